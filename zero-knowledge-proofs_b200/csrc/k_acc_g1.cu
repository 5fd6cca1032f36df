// hot kernel: bucket accumulation (g1)
#include "kernel_impl.cuh"
namespace g16 {
template void k_accumulate<Fq>(stream_t, size_t, const uint32_t *, const uint32_t *, const WorkItem *, const uint32_t *, uint32_t *, uint32_t *, bool);
}
