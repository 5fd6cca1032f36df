// hot kernel: bucket accumulation (g2)
#include "kernel_impl.cuh"
namespace g16 {
template void k_accumulate<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, const WorkItem *, const uint32_t *, uint32_t *, uint32_t *, bool);
}
