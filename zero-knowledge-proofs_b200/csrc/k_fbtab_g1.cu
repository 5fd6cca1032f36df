// cold kernels: fixed-base window table (g1)
#define G16_COLD 1
#include "kernel_impl.cuh"
namespace g16 {
template void k_fb_powers<Fq>(stream_t, const uint32_t *, uint32_t *);
template void k_fb_table<Fq>(stream_t, const uint32_t *, uint32_t *);
}
