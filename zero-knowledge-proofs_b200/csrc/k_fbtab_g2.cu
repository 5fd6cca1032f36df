// cold kernels: fixed-base window table (g2)
#define G16_COLD 1
#include "kernel_impl.cuh"
namespace g16 {
template void k_fb_powers<Fq2>(stream_t, const uint32_t *, uint32_t *);
template void k_fb_table<Fq2>(stream_t, const uint32_t *, uint32_t *);
}
