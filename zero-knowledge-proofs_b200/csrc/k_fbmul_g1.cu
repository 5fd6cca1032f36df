// fixed-base batch scalar multiplication (g1)
// hot: one inlined mixed addition per window (setup throughput)
#include "kernel_impl.cuh"
namespace g16 { template void k_fb_mul<Fq>(stream_t, size_t, const uint32_t *, bool, const uint32_t *, uint32_t *); }
