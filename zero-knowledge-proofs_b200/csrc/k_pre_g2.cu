// base-point precomputation 2^(c w) P (g2); one-time work at upload
#define G16_COLD 1
#include "kernel_impl.cuh"
namespace g16 { template void k_precompute_bases<Fq2>(stream_t, size_t, const uint32_t *, uint32_t, uint32_t, uint32_t *); }
