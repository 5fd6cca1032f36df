// base-point precomputation 2^(c w) P (g1); one-time work at upload
#include "kernel_impl.cuh"
namespace g16 { template void k_precompute_bases<Fq>(stream_t, size_t, const uint32_t *, uint32_t, uint32_t, uint32_t *); }
