// cold kernel: bucket reduction level (g1)
// hot: field multiply inlined (G1 tail latency matters at small N and in strong scaling)
#include "kernel_impl.cuh"
namespace g16 {
template void k_tile_reduce<Fq>(stream_t, uint32_t, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *, uint32_t *);
template void k_reduce_level<Fq>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *); }
