// cold kernel: bucket reduction level (g2)
#define G16_COLD 1
#include "kernel_impl.cuh"
namespace g16 {
template void k_tile_reduce<Fq2>(stream_t, uint32_t, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *, uint32_t *);
template void k_reduce_level<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *); }
