// bucket reduction (g2): Fq multiplication inline, Fq2 multiplication / squaring as calls (g16_defs.cuh)
#define G16_COLD_FQ2 1
#include "kernel_impl.cuh"
namespace g16 {
template void k_tile_reduce<Fq2>(stream_t, uint32_t, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *, uint32_t *);
template void k_reduce_level<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *); }
