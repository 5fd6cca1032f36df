// bucket reduction (g2): Fq multiplication inline, Fq2 multiplication / squaring as calls (g16_defs.cuh), and the
// Karatsuba form of the Fq2 multiplication: behind a call boundary it beats the two-product form here (reduce 4.77 ms
// against 5.29 ms at 2^19 buckets, profiles/r02_run21_lab_g2_dual_per_kernel.txt)
#define G16_COLD_FQ2 1
#define G16_FQ2_DUAL 0
#define G16_FQ2_QUAD 0
#include "kernel_impl.cuh"
namespace g16 {
template void k_tile_reduce<Fq2>(stream_t, uint32_t, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *, uint32_t *);
template void k_reduce_level<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t *, uint32_t *); }
