// base-point precomputation 2^(c w) P (g2); one-time work at upload
#define G16_COLD 1
#define G16_FQ2_DUAL 1   // 372 ms against 404 ms with Karatsuba for 2^20 bases (profiles/r02_run21_lab_g2_dual_per_kernel.txt)
#include "kernel_impl.cuh"
namespace g16 { template void k_precompute_bases<Fq2>(stream_t, size_t, const uint32_t *, uint32_t, uint32_t, uint32_t *); }
