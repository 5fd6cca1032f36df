// test hooks (see debug_kernels.cuh)
#define G16_COLD 1
#define G16_PAIR_DEBUG_KERNEL 1
#include "kernel_impl.cuh"
#include "debug_kernels.cuh"
namespace g16 {
void k_debug_fq_op(stream_t s, size_t n, int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    launch<DebugFqOp>(n, s, op, a, b, out);
}
void k_debug_fr_from_mont(stream_t s, size_t n, const uint32_t *a, uint32_t *out) { launch<DebugFrFromMont>(n, s, a, out); }
template <class F>
void k_debug_add(stream_t s, size_t n, const uint32_t *p, const uint32_t *q, uint32_t *out) {
#if !defined(G16_EMU) && !G16_G2_ACC_THREAD
    if constexpr (std::is_same<F, Fq2>::value) {   // G2 on the device: the lane-pair addition of the hot kernel (pair_g2.cuh)
        if (n == 0) return;
        debug_pair_add_kernel<<<(unsigned)((2 * n + 63) / 64), 64, 0, s>>>(n, p, q, out);
        G16_CUDA_CHECK(cudaGetLastError());
        note_launch();
        return;
    }
#endif
    launch<DebugAdd<F>>(n, s, p, q, out);
}
template void k_debug_add<Fq>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t *);
template void k_debug_add<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t *);
}  // namespace g16
