// test hooks (see debug_kernels.cuh)
#define G16_COLD 1
#include "kernel_impl.cuh"
#include "debug_kernels.cuh"
namespace g16 {
void k_debug_fq_op(stream_t s, size_t n, int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    launch<DebugFqOp>(n, s, op, a, b, out);
}
void k_debug_fr_from_mont(stream_t s, size_t n, const uint32_t *a, uint32_t *out) { launch<DebugFrFromMont>(n, s, a, out); }
template <class F>
void k_debug_add(stream_t s, size_t n, const uint32_t *p, const uint32_t *q, uint32_t *out) {
    launch<DebugAdd<F>>(n, s, p, q, out);
}
template void k_debug_add<Fq>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t *);
template void k_debug_add<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t *);
}  // namespace g16
