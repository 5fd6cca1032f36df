// cold kernels: window fold, partial-sum combine, base import/export (g1)
// hot: field multiply inlined (G1 tail latency matters at small N and in strong scaling)
#include "kernel_impl.cuh"
namespace g16 {
template void k_window_combine<Fq>(stream_t, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t *, uint32_t *);
template void k_partial_combine<Fq>(stream_t, const uint32_t *, uint32_t, uint32_t *, uint32_t *);
template void k_scalar_mul_affine<Fq>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t, uint32_t *);
template void k_chunk_merge<Fq>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t *, bool, uint32_t);
template void k_import_bases<Fq>(stream_t, size_t, const uint32_t *, const uint8_t *, uint32_t *);
template void k_export_flags<Fq>(stream_t, size_t, const uint32_t *, uint8_t *);
}
