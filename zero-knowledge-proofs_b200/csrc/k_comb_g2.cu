// cold kernels: window fold, partial-sum combine, base import/export (g2)
#define G16_COLD 1
#include "kernel_impl.cuh"
namespace g16 {
template void k_window_combine<Fq2>(stream_t, const uint32_t *, const uint32_t *, const uint32_t *, uint32_t, uint32_t, uint32_t *, uint32_t *);
template void k_partial_combine<Fq2>(stream_t, const uint32_t *, uint32_t, uint32_t *, uint32_t *);
template void k_chunk_merge<Fq2>(stream_t, size_t, const uint32_t *, const uint32_t *, uint32_t *, bool, uint32_t);
template void k_import_bases<Fq2>(stream_t, size_t, const uint32_t *, const uint8_t *, uint32_t *);
template void k_export_flags<Fq2>(stream_t, size_t, const uint32_t *, uint8_t *);
}
