// wire format kernels (cold: out-of-line field multiplications)
#define G16_COLD 1
#include "kernel_impl.cuh"
#include "wire_kernels.cuh"
namespace g16 {
template <class F>
void k_point_encode(stream_t s, size_t n, const uint32_t *pts, bool compressed, uint32_t *out_bytes) {
    launch<PointEncode<F>>(n, s, pts, compressed ? 1u : 0u, out_bytes);
}
template <class F>
void k_point_decode(stream_t s, size_t n, const uint32_t *in_bytes, bool compressed, bool validate, uint32_t *pts, uint8_t *status) {
    launch<PointDecode<F>>(n, s, in_bytes, compressed ? 1u : 0u, validate ? 1u : 0u, pts, status);
}
template void k_point_encode<Fq>(stream_t, size_t, const uint32_t *, bool, uint32_t *);
template void k_point_encode<Fq2>(stream_t, size_t, const uint32_t *, bool, uint32_t *);
template void k_point_decode<Fq>(stream_t, size_t, const uint32_t *, bool, bool, uint32_t *, uint8_t *);
template void k_point_decode<Fq2>(stream_t, size_t, const uint32_t *, bool, bool, uint32_t *, uint8_t *);
}  // namespace g16
