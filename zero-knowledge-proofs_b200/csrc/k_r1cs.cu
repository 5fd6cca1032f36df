// sparse R1CS kernels (Fr): matrix-vector products, Lagrange basis at s, CRS exponents
#include "r1cs_kernels.cuh"
namespace g16 {
void k_spmv(stream_t s, size_t lines, const uint32_t *line_ptr, const uint32_t *idx, const uint32_t *val, const uint32_t *vec,
            uint32_t seg, uint32_t seg_out, const uint32_t *long_lines, size_t n_long, uint32_t *out) {
    launch<SpmvThread>(lines, s, line_ptr, idx, val, vec, seg, seg_out, out);
    if (!n_long) return;
#ifndef G16_EMU
    spmv_long_kernel<<<(unsigned)n_long, SPMV_LONG_THREADS, 0, s>>>(long_lines, line_ptr, idx, val, vec, seg, seg_out, out);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
#else
    launch<SpmvLongSerial>(n_long, s, long_lines, line_ptr, idx, val, vec, seg, seg_out, out);
#endif
}
uint32_t k_spmv_long_threshold() { return SPMV_LONG; }
void k_truncate64(stream_t s, size_t n, const uint32_t *in, uint32_t *out) { launch<Truncate64>(n, s, in, out); }
size_t k_setup_scalar_words() { return SETUP_SCALARS * 8; }
void k_setup_scalars(stream_t s, const uint32_t *params, const uint32_t *consts, uint32_t log_n, uint32_t truncate, uint32_t *blk) {
    launch<SetupScalars>(1, s, params, consts, log_n, truncate, blk);
}
void k_lagrange_at(stream_t s, size_t n, const uint32_t *consts, const uint32_t *blk, uint32_t *out) {
    launch<LagrangeAt>(n, s, consts, blk, out);
}
void k_crs_exponents(stream_t s, size_t num_vars, const uint32_t *vals, const uint32_t *blk, uint32_t num_public, uint32_t *ab,
                     uint32_t *ic) {
    launch<CrsExponents>(num_vars, s, vals, blk, (uint32_t)num_vars, num_public, ab, ic);
}
void k_crs_h_exponents(stream_t s, size_t n, const uint32_t *blk, uint32_t *h) { launch<CrsHExponents>(n, s, blk, h); }
void k_validate_row(stream_t s, const uint32_t *abc, uint32_t n, uint32_t *flag) { launch<ValidateRow>(1, s, abc, n, flag); }
}  // namespace g16
