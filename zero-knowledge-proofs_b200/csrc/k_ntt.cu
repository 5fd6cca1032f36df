// quotient polynomial kernels (Fr NTT)
#include "ntt_kernels.cuh"
namespace g16 {
size_t k_ntt_const_words() { return NTT_CONST_WORDS; }
void k_ntt_setup(stream_t s, uint32_t log_n, uint32_t *consts) { launch<NttSetup>(1, s, log_n, consts); }
void k_ntt_twiddles(stream_t s, uint32_t n, const uint32_t *consts, uint32_t *tw, uint32_t *twi) {
    launch<NttTwiddles>(n / 2, s, consts, tw, twi);
}
void k_ntt_stage(stream_t s, bool dit, size_t batch, uint32_t *x, const uint32_t *tw, uint32_t n, uint32_t half) {
    if (dit) launch<NttStageDit>(batch * (n / 2), s, x, tw, n, half);
    else launch<NttStageDif>(batch * (n / 2), s, x, tw, n, half);
}
void k_ntt_quotient_pointwise(stream_t s, uint32_t *abc, const uint32_t *consts, uint32_t n) {
    launch<NttQuotientPointwise>(n, s, abc, consts, n);
}
void k_ntt_check_vanish(stream_t s, const uint32_t *abc, uint32_t n, uint32_t *flag) { launch<NttCheckVanish>(n, s, abc, n, flag); }
void k_ntt_scale_tables(stream_t s, uint32_t n, uint32_t log_n, const uint32_t *consts, uint32_t *scale, uint32_t *fscale) {
    launch<NttCosetTable>(n, s, consts, log_n, scale);
    launch<NttFinalTable>(n, s, consts, fscale);
}
void k_ntt_final_permute(stream_t s, const uint32_t *x, const uint32_t *fscale, uint32_t n, uint32_t log_n, uint32_t *out) {
    launch<NttFinalPermute>(n, s, x, fscale, log_n, out);
}
// One whole transform of `batch` arrays: DIF (natural in, bit-reversed out; tw = inverse twiddles) or DIT (bit-reversed in,
// natural out; tw = forward twiddles).  scale (DIF only, may be null): per-position factors applied with the last pass.
void k_ntt_transform(stream_t s, bool dit, size_t batch, uint32_t *x, const uint32_t *tw, uint32_t log_n, const uint32_t *scale) {
    uint32_t n = 1u << log_n;
#ifndef G16_EMU
    if (log_n == 0) {
        if (scale && !dit) launch<NttCosetScaleTable>(batch * n, s, x, scale, n);
        return;
    }
    // passes of at most NTT_FUSED_MAX_STAGES stages, sizes as even as possible; pass i covers stages k0[i] .. k0[i] + S[i] - 1
    uint32_t P = (log_n + NTT_FUSED_MAX_STAGES - 1) / NTT_FUSED_MAX_STAGES, S[32], k0[32];
    for (uint32_t i = 0, k = 0; i < P; ++i) {
        S[i] = log_n / P + (i < log_n % P ? 1u : 0u);
        k0[i] = k;
        k += S[i];
    }
    for (uint32_t step = 0; step < P; ++step) {
        uint32_t i = dit ? P - 1 - step : step;                 // DIT runs the stages from the innermost pass outwards
        uint32_t log_stride = log_n - k0[i] - S[i];
        uint32_t log_c = log_stride == 0 ? k0[i] : log_stride;
        if (log_c > 3) log_c = 3;                               // NTT_FUSED_MAX_COLS = 8
        uint32_t tiles = n >> (S[i] + log_c);
        size_t smem = ((size_t)1 << (S[i] + log_c)) * 32;
        bool last_dif = !dit && i == P - 1;
        ntt_fused_kernel<<<(unsigned)(batch * tiles), NTT_FUSED_THREADS, smem, s>>>(x, tw, log_n, k0[i], S[i], log_c, tiles, dit ? 1 : 0,
                                                                                  last_dif ? scale : nullptr);
        G16_CUDA_CHECK(cudaGetLastError());
        note_launch();
    }
#else
    if (dit) for (uint32_t half = 1; half < n; half <<= 1) k_ntt_stage(s, true, batch, x, tw, n, half);
    else for (uint32_t half = n / 2; half >= 1; half >>= 1) k_ntt_stage(s, false, batch, x, tw, n, half);
    if (scale && !dit) launch<NttCosetScaleTable>(batch * n, s, x, scale, n);
#endif
}
}  // namespace g16
