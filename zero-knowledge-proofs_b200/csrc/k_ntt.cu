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
void k_ntt_coset_scale(stream_t s, size_t batch, uint32_t *x, const uint32_t *consts, uint32_t n, uint32_t log_n) {
    launch<NttCosetScale>(batch * n, s, x, consts, n, log_n);
}
void k_ntt_quotient_pointwise(stream_t s, uint32_t *abc, const uint32_t *consts, uint32_t n) {
    launch<NttQuotientPointwise>(n, s, abc, consts, n);
}
void k_ntt_final_scale(stream_t s, const uint32_t *x, const uint32_t *consts, uint32_t n, uint32_t log_n, uint32_t *out) {
    launch<NttFinalScale>(n, s, x, consts, log_n, out);
}
void k_ntt_check_vanish(stream_t s, const uint32_t *abc, uint32_t n, uint32_t *flag) { launch<NttCheckVanish>(n, s, abc, n, flag); }
}  // namespace g16
