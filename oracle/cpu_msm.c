/* CPU oracle / CPU baseline -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load the library built from this file (oracle/_build/liboracle.so).
 *
 * What it restates: the CPU path the reference executes behind
 *   Prover::multi_scalar_mult_g1 / _g2      /root/reference/crates/groth16-core/src/lib.rs:275-300
 *   CRS::generate_from_qap (group part)     /root/reference/crates/groth16-setup/src/lib.rs:162-241
 * i.e. ark-ec 0.4.2 `VariableBaseMSM::msm` (msm_bigint_wnaf, sequential) + `into_affine`,
 * and `Projective * Fr` + `into_affine` per element.  ark-ec/ark-ff are third-party crates
 * that are not vendored in /root/reference and there is no Rust toolchain here, so this is a
 * "port" baseline (cpu_baseline.kind = "port"), validated against oracle/bls12_381.py.
 * PARITY UNPINNED by the reference (no golden vectors exist upstream).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "cpu_field.h"

/* ark_std::log2 = ceil(log2(x)); ln_without_floats(a) = log2(a) * 69 / 100 */
static size_t ark_log2(size_t x) {
    if (x <= 1) return 0;
    size_t n = 0, v = x - 1;
    while (v) { ++n; v >>= 1; }
    return n;
}
static size_t ln_without_floats(size_t a) { return ark_log2(a) * 69 / 100; }

/* ark-ec 0.4.2 make_digits: signed radix-2^w digits in [-2^(w-1), 2^(w-1)) */
static void make_digits(int64_t *digits, const uint64_t a[4], int w, int num_bits) {
    const uint64_t radix = 1ULL << w, window_mask = radix - 1;
    uint64_t carry = 0;
    int digits_count = (num_bits + w - 1) / w;
    for (int i = 0; i < digits_count; ++i) {
        int bit_offset = i * w, u64_idx = bit_offset / 64, bit_idx = bit_offset % 64;
        uint64_t bit_buf;
        if (bit_idx < 64 - w || u64_idx == 3) bit_buf = a[u64_idx] >> bit_idx;
        else bit_buf = (a[u64_idx] >> bit_idx) | (a[u64_idx + 1] << (64 - bit_idx));
        uint64_t coef = carry + (bit_buf & window_mask);
        carry = (coef + radix / 2) >> w;
        digits[i] = (int64_t)coef - (int64_t)(carry << w);
    }
    digits[digits_count - 1] += (int64_t)(carry << w);
}

#define FE fq_t
#define F(n) fq_##n
#define C(n) g1_##n
#include "cpu_curve_tmpl.h"
#undef FE
#undef F
#undef C
#define FE fq2_t
#define F(n) fq2_##n
#define C(n) g2_##n
#include "cpu_curve_tmpl.h"
#undef FE
#undef F
#undef C

/* ------------------------------------------------------------------ marshalling
 * Same packed layout as include/g16_cuda.h: G1 point = 12 u64 (x[6], y[6]) Montgomery LE,
 * G2 point = 24 u64 (x.c0, x.c1, y.c0, y.c1), infinity flags in a separate byte array
 * (NULL = none), scalars = 4 u64 Montgomery LE.  */
static void g1_load(g1_aff_t *p, const uint64_t *xy, const uint8_t *inf, size_t i) {
    memcpy(p->x.l, xy + 12 * i, 48); memcpy(p->y.l, xy + 12 * i + 6, 48); p->inf = inf ? inf[i] != 0 : 0;
}
static void g1_store(const g1_aff_t *p, uint64_t *xy, uint8_t *inf, size_t i) {
    memcpy(xy + 12 * i, p->x.l, 48); memcpy(xy + 12 * i + 6, p->y.l, 48); if (inf) inf[i] = (uint8_t)p->inf;
}
static void g2_load(g2_aff_t *p, const uint64_t *xy, const uint8_t *inf, size_t i) {
    memcpy(&p->x, xy + 24 * i, 96); memcpy(&p->y, xy + 24 * i + 12, 96); p->inf = inf ? inf[i] != 0 : 0;
}
static void g2_store(const g2_aff_t *p, uint64_t *xy, uint8_t *inf, size_t i) {
    memcpy(xy + 24 * i, &p->x, 96); memcpy(xy + 24 * i + 12, &p->y, 96); if (inf) inf[i] = (uint8_t)p->inf;
}

#define DEFINE_API(G, LOAD, STORE)                                                                   \
    int ora_##G##_msm(const uint64_t *bases, const uint8_t *inf, const uint64_t *scalars_mont, size_t n, \
                      int threads, uint64_t *out_xy, uint8_t *out_inf) {                             \
        G##_aff_t *pts = (G##_aff_t *)malloc(sizeof(G##_aff_t) * (n ? n : 1));                       \
        uint64_t *big = (uint64_t *)malloc(32 * (n ? n : 1));                                        \
        if (!pts || !big) { free(pts); free(big); return -1; }                                       \
        for (size_t i = 0; i < n; ++i) { LOAD(&pts[i], bases, inf, i); fr_from_mont(big + 4 * i, scalars_mont + 4 * i); } \
        G##_jac_t acc; int rc = G##_msm_bigint_wnaf(&acc, pts, big, n, threads < 1 ? 1 : threads);   \
        G##_aff_t res; G##_jac_to_affine(&res, &acc); STORE(&res, out_xy, out_inf, 0);               \
        free(pts); free(big); return rc;                                                             \
    }                                                                                                \
    int ora_##G##_fixed_base_mul(const uint64_t *base_xy, const uint64_t *scalars_mont, size_t n,    \
                                 int threads, uint64_t *out_xy, uint8_t *out_inf) {                  \
        G##_aff_t base; LOAD(&base, base_xy, NULL, 0);                                               \
        (void)threads;                                                                               \
        _Pragma("omp parallel for schedule(static) num_threads(threads) if (threads > 1)")           \
        for (size_t i = 0; i < n; ++i) {                                                             \
            uint64_t k[4]; fr_from_mont(k, scalars_mont + 4 * i);                                    \
            G##_jac_t acc; G##_mul_bigint(&acc, &base, k);                                           \
            G##_aff_t res; G##_jac_to_affine(&res, &acc); STORE(&res, out_xy, out_inf, i);           \
        }                                                                                            \
        return 0;                                                                                    \
    }                                                                                                \
    /* naive sum_i s_i P_i by double-and-add: independent of the bucket method */                    \
    int ora_##G##_msm_naive(const uint64_t *bases, const uint8_t *inf, const uint64_t *scalars_mont, \
                            size_t n, uint64_t *out_xy, uint8_t *out_inf) {                          \
        G##_jac_t acc; G##_jac_set_zero(&acc);                                                       \
        for (size_t i = 0; i < n; ++i) {                                                             \
            G##_aff_t p; LOAD(&p, bases, inf, i);                                                    \
            uint64_t k[4]; fr_from_mont(k, scalars_mont + 4 * i);                                    \
            G##_jac_t t; G##_mul_bigint(&t, &p, k); G##_jac_add(&acc, &t);                           \
        }                                                                                            \
        G##_aff_t res; G##_jac_to_affine(&res, &acc); STORE(&res, out_xy, out_inf, 0);               \
        return 0;                                                                                    \
    }

DEFINE_API(g1, g1_load, g1_store)
DEFINE_API(g2, g2_load, g2_store)

/* ------------------------------------------------------------------ small helpers for parity tests */
void ora_fq_mul(const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; ++i) mont_mul(r + 6 * i, a + 6 * i, b + 6 * i, FQ_P, FQ_NINV, 6);
}
void ora_fq_add(const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; ++i) mod_add(r + 6 * i, a + 6 * i, b + 6 * i, FQ_P, 6);
}
void ora_fq_sub(const uint64_t *a, const uint64_t *b, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; ++i) mod_sub(r + 6 * i, a + 6 * i, b + 6 * i, FQ_P, 6);
}
void ora_fq_inv(const uint64_t *a, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; ++i) { fq_t x, y; memcpy(x.l, a + 6 * i, 48); fq_inv(&y, &x); memcpy(r + 6 * i, y.l, 48); }
}
void ora_fr_from_mont(const uint64_t *a, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; ++i) fr_from_mont(r + 4 * i, a + 4 * i);
}
void ora_fr_to_mont(const uint64_t *a, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; ++i) fr_to_mont(r + 4 * i, a + 4 * i);
}
/* sum_i a_i * b_i mod r, all values Fr in Montgomery form (the "discrete-log" side of the exact
 * large-size checks: with bases P_i = k_i G, sum s_i P_i = (sum s_i k_i) G).  Result in Montgomery form. */
void ora_dot_mod_r(const uint64_t *a_mont, const uint64_t *b_mont, size_t n, int threads, uint64_t out_mont[4]) {
    if (threads < 1) threads = 1;
    uint64_t *part = (uint64_t *)calloc((size_t)threads * 4, 8);
    _Pragma("omp parallel num_threads(threads) if (threads > 1)")
    {
#ifdef _OPENMP
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int t = 0, nt = 1;
#endif
        uint64_t acc[4] = {0, 0, 0, 0}, prod[4];
        size_t lo = n * (size_t)t / (size_t)nt, hi = n * ((size_t)t + 1) / (size_t)nt;
        for (size_t i = lo; i < hi; ++i) {
            mont_mul(prod, a_mont + 4 * i, b_mont + 4 * i, FR_P, FR_NINV, 4);
            mod_add(acc, acc, prod, FR_P, 4);
        }
        memcpy(part + 4 * t, acc, 32);
    }
    uint64_t acc[4] = {0, 0, 0, 0};
    for (int t = 0; t < threads; ++t) mod_add(acc, acc, part + 4 * t, FR_P, 4);
    memcpy(out_mont, acc, 32);
    free(part);
}
int ora_msm_window(size_t n) { return n < 32 ? 3 : (int)ln_without_floats(n) + 2; }
void ora_make_digits(const uint64_t a[4], int w, int64_t *digits) { make_digits(digits, a, w, 255); }
int ora_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* SplitMix64-driven inputs, identical to oracle/bls12_381.py:SplitMix64 */
static uint64_t splitmix_next(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* n scalars uniform in [0, min(r, 2^bits)), written in Montgomery form */
void ora_gen_scalars(uint64_t seed, int bits, size_t n, uint64_t *out_mont) {
    uint64_t s = seed;
    for (size_t i = 0; i < n; ++i) {
        uint64_t v[4];
        for (;;) {
            for (int k = 0; k < 4; ++k) v[k] = splitmix_next(&s);
            for (int k = 0; k < 4; ++k) {
                int lo = 64 * k;
                if (bits <= lo) v[k] = 0;
                else if (bits < lo + 64) v[k] &= (1ULL << (bits - lo)) - 1;
            }
            if (!mp_geq(v, FR_P, 4)) break;
        }
        fr_to_mont(out_mont + 4 * i, v);
    }
}
