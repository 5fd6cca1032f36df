/* CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle/README.md.
 *
 * 64-bit-limb Montgomery arithmetic for BLS12-381 Fq (6 limbs), Fr (4 limbs) and
 * Fq2 = Fq[u]/(u^2+1).  Restates the published algorithm of ark-ff 0.4.2 `MontBackend`
 * (third-party, not vendored; pinned in /root/reference/Cargo.lock:118-119), i.e. the
 * representation `a*R mod p` in little-endian u64 limbs that the reference passes around
 * at crates/groth16-core/src/lib.rs:275-300 and crates/groth16-setup/src/lib.rs:166-241.
 * Values are always kept fully reduced (< p), as ark does.
 */
#ifndef ORA_CPU_FIELD_H
#define ORA_CPU_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fq_t;
typedef struct { uint64_t l[4]; } fr_t;
typedef struct { fq_t c0, c1; } fq2_t;

static const uint64_t FQ_P[6] = {
    0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
    0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const uint64_t FQ_NINV = 0x89f3fffcfffcfffdULL;
/* R mod q (Montgomery one) and R^2 mod q */
static const uint64_t FQ_ONE[6] = {
    0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
    0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL};
static const uint64_t FQ_R2[6] = {
    0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
    0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL};

static const uint64_t FR_P[4] = {
    0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t FR_NINV = 0xfffffffeffffffffULL;
static const uint64_t FR_ONE[4] = {
    0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL};
static const uint64_t FR_R2[4] = {
    0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL};

/* ---- generic n-limb helpers (n is a compile-time constant at every call site) ---- */
static inline int mp_geq(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; --i) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline uint64_t mp_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    u128 c = 0;
    for (int i = 0; i < n; ++i) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t mp_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    uint64_t borrow = 0;
    for (int i = 0; i < n; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (uint64_t)d; borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
static inline int mp_is_zero(const uint64_t *a, int n) {
    uint64_t v = 0; for (int i = 0; i < n; ++i) v |= a[i]; return v == 0;
}
static inline void mod_add(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *p, int n) {
    uint64_t c = mp_add(r, a, b, n);
    if (c || mp_geq(r, p, n)) mp_sub(r, r, p, n);
}
static inline void mod_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *p, int n) {
    if (mp_sub(r, a, b, n)) mp_add(r, r, p, n);
}
static inline void mod_neg(uint64_t *r, const uint64_t *a, const uint64_t *p, int n) {
    if (mp_is_zero(a, n)) { memset(r, 0, 8 * (size_t)n); } else mp_sub(r, p, a, n);
}
/* CIOS Montgomery product r = a*b/R mod p */
static inline __attribute__((always_inline)) void mont_mul(uint64_t *r, const uint64_t *a, const uint64_t *b,
                            const uint64_t *p, uint64_t ninv, int n) {
    uint64_t t[8] = {0};
    for (int i = 0; i < n; ++i) {
        u128 c = 0;
        for (int j = 0; j < n; ++j) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[n]; t[n] = (uint64_t)c; t[n + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * ninv;
        c = (u128)m * p[0] + t[0]; c >>= 64;
        for (int j = 1; j < n; ++j) { c += (u128)m * p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[n]; t[n - 1] = (uint64_t)c; t[n] = t[n + 1] + (uint64_t)(c >> 64);
    }
    if (t[n] || mp_geq(t, p, n)) mp_sub(t, t, p, n);
    memcpy(r, t, 8 * (size_t)n);
}

/* fully unrolled 6-limb CIOS for Fq (same result as mont_mul(...,6); ~1.5x faster) */
static inline __attribute__((always_inline)) void fq_mont_mul6(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0, t7;
#define FQ_ROUND(bi) { u128 c; uint64_t m; \
    c = (u128)a[0] * (bi) + t0; t0 = (uint64_t)c; c >>= 64; \
    c += (u128)a[1] * (bi) + t1; t1 = (uint64_t)c; c >>= 64; \
    c += (u128)a[2] * (bi) + t2; t2 = (uint64_t)c; c >>= 64; \
    c += (u128)a[3] * (bi) + t3; t3 = (uint64_t)c; c >>= 64; \
    c += (u128)a[4] * (bi) + t4; t4 = (uint64_t)c; c >>= 64; \
    c += (u128)a[5] * (bi) + t5; t5 = (uint64_t)c; c >>= 64; \
    c += t6; t6 = (uint64_t)c; t7 = (uint64_t)(c >> 64); \
    m = t0 * FQ_NINV; \
    c = (u128)m * FQ_P[0] + t0; c >>= 64; \
    c += (u128)m * FQ_P[1] + t1; t0 = (uint64_t)c; c >>= 64; \
    c += (u128)m * FQ_P[2] + t2; t1 = (uint64_t)c; c >>= 64; \
    c += (u128)m * FQ_P[3] + t3; t2 = (uint64_t)c; c >>= 64; \
    c += (u128)m * FQ_P[4] + t4; t3 = (uint64_t)c; c >>= 64; \
    c += (u128)m * FQ_P[5] + t5; t4 = (uint64_t)c; c >>= 64; \
    c += t6; t5 = (uint64_t)c; t6 = t7 + (uint64_t)(c >> 64); }
    uint64_t b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3], b4 = b[4], b5 = b[5];
    FQ_ROUND(b0) FQ_ROUND(b1) FQ_ROUND(b2) FQ_ROUND(b3) FQ_ROUND(b4) FQ_ROUND(b5)
#undef FQ_ROUND
    uint64_t t[6] = {t0, t1, t2, t3, t4, t5}, s[6], br = 0;
    for (int i = 0; i < 6; ++i) { u128 d = (u128)t[i] - FQ_P[i] - br; s[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    int ge = t6 || !br;
    for (int i = 0; i < 6; ++i) r[i] = ge ? s[i] : t[i];
}

/* ---- Fq ---- */
static inline void fq_add(fq_t *r, const fq_t *a, const fq_t *b) { mod_add(r->l, a->l, b->l, FQ_P, 6); }
static inline void fq_sub(fq_t *r, const fq_t *a, const fq_t *b) { mod_sub(r->l, a->l, b->l, FQ_P, 6); }
static inline void fq_neg(fq_t *r, const fq_t *a) { mod_neg(r->l, a->l, FQ_P, 6); }
static inline void fq_dbl(fq_t *r, const fq_t *a) { mod_add(r->l, a->l, a->l, FQ_P, 6); }
static inline void fq_mul(fq_t *r, const fq_t *a, const fq_t *b) { fq_mont_mul6(r->l, a->l, b->l); }
static inline void fq_sqr(fq_t *r, const fq_t *a) { fq_mont_mul6(r->l, a->l, a->l); }
static inline int fq_is_zero(const fq_t *a) { return mp_is_zero(a->l, 6); }
static inline int fq_eq(const fq_t *a, const fq_t *b) { return memcmp(a, b, sizeof(fq_t)) == 0; }
static inline void fq_set_zero(fq_t *r) { memset(r, 0, sizeof *r); }
static inline void fq_set_one(fq_t *r) { memcpy(r->l, FQ_ONE, sizeof r->l); }
/* a^(q-2): Fermat inversion (ark uses binary EGCD; the value is the same) */
static inline void fq_inv(fq_t *r, const fq_t *a) {
    uint64_t e[6]; memcpy(e, FQ_P, sizeof e); e[0] -= 2;
    fq_t acc; fq_set_one(&acc);
    for (int i = 383; i >= 0; --i) {
        fq_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fq_mul(&acc, &acc, a);
    }
    *r = acc;
}

/* ---- Fq2 ---- */
static inline void fq2_add(fq2_t *r, const fq2_t *a, const fq2_t *b) { fq_add(&r->c0, &a->c0, &b->c0); fq_add(&r->c1, &a->c1, &b->c1); }
static inline void fq2_sub(fq2_t *r, const fq2_t *a, const fq2_t *b) { fq_sub(&r->c0, &a->c0, &b->c0); fq_sub(&r->c1, &a->c1, &b->c1); }
static inline void fq2_neg(fq2_t *r, const fq2_t *a) { fq_neg(&r->c0, &a->c0); fq_neg(&r->c1, &a->c1); }
static inline void fq2_dbl(fq2_t *r, const fq2_t *a) { fq_dbl(&r->c0, &a->c0); fq_dbl(&r->c1, &a->c1); }
static inline void fq2_mul(fq2_t *r, const fq2_t *a, const fq2_t *b) {
    fq_t v0, v1, s, t;
    fq_mul(&v0, &a->c0, &b->c0); fq_mul(&v1, &a->c1, &b->c1);
    fq_add(&s, &a->c0, &a->c1); fq_add(&t, &b->c0, &b->c1);
    fq_mul(&s, &s, &t); fq_sub(&s, &s, &v0); fq_sub(&s, &s, &v1);
    fq_sub(&r->c0, &v0, &v1); r->c1 = s;
}
static inline void fq2_sqr(fq2_t *r, const fq2_t *a) {
    fq_t s, d, m;
    fq_add(&s, &a->c0, &a->c1); fq_sub(&d, &a->c0, &a->c1); fq_mul(&m, &a->c0, &a->c1);
    fq_mul(&r->c0, &s, &d); fq_dbl(&r->c1, &m);
}
static inline int fq2_is_zero(const fq2_t *a) { return fq_is_zero(&a->c0) && fq_is_zero(&a->c1); }
static inline int fq2_eq(const fq2_t *a, const fq2_t *b) { return memcmp(a, b, sizeof(fq2_t)) == 0; }
static inline void fq2_set_zero(fq2_t *r) { memset(r, 0, sizeof *r); }
static inline void fq2_set_one(fq2_t *r) { fq_set_one(&r->c0); fq_set_zero(&r->c1); }
static inline void fq2_inv(fq2_t *r, const fq2_t *a) {
    fq_t n, t; fq_sqr(&n, &a->c0); fq_sqr(&t, &a->c1); fq_add(&n, &n, &t); fq_inv(&n, &n);
    fq_mul(&r->c0, &a->c0, &n); fq_mul(&t, &a->c1, &n); fq_neg(&r->c1, &t);
}

/* ---- Fr: only what the MSM front-end needs ---- */
/* Montgomery -> canonical integer (`into_bigint()`, ark-ec msm_unchecked) */
static inline void fr_from_mont(uint64_t out[4], const uint64_t a[4]) {
    static const uint64_t one[4] = {1, 0, 0, 0};
    mont_mul(out, a, one, FR_P, FR_NINV, 4);
}
static inline void fr_to_mont(uint64_t out[4], const uint64_t a[4]) { mont_mul(out, a, FR_R2, FR_P, FR_NINV, 4); }
#endif
