// Quotient polynomial H = (A*B - C) / Z on the GPU -- SURVEY.md 8(f) rank 2, the stage that feeds the H MSM.
//
// Replaces `QAP::compute_quotient_polynomial` (/root/reference/crates/groth16-qap/src/lib.rs:225-271, called
// from Prover::prove at crates/groth16-core/src/lib.rs:200): the reference sums dense per-variable polynomials
// (Theta(N*n)), multiplies by FFT and divides by Z(x) = x^n - 1.  Given the evaluations of A, B, C on the
// radix-2 domain (a_i = <A-row i, w>, which a sparse R1CS yields in O(nnz)), the same H is obtained with
// seven size-n NTTs over Fr:
//     a, b, c  --iNTT-->  coefficients  --x g^j, NTT-->  values on the coset g*<omega>
//     h~_i = (a~_i b~_i - c~_i) / (g^n - 1)            (Z is the constant g^n - 1 on the coset)
//     h~  --iNTT, x g^-j-->  coefficients of H  (degree <= n - 2, so n coset points determine it)
// Domain = ark-poly's Radix2EvaluationDomain: omega = TWO_ADIC_ROOT^(2^(32 - log n)), TWO_ADIC_ROOT =
// 7^((r-1)/2^32); coset generator g = 7 (any g outside the domain gives the same H).
//
// Transforms: decimation-in-frequency forward (natural in, bit-reversed out) and decimation-in-time
// (bit-reversed in, natural out), so no separate permutation pass is needed; one kernel per stage, one
// thread per butterfly (the stages stream 64 B per butterfly: HBM bound).
#pragma once
#include "fp.cuh"
#include "kernel_api.cuh"

namespace g16 {

G16_HD Fr fr_load(const uint32_t *p, size_t i) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = p[8 * i + k];
    return r;
}
G16_HD void fr_store(uint32_t *p, size_t i, const Fr &v) {
#pragma unroll
    for (int k = 0; k < 8; ++k) p[8 * i + k] = v.l[k];
}
G16_HD uint32_t bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t b = 0; b < bits; ++b) { r = (r << 1) | (x & 1u); x >>= 1; }
    return r;
}

// Layout of the constant block (Fr elements, Montgomery form):
//   [0, 32)   w2[j]  = omega^(2^j)          [32, 64)  wi2[j] = omega^-(2^j)
//   [64, 96)  g2[j]  = g^(2^j)              [96, 128) gi2[j] = g^-(2^j)
//   128: n^-1      129: (g^n - 1)^-1
constexpr uint32_t NTT_CONST_WORDS = 130 * 8;
// Fr(7) and TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32) in Montgomery form (ark-bls12-381 Fr)
G16_HD Fr fr_small(uint32_t v) {
    Fr x = Fr::zero();
    x.l[0] = v;
    return Fr::to_mont(x);
}

struct NttSetup {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, uint32_t log_n, uint32_t *consts) {
        // TWO_ADIC_ROOT = 7^((r-1) >> 32): exponent = (r - 1) / 2^32, square-and-multiply over its 223 bits
        Fr g = fr_small(7);
        uint32_t e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = FrParams::MOD(k);
        e[0] -= 1u;   // r - 1 (r is odd, no borrow)
        Fr root = Fr::one();
        for (int bit = 255; bit >= 32; --bit) {
            root = Fr::sqr(root);
            if ((e[bit >> 5] >> (bit & 31)) & 1u) root = Fr::mul(root, g);
        }
        // omega = root^(2^(32 - log_n))
        Fr w = root;
        for (uint32_t k = log_n; k < 32; ++k) w = Fr::sqr(w);
        Fr wi = Fr::inv(w), gi = Fr::inv(g);
        Fr a = w, b = wi, c = g, d = gi;
        for (uint32_t j = 0; j < 32; ++j) {
            fr_store(consts, j, a); fr_store(consts, 32 + j, b); fr_store(consts, 64 + j, c); fr_store(consts, 96 + j, d);
            a = Fr::sqr(a); b = Fr::sqr(b); c = Fr::sqr(c); d = Fr::sqr(d);
        }
        // n^-1 and (g^n - 1)^-1 ;  g^n = g2[log_n]
        Fr n_fr = Fr::zero();
        n_fr.l[log_n >> 5] = 1u << (log_n & 31);
        fr_store(consts, 128, Fr::inv(Fr::to_mont(n_fr)));
        Fr gn = fr_load(consts, 64 + log_n);
        fr_store(consts, 129, Fr::inv(Fr::sub(gn, Fr::one())));
    }
};

// base^k from the table of base^(2^j)
G16_HD Fr pow_from_table(const uint32_t *consts, uint32_t table, uint32_t k) {
    Fr acc = Fr::one();
    for (uint32_t j = 0; k; ++j, k >>= 1)
        if (k & 1u) acc = Fr::mul(acc, fr_load(consts, table + j));
    return acc;
}

// tw[k] = omega^k, twi[k] = omega^-k for k < n/2
struct NttTwiddles {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t k, const uint32_t *consts, uint32_t *tw, uint32_t *twi) {
        fr_store(tw, k, pow_from_table(consts, 0, (uint32_t)k));
        fr_store(twi, k, pow_from_table(consts, 32, (uint32_t)k));
    }
};

// one DIF stage on `batch` arrays of n elements stored back to back (half = distance of the pair)
struct NttStageDif {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t t, uint32_t *x, const uint32_t *tw, uint32_t n, uint32_t half) {
        size_t arr = t / (n / 2);
        uint32_t q = (uint32_t)(t % (n / 2));
        uint32_t j = q % half, blk = q / half;
        size_t i = arr * n + (size_t)blk * 2 * half + j;
        Fr u = fr_load(x, i), v = fr_load(x, i + half);
        fr_store(x, i, Fr::add(u, v));
        fr_store(x, i + half, Fr::mul(Fr::sub(u, v), fr_load(tw, (size_t)j * (n / (2 * half)))));
    }
};
struct NttStageDit {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t t, uint32_t *x, const uint32_t *tw, uint32_t n, uint32_t half) {
        size_t arr = t / (n / 2);
        uint32_t q = (uint32_t)(t % (n / 2));
        uint32_t j = q % half, blk = q / half;
        size_t i = arr * n + (size_t)blk * 2 * half + j;
        Fr u = fr_load(x, i), v = Fr::mul(fr_load(x, i + half), fr_load(tw, (size_t)j * (n / (2 * half))));
        fr_store(x, i, Fr::add(u, v));
        fr_store(x, i + half, Fr::sub(u, v));
    }
};

// coefficients in bit-reversed order (after the inverse DIF): x[p] *= g^(br(p)) / n
struct NttCosetScale {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t, uint32_t *x, const uint32_t *consts, uint32_t n, uint32_t log_n) {
        uint32_t p = (uint32_t)(t % n);
        Fr s = Fr::mul(pow_from_table(consts, 64, bitrev(p, log_n)), fr_load(consts, 128));
        fr_store(x, t, Fr::mul(fr_load(x, t), s));
    }
};

// h~ = (a~ b~ - c~) / (g^n - 1) into a~   (a, b, c stored back to back)
struct NttQuotientPointwise {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, uint32_t *abc, const uint32_t *consts, uint32_t n) {
        Fr a = fr_load(abc, i), b = fr_load(abc, (size_t)n + i), c = fr_load(abc, 2 * (size_t)n + i);
        fr_store(abc, i, Fr::mul(Fr::sub(Fr::mul(a, b), c), fr_load(consts, 129)));
    }
};

// out[j] = x[br(j)] * g^-j / n   (x = inverse DIF of h~, bit-reversed)
struct NttFinalScale {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t j, const uint32_t *x, const uint32_t *consts, uint32_t log_n, uint32_t *out) {
        Fr s = Fr::mul(pow_from_table(consts, 96, (uint32_t)j), fr_load(consts, 128));
        fr_store(out, j, Fr::mul(fr_load(x, bitrev((uint32_t)j, log_n)), s));
    }
};

// the reference fails with PolynomialDivisionFailed unless A*B - C vanishes on the domain
struct NttCheckVanish {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, const uint32_t *abc, uint32_t n, uint32_t *flag) {
        Fr a = fr_load(abc, i), b = fr_load(abc, (size_t)n + i), c = fr_load(abc, 2 * (size_t)n + i);
        if (Fr::mul(a, b) != c) atomic_add_u32(flag, 1u);   // Montgomery forms: mont(a) * mont(b) = mont(ab)
    }
};

}  // namespace g16
