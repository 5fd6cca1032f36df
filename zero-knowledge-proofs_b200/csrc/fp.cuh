// Montgomery prime-field arithmetic on 32-bit limbs for BLS12-381 Fq (12 limbs) and Fr
// (8 limbs).  Replaces ark-ff 0.4.2 `MontBackend` (6 / 4 x u64 limbs) that runs under every
// group operation of /root/reference/crates/groth16-core/src/lib.rs:282,296 and
// crates/groth16-setup/src/lib.rs:166-241.  Because R = 2^384 (2^256) is the same power of
// two in both limb widths, ark's little-endian u64 limbs reinterpret directly as our u32 limbs.
//
// Representation: fully reduced Montgomery residues (< p) at every function boundary, the
// same invariant ark keeps -- so limbs can be handed back to the host verbatim.
//
// Multiplication: CIOS Montgomery with the even/odd column split so that every partial
// product is a (mad.lo.cc, madc.hi.cc) pair on a 64-bit aligned column; ptxas fuses each
// pair into one IMAD.WIDE.U32 with carry-in/out (see DESIGN.md "Field multiply").
#pragma once
#include "g16_defs.cuh"

namespace g16 {

struct FqParams {
    static constexpr int N = 12;
    static constexpr uint32_t NINV = 0xfffcfffdu;  // -q^-1 mod 2^32
    G16_HD static constexpr uint32_t MOD(int i) {
        constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return m[i];
    }
    G16_HD static constexpr uint32_t ONE(int i) {  // R mod q
        constexpr uint32_t m[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return m[i];
    }
    G16_HD static constexpr uint32_t R2(int i) {  // R^2 mod q
        constexpr uint32_t m[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return m[i];
    }
};

struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t NINV = 0xffffffffu;  // -r^-1 mod 2^32
    G16_HD static constexpr uint32_t MOD(int i) {
        constexpr uint32_t m[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return m[i];
    }
    G16_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t m[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return m[i];
    }
    G16_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t m[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return m[i];
    }
};

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    G16_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.l[i] = 0;
        return r;
    }
    G16_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.l[i] = P::ONE(i);
        return r;
    }
    G16_HD bool is_zero() const {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) v |= l[i];
        return v == 0;
    }
    G16_HD bool operator==(const Fp &o) const {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) v |= l[i] ^ o.l[i];
        return v == 0;
    }
    G16_HD bool operator!=(const Fp &o) const { return !(*this == o); }

    // r = x - p if x >= p else x   (x < 2p)
    G16_HD static void final_sub(uint32_t *x) {
        uint32_t s[N];
        s[0] = sub_cc(x[0], P::MOD(0));
#pragma unroll
        for (int i = 1; i < N; ++i) s[i] = subc_cc(x[i], P::MOD(i));
        uint32_t borrow = subc(0u, 0u);  // 0xffffffff when x < p
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = borrow ? x[i] : s[i];
    }

    G16_HD static Fp add(const Fp &a, const Fp &b) {
        Fp r;
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(a.l[i], b.l[i]);
        r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);  // p has spare top bits: no carry out
        final_sub(r.l);
        return r;
    }
    G16_HD static Fp dbl(const Fp &a) { return add(a, a); }

    G16_HD static Fp sub(const Fp &a, const Fp &b) {
        Fp r;
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) r.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = subc(0u, 0u);  // all ones when a < b
        // add back p & borrow
        r.l[0] = add_cc(r.l[0], P::MOD(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(r.l[i], P::MOD(i) & borrow);
        r.l[N - 1] = addc(r.l[N - 1], P::MOD(N - 1) & borrow);
        return r;
    }
    G16_HD static Fp neg(const Fp &a) { return sub(zero(), a); }

    // ---- Montgomery multiplication ------------------------------------------------------
    // One CIOS round on the split accumulator: value = E + O * 2^32.
    // E, O: N words each.  On exit E[0] == 0 (mod 2^32) and the caller swaps roles.
    template <bool FIRST>
    G16_HD static void round(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi) {
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                E[j] = mul_lo(a[j], bi);
                E[j + 1] = mul_hi(a[j], bi);
                O[j] = mul_lo(a[j + 1], bi);
                O[j + 1] = mul_hi(a[j + 1], bi);
            }
        } else {
            // previous round left O[0] == 0; O[1] sits on column 0, O[2..] become the new O
            E[0] = add_cc(E[0], O[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2]);
                O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 3]);
            }
            O[N - 2] = madc_lo_cc(a[N - 1], bi, 0u);
            O[N - 1] = madc_hi(a[N - 1], bi, 0u);
            E[0] = mad_lo_cc(a[0], bi, E[0]);
            E[1] = madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                E[j] = madc_lo_cc(a[j], bi, E[j]);
                E[j + 1] = madc_hi_cc(a[j], bi, E[j + 1]);
            }
            O[N - 1] = addc(O[N - 1], 0u);
        }
        uint32_t m = E[0] * P::NINV;
        O[0] = mad_lo_cc(P::MOD(1), m, O[0]);
        O[1] = madc_hi_cc(P::MOD(1), m, O[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            O[j] = madc_lo_cc(P::MOD(j + 1), m, O[j]);
            O[j + 1] = madc_hi_cc(P::MOD(j + 1), m, O[j + 1]);
        }
        E[0] = mad_lo_cc(P::MOD(0), m, E[0]);
        E[1] = madc_hi_cc(P::MOD(0), m, E[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            E[j] = madc_lo_cc(P::MOD(j), m, E[j]);
            E[j + 1] = madc_hi_cc(P::MOD(j), m, E[j + 1]);
        }
        O[N - 1] = addc(O[N - 1], 0u);
    }

    G16_MUL_HD static Fp mul(const Fp &a, const Fp &b) {
        uint32_t ev[N], od[N];
        round<true>(ev, od, a.l, b.l[0]);
        round<false>(od, ev, a.l, b.l[1]);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            round<false>(ev, od, a.l, b.l[i]);
            round<false>(od, ev, a.l, b.l[i + 1]);
        }
        // last round had E = od, O = ev:  result = ev + (od >> 32)
        Fp r;
        r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(ev[i], od[i + 1]);
        r.l[N - 1] = addc(ev[N - 1], 0u);
        final_sub(r.l);
        return r;
    }
    // ---- K products, one reduction:  (sum_k a_k x_k) R^-1 mod p --------------------------------------------------
    // The same CIOS rounds with K product terms per round: (K + 1) N^2 + N wide MADs instead of the K (2 N^2 + N) of
    // K multiplications, and the sum of the products never exists as a 2N-word number -- the accumulator stays N + 1
    // words (the running value is below (K + 1) p 2^32 < 2^(32 N + 32) for K <= 4: p has three spare top bits).  With
    // all a_k < p the result is below p (K p / R + 1) < 2p for K <= 4, so one conditional subtraction restores the
    // invariant.  Uses: Fq2 products c0 = a0 b0 + a1 (-b1), c1 = a0 b1 + a1 b0 (fq2.cuh), and the y coordinate of every
    // group addition, y3 = r (q - x3) - y1 ppp (ec.cuh): one reduction saved each time.
    // term k of one round: value += a * bi on the split accumulator (E, O), continuing a value already in place
    G16_HD static void round_term(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi) {
        O[0] = mad_lo_cc(a[1], bi, O[0]);
        O[1] = madc_hi_cc(a[1], bi, O[1]);
#pragma unroll
        for (int j = 2; j < N - 2; j += 2) {
            O[j] = madc_lo_cc(a[j + 1], bi, O[j]);
            O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 1]);
        }
        O[N - 2] = madc_lo_cc(a[N - 1], bi, O[N - 2]);
        O[N - 1] = madc_hi(a[N - 1], bi, O[N - 1]);   // no carry leaves O[N-1] (bound above)
        E[0] = mad_lo_cc(a[0], bi, E[0]);
        E[1] = madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            E[j] = madc_lo_cc(a[j], bi, E[j]);
            E[j + 1] = madc_hi_cc(a[j], bi, E[j + 1]);
        }
        O[N - 1] = addc(O[N - 1], 0u);
    }
    // first term of a round (merges the shift of the previous round's O, as `round` does) ...
    template <bool FIRST>
    G16_HD static void round_open(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi) {
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                E[j] = mul_lo(a[j], bi);
                E[j + 1] = mul_hi(a[j], bi);
                O[j] = mul_lo(a[j + 1], bi);
                O[j + 1] = mul_hi(a[j + 1], bi);
            }
        } else {
            E[0] = add_cc(E[0], O[1]);
#pragma unroll
            for (int j = 0; j < N - 2; j += 2) {
                O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2]);
                O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 3]);
            }
            O[N - 2] = madc_lo_cc(a[N - 1], bi, 0u);
            O[N - 1] = madc_hi(a[N - 1], bi, 0u);
            E[0] = mad_lo_cc(a[0], bi, E[0]);
            E[1] = madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
            for (int j = 2; j < N; j += 2) {
                E[j] = madc_lo_cc(a[j], bi, E[j]);
                E[j + 1] = madc_hi_cc(a[j], bi, E[j + 1]);
            }
            O[N - 1] = addc(O[N - 1], 0u);
        }
    }
    // ... and the reduction step that closes it: on exit E[0] == 0 (mod 2^32) and the caller swaps roles
    G16_HD static void round_close(uint32_t *E, uint32_t *O) {
        uint32_t m = E[0] * P::NINV;
        O[0] = mad_lo_cc(P::MOD(1), m, O[0]);
        O[1] = madc_hi_cc(P::MOD(1), m, O[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            O[j] = madc_lo_cc(P::MOD(j + 1), m, O[j]);
            O[j + 1] = madc_hi_cc(P::MOD(j + 1), m, O[j + 1]);
        }
        E[0] = mad_lo_cc(P::MOD(0), m, E[0]);
        E[1] = madc_hi_cc(P::MOD(0), m, E[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            E[j] = madc_lo_cc(P::MOD(j), m, E[j]);
            E[j + 1] = madc_hi_cc(P::MOD(j), m, E[j + 1]);
        }
        O[N - 1] = addc(O[N - 1], 0u);
    }
    template <bool FIRST, int K>
    G16_HD static void round_sum(uint32_t *E, uint32_t *O, const Fp *const *a, const Fp *const *x, int i) {
        round_open<FIRST>(E, O, a[0]->l, x[0]->l[i]);
#pragma unroll
        for (int k = 1; k < K; ++k) round_term(E, O, a[k]->l, x[k]->l[i]);
        round_close(E, O);
    }
    template <int K>
    G16_HD static Fp mul_sum(const Fp *const *a, const Fp *const *x) {
        static_assert(K >= 1 && K <= 4, "the one-subtraction bound holds up to four products");
        uint32_t ev[N], od[N];
        round_sum<true, K>(ev, od, a, x, 0);
        round_sum<false, K>(od, ev, a, x, 1);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            round_sum<false, K>(ev, od, a, x, i);
            round_sum<false, K>(od, ev, a, x, i + 1);
        }
        Fp r;
        r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(ev[i], od[i + 1]);
        r.l[N - 1] = addc(ev[N - 1], 0u);
        final_sub(r.l);
        return r;
    }
    // a x + a2 y
    G16_MUL_HD static Fp mul_dual(const Fp &a, const Fp &x, const Fp &a2, const Fp &y) {
        const Fp *as[2] = {&a, &a2}, *xs[2] = {&x, &y};
        return mul_sum<2>(as, xs);
    }
    // a x + a2 y + a3 z + a4 w
    G16_MUL_HD static Fp mul_quad(const Fp &a, const Fp &x, const Fp &a2, const Fp &y, const Fp &a3, const Fp &z,
                                  const Fp &a4, const Fp &w) {
        const Fp *as[4] = {&a, &a2, &a3, &a4}, *xs[4] = {&x, &y, &z, &w};
        return mul_sum<4>(as, xs);
    }
    // a x - b y  (the y coordinate of the group additions); same interface on Fq2
    // (-DG16_MUL_DIFF=0 builds the two-multiplication form for A/B runs, tools/lab_build.py)
#ifndef G16_MUL_DIFF
#define G16_MUL_DIFF 1
#endif
    G16_HD static Fp mul_diff(const Fp &a, const Fp &x, const Fp &b, const Fp &y) {
        return G16_MUL_DIFF ? mul_dual(a, x, neg(b), y) : sub(mul(a, x), mul(b, y));
    }
    // A dedicated squaring (fewer IMAD.WIDE, longer dependent structure) was measured 5 % slower inside the bucket
    // kernel (profiles/README.md run 7); it lives in experiments/fp_wide.cuh, not in the library.
    G16_HD static Fp sqr(const Fp &a) { return mul(a, a); }

    // Montgomery form <-> canonical integer limbs
    G16_HD static Fp from_mont(const Fp &a) {
        Fp o = zero();
        o.l[0] = 1;
        return mul(a, o);
    }
    G16_HD static Fp to_mont(const Fp &a) {
        Fp r2;
#pragma unroll
        for (int i = 0; i < N; ++i) r2.l[i] = P::R2(i);
        return mul(a, r2);
    }

    // a >= b on raw limbs (wire format: canonical-range and "larger root" tests)
    G16_HD static bool geq_raw(const uint32_t *a, const uint32_t *b) {
        sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) subc_cc(a[i], b[i]);
        return subc(0u, 0u) == 0u;
    }

    // ---- inversion: Bernstein-Yang "safegcd" division steps, 30 per batch ----------------------------------------------
    // State (eta, f, g) with f odd, f = p, g = a.  A division step is
    //     eta < 0 and g odd:  (eta, f, g) <- (-eta - 1, g, (g - f) / 2)        otherwise:  (eta - 1, f, (g + (g odd) f) / 2)
    // and depends only on the low bits, so 30 steps are run on the low words and collected in a 2 x 2 matrix t with
    // entries below 2^30 in magnitude; then (f, g) <- t (f, g) / 2^30 exactly and (d, e) <- t (d, e) / 2^30 mod p, where
    // d a = f and e a = g (mod p) throughout.  After at most (49 bits + 57) / 17 steps g = 0 and f = +-1, so d = +-a^-1.
    // About 600 cheap instructions per batch and no data-dependent branch inside a batch -- against ~760 multi-word
    // steps with divergent inner loops of the binary extended Euclid it replaces (the to-affine tail of every MSM, the
    // shared inversions of the fixed-base and precompute kernels).  Numbers are held in signed 30-bit limbs (value =
    // sum v[i] 2^(30 i), lower limbs in [0, 2^30), the top limb carries the sign) so that products accumulate in 64 bits.
    // Input/output in Montgomery form; inv(0) = 0 (never used: callers test for infinity first).
    static constexpr int L30 = (32 * N + 29) / 30;                             // Fq: 13 limbs, Fr: 9
    static constexpr int INV_BATCHES = ((49 * 32 * N + 57) / 17 + 29) / 30;    // Fq: 37, Fr: 25
    G16_HD static void to_limbs30(const uint32_t *x, int32_t *v) {
#pragma unroll
        for (int i = 0; i < L30; ++i) {
            const int w = (30 * i) / 32, sh = (30 * i) % 32;
            uint32_t lo = w < N ? x[w] >> sh : 0u;
            if (sh > 2 && w + 1 < N) lo |= x[w + 1] << (32 - sh);
            v[i] = (int32_t)(lo & 0x3fffffffu);
        }
    }
    G16_HD static void from_limbs30(const int32_t *v, uint32_t *x) {   // limbs in [0, 2^30)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int i = (32 * j) / 30, o = (32 * j) % 30;
            uint64_t t = (uint64_t)(uint32_t)v[i] >> o;
            if (i + 1 < L30) t |= (uint64_t)(uint32_t)v[i + 1] << (30 - o);
            if (i + 2 < L30 && 60 - o < 32) t |= (uint64_t)(uint32_t)v[i + 2] << (60 - o);
            x[j] = (uint32_t)t;
        }
    }
    G16_HD static Fp inv(const Fp &a) {
        if (a.is_zero()) return zero();
        constexpr int32_t M30 = 0x3fffffff;
        Fp x = from_mont(a);
        uint32_t mod[N];
#pragma unroll
        for (int i = 0; i < N; ++i) mod[i] = P::MOD(i);
        int32_t m[L30], f[L30], g[L30], d[L30], e[L30];
        to_limbs30(mod, m);
        to_limbs30(x.l, g);
#pragma unroll
        for (int i = 0; i < L30; ++i) { f[i] = m[i]; d[i] = 0; e[i] = 0; }
        e[0] = 1;
        // p^-1 mod 2^30 by Newton's iteration (p p = 1 mod 8, every step doubles the number of correct bits)
        const uint32_t p0 = (uint32_t)m[0] | ((uint32_t)m[1] << 30);
        uint32_t pinv = p0;
        pinv *= 2u - p0 * pinv; pinv *= 2u - p0 * pinv; pinv *= 2u - p0 * pinv; pinv *= 2u - p0 * pinv;
        int32_t eta = -1;
#pragma unroll 1
        for (int it = 0; it < INV_BATCHES; ++it) {
            // 30 division steps on the low words; u, v, q, r are signed entries kept mod 2^32
            uint32_t u = 1, v = 0, q = 0, r = 1;
            uint32_t fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
#pragma unroll 6
            for (int i = 0; i < 30; ++i) {
                uint32_t c1 = (uint32_t)(eta >> 31), c2 = 0u - (gl & 1u);
                uint32_t xf = (fl ^ c1) - c1, yu = (u ^ c1) - c1, zv = (v ^ c1) - c1;   // -f, -u, -v when eta < 0
                gl += xf & c2; q += yu & c2; r += zv & c2;
                c1 &= c2;                                                              // eta < 0 and g odd: swap
                eta = (int32_t)(((uint32_t)eta ^ c1) - (c1 + 1u));
                fl += gl & c1; u += q & c1; v += r & c1;
                gl >>= 1; u <<= 1; v <<= 1;
            }
            const int32_t tu = (int32_t)u, tv = (int32_t)v, tq = (int32_t)q, tr = (int32_t)r;
            // (d, e) <- t (d, e) / 2^30 mod p: add the multiple (md, me) of p that clears the low 30 bits
            {
                const int32_t sd = d[L30 - 1] >> 31, se = e[L30 - 1] >> 31;
                int32_t md = (tu & sd) + (tv & se), me = (tq & sd) + (tr & se);
                int64_t cd = (int64_t)tu * d[0] + (int64_t)tv * e[0], ce = (int64_t)tq * d[0] + (int64_t)tr * e[0];
                md -= (int32_t)((pinv * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
                me -= (int32_t)((pinv * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
                cd += (int64_t)m[0] * md; ce += (int64_t)m[0] * me;
                cd >>= 30; ce >>= 30;
#pragma unroll
                for (int i = 1; i < L30; ++i) {
                    cd += (int64_t)tu * d[i] + (int64_t)tv * e[i] + (int64_t)m[i] * md;
                    ce += (int64_t)tq * d[i] + (int64_t)tr * e[i] + (int64_t)m[i] * me;
                    d[i - 1] = (int32_t)cd & M30; cd >>= 30;
                    e[i - 1] = (int32_t)ce & M30; ce >>= 30;
                }
                d[L30 - 1] = (int32_t)cd; e[L30 - 1] = (int32_t)ce;
            }
            // (f, g) <- t (f, g) / 2^30 (exact)
            {
                int64_t cf = (int64_t)tu * f[0] + (int64_t)tv * g[0], cg = (int64_t)tq * f[0] + (int64_t)tr * g[0];
                cf >>= 30; cg >>= 30;
                int32_t nz = 0;
#pragma unroll
                for (int i = 1; i < L30; ++i) {
                    cf += (int64_t)tu * f[i] + (int64_t)tv * g[i];
                    cg += (int64_t)tq * f[i] + (int64_t)tr * g[i];
                    f[i - 1] = (int32_t)cf & M30; cf >>= 30;
                    g[i - 1] = (int32_t)cg & M30; cg >>= 30;
                    nz |= g[i - 1];
                }
                f[L30 - 1] = (int32_t)cf; g[L30 - 1] = (int32_t)cg;
                if ((nz | g[L30 - 1]) == 0) break;
            }
        }
        // f = +-1 and d = +-a^-1 in (-2p, p): add p if negative, negate if f < 0, add p again if still negative
        {
            const int32_t neg = f[L30 - 1] >> 31;
            int32_t add = d[L30 - 1] >> 31;
#pragma unroll
            for (int i = 0; i < L30; ++i) d[i] = ((d[i] + (m[i] & add)) ^ neg) - neg;
#pragma unroll
            for (int i = 0; i < L30 - 1; ++i) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
            add = d[L30 - 1] >> 31;
#pragma unroll
            for (int i = 0; i < L30; ++i) d[i] += m[i] & add;
#pragma unroll
            for (int i = 0; i < L30 - 1; ++i) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
        }
        Fp r;
        from_limbs30(d, r.l);
        return to_mont(r);   // canonical a^-1 -> Montgomery form
    }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace g16
