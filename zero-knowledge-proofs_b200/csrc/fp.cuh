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

    // ---- inversion: binary extended Euclid on canonical integers ----------------------------------
    // ~2 * bits cheap multi-word steps instead of the ~1.5 * bits Montgomery multiplications of a
    // Fermat ladder: an order of magnitude shorter serial chain for the single-threaded to-affine tails.
    // Input/output in Montgomery form; inv(0) = 0 (never used: callers test for infinity first).
    G16_HD static bool geq_raw(const uint32_t *a, const uint32_t *b) {   // a >= b
        sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) subc_cc(a[i], b[i]);
        return subc(0u, 0u) == 0u;
    }
    G16_HD static void sub_raw(uint32_t *a, const uint32_t *b) {         // a -= b (a >= b)
        a[0] = sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) a[i] = subc_cc(a[i], b[i]);
        a[N - 1] = subc(a[N - 1], b[N - 1]);
    }
    G16_HD static void shr1_raw(uint32_t *a, uint32_t top) {             // a = (top:a) >> 1
#pragma unroll
        for (int i = 0; i < N - 1; ++i) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[N - 1] = (a[N - 1] >> 1) | (top << 31);
    }
    // x = x / 2 mod p  (x < p)
    G16_HD static void halve_mod(uint32_t *x) {
        uint32_t odd = 0u - (x[0] & 1u), carry;
        x[0] = add_cc(x[0], P::MOD(0) & odd);
#pragma unroll
        for (int i = 1; i < N; ++i) x[i] = addc_cc(x[i], P::MOD(i) & odd);
        carry = addc(0u, 0u);
        shr1_raw(x, carry);
    }
    G16_HD static bool is_one_raw(const uint32_t *a) {
        uint32_t v = a[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; ++i) v |= a[i];
        return v == 0;
    }
    G16_HD static Fp inv(const Fp &a) {
        if (a.is_zero()) return zero();
        Fp u = from_mont(a), v, x1 = zero(), x2 = zero();
        x1.l[0] = 1;
#pragma unroll
        for (int i = 0; i < N; ++i) v.l[i] = P::MOD(i);
        // invariants: x1 * a == u, x2 * a == v (mod p); gcd(u, v) = 1
        while (!is_one_raw(u.l) && !is_one_raw(v.l)) {
            while (!(u.l[0] & 1u)) { shr1_raw(u.l, 0u); halve_mod(x1.l); }
            while (!(v.l[0] & 1u)) { shr1_raw(v.l, 0u); halve_mod(x2.l); }
            if (geq_raw(u.l, v.l)) { sub_raw(u.l, v.l); x1 = sub(x1, x2); }
            else { sub_raw(v.l, u.l); x2 = sub(x2, x1); }
        }
        Fp r = is_one_raw(u.l) ? x1 : x2;   // canonical a^-1 (of the canonical a)
        // Montgomery form of the inverse: a^-1 * R = mont_mul(mont_mul(a^-1, R^2), ...) -> one to_mont
        return to_mont(r);
    }
};

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace g16
