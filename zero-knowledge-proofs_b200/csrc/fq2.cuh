// Fq2 = Fq[u]/(u^2+1) for G2.  Replaces ark-ff `QuadExtField` under `G2Projective::msm`
// (/root/reference/crates/groth16-core/src/lib.rs:296) and `g2_gen * fr`
// (crates/groth16-setup/src/lib.rs:168-171,205).  Same interface as Fp so that the curve
// templates in ec.cuh work for both groups.
#pragma once
#include "fp.cuh"

namespace g16 {

struct Fq2 {
    static constexpr int N = 24;  // u32 limbs: c0[12] || c1[12]  (ark layout: c0 then c1)
    Fq c0, c1;

    G16_HD static Fq2 zero() { return Fq2{Fq::zero(), Fq::zero()}; }
    G16_HD static Fq2 one() { return Fq2{Fq::one(), Fq::zero()}; }
    G16_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    G16_HD bool operator==(const Fq2 &o) const { return c0 == o.c0 && c1 == o.c1; }
    G16_HD bool operator!=(const Fq2 &o) const { return !(*this == o); }
    G16_HD static Fq2 add(const Fq2 &a, const Fq2 &b) { return Fq2{Fq::add(a.c0, b.c0), Fq::add(a.c1, b.c1)}; }
    G16_HD static Fq2 sub(const Fq2 &a, const Fq2 &b) { return Fq2{Fq::sub(a.c0, b.c0), Fq::sub(a.c1, b.c1)}; }
    G16_HD static Fq2 dbl(const Fq2 &a) { return Fq2{Fq::dbl(a.c0), Fq::dbl(a.c1)}; }
    G16_HD static Fq2 neg(const Fq2 &a) { return Fq2{Fq::neg(a.c0), Fq::neg(a.c1)}; }
    // Karatsuba with lazy reduction: 3 wide products, 2 Montgomery reductions (720 instead of 864 wide MADs).
    //   c1 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1 = a0 b1 + a1 b0 < 2 q^2 < q R
    //   c0 = a0 b0 - a1 b1 (+ q R when negative)                  < q R
    G16_HD static Fq2 mul_lazy(const Fq2 &a, const Fq2 &b) {
        constexpr int N = Fq::N;
        uint32_t v0[2 * N], v1[2 * N], sw[2 * N], sa[N], sb[N];
        Fq::mul_wide(a.c0.l, b.c0.l, v0);
        Fq::mul_wide(a.c1.l, b.c1.l, v1);
        Fq::add_noreduce(a.c0.l, a.c1.l, sa);
        Fq::add_noreduce(b.c0.l, b.c1.l, sb);
        Fq::mul_wide(sa, sb, sw);
        Fq::sub_wide(sw, v0);
        Fq::sub_wide(sw, v1);
        uint32_t borrow = Fq::sub_wide(v0, v1);
        v0[N] = add_cc(v0[N], FqParams::MOD(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) v0[N + i] = addc_cc(v0[N + i], FqParams::MOD(i) & borrow);
        v0[2 * N - 1] = addc(v0[2 * N - 1], FqParams::MOD(N - 1) & borrow);
        return Fq2{Fq::redc_wide(v0), Fq::redc_wide(sw)};
    }
    // Karatsuba: 3 Fq multiplications
    G16_HD static Fq2 mul_karatsuba(const Fq2 &a, const Fq2 &b) {
        Fq v0 = Fq::mul(a.c0, b.c0);
        Fq v1 = Fq::mul(a.c1, b.c1);
        Fq s = Fq::mul(Fq::add(a.c0, a.c1), Fq::add(b.c0, b.c1));
        return Fq2{Fq::sub(v0, v1), Fq::sub(Fq::sub(s, v0), v1)};
    }
    G16_HD static Fq2 mul(const Fq2 &a, const Fq2 &b) {
#if defined(G16_FQ2_LAZY)
        return mul_lazy(a, b);
#else
        return mul_karatsuba(a, b);
#endif
    }
    // complex squaring: 2 Fq multiplications
    G16_HD static Fq2 sqr(const Fq2 &a) {
        Fq m = Fq::mul(a.c0, a.c1);
        Fq t = Fq::mul(Fq::add(a.c0, a.c1), Fq::sub(a.c0, a.c1));
        return Fq2{t, Fq::dbl(m)};
    }
    G16_HD static Fq2 inv(const Fq2 &a) {
        Fq n = Fq::inv(Fq::add(Fq::sqr(a.c0), Fq::sqr(a.c1)));
        return Fq2{Fq::mul(a.c0, n), Fq::neg(Fq::mul(a.c1, n))};
    }
};

// limb-wise load/store helpers shared by Fq and Fq2 (both are plain arrays of u32)
template <class F>
G16_HD const uint32_t *limbs(const F &f) { return reinterpret_cast<const uint32_t *>(&f); }
template <class F>
G16_HD uint32_t *limbs(F &f) { return reinterpret_cast<uint32_t *>(&f); }

}  // namespace g16
