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
    // Karatsuba: 3 Fq multiplications
    G16_HD static Fq2 mul_karatsuba(const Fq2 &a, const Fq2 &b) {
        Fq v0 = Fq::mul(a.c0, b.c0);
        Fq v1 = Fq::mul(a.c1, b.c1);
        Fq s = Fq::mul(Fq::add(a.c0, a.c1), Fq::add(b.c0, b.c1));
        return Fq2{Fq::sub(v0, v1), Fq::sub(Fq::sub(s, v0), v1)};
    }
    // schoolbook with two products per reduction (Fp::mul_dual): 2 x (2 products + 1 reduction) = 888 wide MADs against
    // Karatsuba's 900, one negation instead of five Fq additions / subtractions, no v0 / v1 / s temporaries
    G16_HD static Fq2 mul_dual(const Fq2 &a, const Fq2 &b) {
        Fq nb1 = Fq::neg(b.c1);
        return Fq2{Fq::mul_dual(a.c0, b.c0, a.c1, nb1), Fq::mul_dual(a.c0, b.c1, a.c1, b.c0)};
    }
    // Shipped: the two-product form (G2 accumulate at 2^20: 14.4 ms against 16.9 ms with Karatsuba on one box,
    // profiles/r02_run20_lab_g2_pair_and_dual.txt; 92 bytes of spills instead of 328).  -DG16_FQ2_DUAL=0 builds the
    // Karatsuba form for A/B runs (tools/lab_build.py).
    // In the cold translation units (G16_COLD: the Fq multiplication is an out-of-line call) the multi-product forms
    // lose -- their four / eight operands travel through local memory: G2 fixed-base 39 -> 29 M points/s, G2 precompute
    // 372 -> 527 ms with the four-product y3 (profiles/r02_run22_*) -- so those units keep Karatsuba unless they say
    // otherwise (k_pre_g2.cu: two-product multiplication yes, four-product y3 no).
#ifndef G16_FQ2_DUAL
#ifdef G16_COLD
#define G16_FQ2_DUAL 0
#else
#define G16_FQ2_DUAL 1
#endif
#endif
#ifndef G16_FQ2_QUAD   // y3 of the group additions as two four-product multiplications (mul_diff below)
#ifdef G16_COLD
#define G16_FQ2_QUAD 0
#else
#define G16_FQ2_QUAD G16_FQ2_DUAL
#endif
#endif
    G16_FQ2_MUL_HD static Fq2 mul(const Fq2 &a, const Fq2 &b) { return G16_FQ2_DUAL ? mul_dual(a, b) : mul_karatsuba(a, b); }
    // a x - b y: four products and one reduction per component
    //   c0 = a0 x0 - a1 x1 - b0 y0 + b1 y1,   c1 = a0 x1 + a1 x0 - b0 y1 - b1 y0
    G16_HD static Fq2 mul_diff(const Fq2 &a, const Fq2 &x, const Fq2 &b, const Fq2 &y) {
#if G16_FQ2_QUAD && G16_MUL_DIFF
        Fq nx1 = Fq::neg(x.c1), nb0 = Fq::neg(b.c0), nb1 = Fq::neg(b.c1);
        return Fq2{Fq::mul_quad(a.c0, x.c0, a.c1, nx1, nb0, y.c0, b.c1, y.c1),
                   Fq::mul_quad(a.c0, x.c1, a.c1, x.c0, nb0, y.c1, nb1, y.c0)};
#else
        return sub(mul(a, x), mul(b, y));
#endif
    }
    // complex squaring: 2 Fq multiplications
    G16_FQ2_MUL_HD static Fq2 sqr(const Fq2 &a) {
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
