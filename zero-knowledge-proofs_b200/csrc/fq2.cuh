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
#ifndef G16_FQ2_DUAL
#define G16_FQ2_DUAL 1
#endif
    G16_FQ2_MUL_HD static Fq2 mul(const Fq2 &a, const Fq2 &b) { return G16_FQ2_DUAL ? mul_dual(a, b) : mul_karatsuba(a, b); }
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
