// Short-Weierstrass (a = 0) group law in XYZZ coordinates (X, Y, ZZ, ZZZ with x = X/ZZ,
// y = Y/ZZZ, ZZ^3 = ZZZ^2), templated on the coordinate field (Fq -> G1, Fq2 -> G2).
//
// Replaces ark-ec 0.4.2's Jacobian `Projective += &Affine`, `+= &Projective`,
// `double_in_place` and `into_affine` that run under
// /root/reference/crates/groth16-core/src/lib.rs:282,285,296,299 and
// crates/groth16-setup/src/lib.rs:166-241.  The *result* of every public entry point is
// converted to the canonical affine point, so the coordinate system is free; XYZZ has the
// cheapest mixed addition (8M + 2S, EFD "madd-2008-s").
//
// Exceptional cases are all handled (they are routine with this reference: CRS points are
// k*G for small k, so equal and opposite bases do meet inside one bucket):
//   affine operand at infinity  -> encoded as (0, 0), which is not on either curve -> no-op
//   accumulator at infinity     -> ZZ == 0
//   P + P                       -> doubling formulas,   P + (-P) -> infinity
#pragma once
#include "fq2.cuh"

namespace g16 {

template <class F>
struct Affine {
    F x, y;
    G16_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    G16_HD static Affine inf() { return Affine{F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
    F x, y, zz, zzz;
    G16_HD bool is_inf() const { return zz.is_zero(); }
    G16_HD static XYZZ inf() { return XYZZ{F::one(), F::one(), F::zero(), F::zero()}; }
    G16_HD static XYZZ from_affine(const Affine<F> &p) {
        if (p.is_inf()) return inf();
        return XYZZ{p.x, p.y, F::one(), F::one()};
    }
};

// acc = 2 * (x2, y2)   (EFD mdbl-2008-s-1, a = 0); (x2, y2) finite with y2 != 0
template <class F>
G16_HD XYZZ<F> xyzz_mdbl(const F &x2, const F &y2) {
    F u = F::dbl(y2);
    F v = F::sqr(u);
    F w = F::mul(u, v);
    F s = F::mul(x2, v);
    F xx = F::sqr(x2);
    F m = F::add(F::dbl(xx), xx);
    XYZZ<F> r;
    r.x = F::sub(F::sqr(m), F::dbl(s));
    r.y = F::mul_diff(m, F::sub(s, r.x), w, y2);
    r.zz = v;
    r.zzz = w;
    return r;
}

// acc += (x2, y2)  (EFD madd-2008-s)
template <class F>
G16_HD void xyzz_madd(XYZZ<F> &acc, const F &x2, const F &y2) {
    if (x2.is_zero() && y2.is_zero()) return;  // affine infinity
    if (acc.is_inf()) {
        acc.x = x2; acc.y = y2; acc.zz = F::one(); acc.zzz = F::one();
        return;
    }
    F u2 = F::mul(x2, acc.zz);
    F s2 = F::mul(y2, acc.zzz);
    F p = F::sub(u2, acc.x);
    F r = F::sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = xyzz_mdbl(x2, y2);
        else acc = XYZZ<F>::inf();
        return;
    }
    F pp = F::sqr(p);
    F ppp = F::mul(p, pp);
    F q = F::mul(acc.x, pp);
    F x3 = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q));
    acc.y = F::mul_diff(r, F::sub(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.zz = F::mul(acc.zz, pp);
    acc.zzz = F::mul(acc.zzz, ppp);
}

// acc = 2 * acc  (EFD dbl-2008-s-1, a = 0)
template <class F>
G16_HD void xyzz_dbl(XYZZ<F> &acc) {
    if (acc.is_inf()) return;
    F u = F::dbl(acc.y);
    F v = F::sqr(u);
    F w = F::mul(u, v);
    F s = F::mul(acc.x, v);
    F xx = F::sqr(acc.x);
    F m = F::add(F::dbl(xx), xx);
    F x3 = F::sub(F::sqr(m), F::dbl(s));
    acc.y = F::mul_diff(m, F::sub(s, x3), w, acc.y);
    acc.x = x3;
    acc.zz = F::mul(v, acc.zz);
    acc.zzz = F::mul(w, acc.zzz);
}

// acc += b  (EFD add-2008-s)
template <class F>
G16_HD void xyzz_add(XYZZ<F> &acc, const XYZZ<F> &b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc = b; return; }
    F u1 = F::mul(acc.x, b.zz);
    F u2 = F::mul(b.x, acc.zz);
    F s1 = F::mul(acc.y, b.zzz);
    F s2 = F::mul(b.y, acc.zzz);
    F p = F::sub(u2, u1);
    F r = F::sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) xyzz_dbl(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    F pp = F::sqr(p);
    F ppp = F::mul(p, pp);
    F q = F::mul(u1, pp);
    F x3 = F::sub(F::sub(F::sqr(r), ppp), F::dbl(q));
    acc.y = F::mul_diff(r, F::sub(q, x3), s1, ppp);
    acc.x = x3;
    acc.zz = F::mul(F::mul(acc.zz, b.zz), pp);
    acc.zzz = F::mul(F::mul(acc.zzz, b.zzz), ppp);
}

// Out-of-line full addition for the cold-ish kernels (reduction tree, combine): one copy of the 14
// multiplications per kernel instead of one per call site; the multiplications inside stay inlined.
template <class F>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline __attribute__((noinline))
#endif
void xyzz_add_call(XYZZ<F> &acc, const XYZZ<F> &b) { xyzz_add(acc, b); }
template <class F>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline __attribute__((noinline))
#endif
void xyzz_dbl_call(XYZZ<F> &acc) { xyzz_dbl(acc); }

template <class F>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline __attribute__((noinline))
#endif
void xyzz_madd_call(XYZZ<F> &acc, const F &x2, const F &y2) { xyzz_madd(acc, x2, y2); }
template <class F>
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline __attribute__((noinline))
#endif
F field_inv_call(const F &a) { return F::inv(a); }

// canonical affine point (the form ark's `into_affine` returns); infinity -> (0, 0)
template <class F>
G16_HD Affine<F> xyzz_to_affine(const XYZZ<F> &p) {
    if (p.is_inf()) return Affine<F>::inf();
    F a = F::inv(p.zzz);              // Z^-3
    F zi = F::mul(a, p.zz);           // Z^-1
    Affine<F> r;
    r.x = F::mul(p.x, F::sqr(zi));    // X / ZZ
    r.y = F::mul(p.y, a);             // Y / ZZZ
    return r;
}

}  // namespace g16
