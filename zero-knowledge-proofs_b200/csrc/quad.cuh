// Quad-cooperative XYZZ group law for the latency-bound tails (device only).
//
// The upper levels of the bucket reduction tree have far fewer independent additions than the GPU has
// lanes, so their duration is (number of dependent additions) x (latency of one addition).  A full XYZZ
// addition is 14 field multiplications of which at most 4 depend on each other; here the four lanes of a
// quad hold the SAME operands and each lane computes a different product of every dependency level, the
// products being exchanged with warp shuffles.  One addition then costs 4 multiplication latencies instead
// of 14 (a doubling 4 instead of 9).  All 32 lanes of a warp must call these functions together (the
// shuffles are full-warp); lanes without work pass the point at infinity.
//
// Same formulas as xyzz_add / xyzz_dbl in ec.cuh (EFD add-2008-s, dbl-2008-s-1); exceptional cases are
// resolved after the last shuffle, falling back to the serial routines for P + P.
#pragma once
#include "ec.cuh"

#if !defined(G16_EMU) && defined(__CUDACC__)
namespace g16 {

template <class F>
__device__ __forceinline__ F quad_bcast(const F &v, int src) {
    F r;
    const uint32_t *s = limbs(v);
    uint32_t *d = limbs(r);
#pragma unroll
    for (int k = 0; k < F::N; ++k) d[k] = __shfl_sync(0xffffffffu, s[k], src, 4);
    return r;
}
template <class F>
__device__ __forceinline__ F quad_xor1(const F &v) {
    F r;
    const uint32_t *s = limbs(v);
    uint32_t *d = limbs(r);
#pragma unroll
    for (int k = 0; k < F::N; ++k) d[k] = __shfl_xor_sync(0xffffffffu, s[k], 1, 4);
    return r;
}
template <class F>
__device__ __forceinline__ F fsel(bool c, const F &a, const F &b) {   // c ? a : b
    F r;
    const uint32_t *x = limbs(a), *y = limbs(b);
    uint32_t *d = limbs(r);
#pragma unroll
    for (int k = 0; k < F::N; ++k) d[k] = c ? x[k] : y[k];
    return r;
}
template <class F>
__device__ __forceinline__ F fsel4(int q, const F &a0, const F &a1, const F &a2, const F &a3) {
    return fsel(q < 2, fsel(q == 0, a0, a1), fsel(q == 2, a2, a3));
}

// acc += b, q = lane & 3; every lane of the quad holds the same acc and b and leaves with the same result
template <class F>
__device__ __noinline__ void xyzz_add_quad(XYZZ<F> &acc, const XYZZ<F> &b, int q) {
    const bool binf = b.is_inf(), ainf = acc.is_inf();
    // level 1:  q0 u1 = X1 ZZ2,  q1 u2 = X2 ZZ1,  q2 s1 = Y1 ZZZ2,  q3 s2 = Y2 ZZZ1
    F m1 = F::mul(fsel4(q, acc.x, b.x, acc.y, b.y), fsel4(q, b.zz, acc.zz, b.zzz, acc.zzz));
    F o1 = quad_xor1(m1);
    F first = fsel((q & 1) != 0, o1, m1);    // u1 (q < 2) / s1 (q >= 2)
    F second = fsel((q & 1) != 0, m1, o1);   // u2 / s2
    F d = F::sub(second, first);             // p = u2 - u1 (q < 2) / r = s2 - s1 (q >= 2)
    // level 2:  q0, q1 pp = p^2,  q2 rr = r^2,  q3 zz12 = ZZ1 ZZ2
    F m2 = F::mul(fsel(q == 3, acc.zz, d), fsel(q == 3, b.zz, d));
    F pp = quad_bcast(m2, 0);
    // level 3:  q0 ppp = p pp,  q1 qq = u1 pp,  q2 zzz12 = ZZZ1 ZZZ2,  q3 zz3 = zz12 pp
    F m3 = F::mul(fsel4(q, d, first, acc.zzz, m2), fsel(q == 2, b.zzz, pp));
    F ppp = quad_bcast(m3, 0), qq = quad_bcast(m3, 1), rr = quad_bcast(m2, 2), zz3 = quad_bcast(m3, 3);
    F x3 = F::sub(F::sub(rr, ppp), F::dbl(qq));
    F r = quad_bcast(d, 2), s1 = quad_bcast(first, 2), p = quad_bcast(d, 0);
    // level 4:  q0 t1 = r (qq - x3),  q1 t2 = s1 ppp,  q2 (q3) zzz3 = zzz12 ppp
    F m4 = F::mul(fsel4(q, r, s1, m3, m3), fsel(q == 0, F::sub(qq, x3), ppp));
    F t1 = quad_bcast(m4, 0), t2 = quad_bcast(m4, 1), zzz3 = quad_bcast(m4, 2);
    // exceptional cases (uniform inside the quad; no shuffles below this line)
    if (binf) return;
    if (ainf) { acc = b; return; }
    if (p.is_zero()) {
        if (r.is_zero()) xyzz_dbl_call(acc);
        else acc = XYZZ<F>::inf();
        return;
    }
    acc.x = x3;
    acc.y = F::sub(t1, t2);
    acc.zz = zz3;
    acc.zzz = zzz3;
}

// acc = 2 acc
template <class F>
__device__ __noinline__ void xyzz_dbl_quad(XYZZ<F> &acc, int q) {
    const bool ainf = acc.is_inf();
    F u = F::dbl(acc.y);
    // level 1:  q1 xx = X^2,  others v = U^2
    F m1 = F::mul(fsel(q == 1, acc.x, u), fsel(q == 1, acc.x, u));
    F v = quad_bcast(m1, 0), xx = quad_bcast(m1, 1);
    F m = F::add(F::dbl(xx), xx);
    // level 2:  q0 w = U V,  q1 s = X V,  q2 (q3) zz3 = V ZZ
    F m2 = F::mul(fsel4(q, u, acc.x, acc.zz, acc.zz), v);
    F w = quad_bcast(m2, 0), s = quad_bcast(m2, 1), zz3 = quad_bcast(m2, 2);
    // level 3:  q0 mm = M^2,  q1 wy = W Y,  q2 (q3) zzz3 = W ZZZ
    F m3 = F::mul(fsel4(q, m, w, w, w), fsel4(q, m, acc.y, acc.zzz, acc.zzz));
    F mm = quad_bcast(m3, 0), wy = quad_bcast(m3, 1), zzz3 = quad_bcast(m3, 2);
    F x3 = F::sub(mm, F::dbl(s));
    // level 4:  every lane t = M (S - X3)
    F t = F::mul(m, F::sub(s, x3));
    if (ainf) return;
    acc.x = x3;
    acc.y = F::sub(t, wy);
    acc.zz = zz3;
    acc.zzz = zzz3;
}

}  // namespace g16
#endif
