// Pippenger bucket-method kernels (per-thread bodies; see rt.cuh for how they are launched).
//
// Replaces ark-ec 0.4.2 `VariableBaseMSM::msm` (msm_bigint_wnaf) as called from
// /root/reference/crates/groth16-core/src/lib.rs:282 (G1) and :296 (G2):
//   scalar `into_bigint()` + `make_digits`       -> DigitCount / DigitScatter  (signed windows)
//   in-order `buckets[|d|-1] +=/-= base`         -> counting sort by bucket + BucketAccumulate
//   sequential running-sum per window            -> ReduceLevel (parallel, log_L levels)
//   Horner fold of window sums + `into_affine`   -> WindowCombine
// The result is the same group element, returned as the canonical affine point.
#pragma once
#include "ec.cuh"
#include "kernel_api.cuh"

namespace g16 {

// Signed window digits of a canonical 256-bit scalar k (8 x u32): k = sum d_w 2^(c w),
// d_w in [-(2^(c-1) - 1), 2^(c-1)].  nwin * c >= 256 guarantees the final carry is zero.
struct DigitIter {
    const uint32_t *k;
    uint32_t c, carry, w;
    G16_HD DigitIter(const uint32_t *k_, uint32_t c_) : k(k_), c(c_), carry(0), w(0) {}
    G16_HD int32_t next() {
        uint32_t bit = w * c;
        uint32_t word = bit >> 5, sh = bit & 31;
        uint64_t lo = word < 8 ? k[word] : 0u;
        uint64_t hi = word + 1 < 8 ? k[word + 1] : 0u;
        uint32_t raw = (uint32_t)(((lo | (hi << 32)) >> sh) & ((1u << c) - 1u));
        raw += carry;
        ++w;
        if (raw > (1u << (c - 1))) { carry = 1; return (int32_t)raw - (int32_t)(1u << c); }
        carry = 0;
        return (int32_t)raw;
    }
};

// scalar -> canonical integer limbs.  ark keeps Fr in Montgomery form; `into_bigint()` is one
// Montgomery multiplication by 1.
G16_HD void load_scalar(const uint32_t *scalars, size_t i, bool mont, uint32_t out[8]) {
    Fr s;
#pragma unroll
    for (int j = 0; j < 8; ++j) s.l[j] = scalars[8 * i + j];
    if (mont) s = Fr::from_mont(s);
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = s.l[j];
}

struct DigitCount {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t i, const uint32_t *scalars, bool mont, MsmPlan plan, uint32_t *counts) {
        uint32_t k[8];
        load_scalar(scalars, i, mont, k);
        DigitIter it(k, plan.c);
        for (uint32_t w = 0; w < plan.nwin; ++w) {
            int32_t d = it.next();
            if (d != 0) {
                uint32_t b = (uint32_t)(d < 0 ? -d : d) - 1u;
                atomic_add_u32(&counts[w * plan.nb + b], 1u);
            }
        }
    }
};

// entries[cursor[bucket]++] = point index | sign << 31
struct DigitScatter {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t i, const uint32_t *scalars, bool mont, MsmPlan plan, uint32_t *cursor,
                           uint32_t *entries) {
        uint32_t k[8];
        load_scalar(scalars, i, mont, k);
        DigitIter it(k, plan.c);
        for (uint32_t w = 0; w < plan.nwin; ++w) {
            int32_t d = it.next();
            if (d != 0) {
                uint32_t b = (uint32_t)(d < 0 ? -d : d) - 1u;
                uint32_t pos = atomic_add_u32(&cursor[w * plan.nb + b], 1u);
                entries[pos] = (uint32_t)i | (d < 0 ? 0x80000000u : 0u);
            }
        }
    }
};

// packed affine point in HBM: x limbs then y limbs, (0,0) = infinity
template <class F>
G16_HD Affine<F> load_affine(const uint32_t *pts, size_t idx) {
    Affine<F> p;
    const uint32_t *src = pts + idx * (2 * F::N);
    uint32_t *dx = limbs(p.x), *dy = limbs(p.y);
#if G16_DEVICE_CODE
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int j = 0; j < F::N / 4; ++j) {
        uint4 v = __ldg(s4 + j);
        dx[4 * j] = v.x; dx[4 * j + 1] = v.y; dx[4 * j + 2] = v.z; dx[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < F::N / 4; ++j) {
        uint4 v = __ldg(s4 + F::N / 4 + j);
        dy[4 * j] = v.x; dy[4 * j + 1] = v.y; dy[4 * j + 2] = v.z; dy[4 * j + 3] = v.w;
    }
#else
    for (int j = 0; j < F::N; ++j) { dx[j] = src[j]; dy[j] = src[F::N + j]; }
#endif
    return p;
}

template <class F>
G16_HD void store_xyzz(uint32_t *dst, size_t idx, const XYZZ<F> &p) {
    uint32_t *d = dst + idx * (4 * F::N);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&p);
#if G16_DEVICE_CODE
    uint4 *d4 = reinterpret_cast<uint4 *>(d);
#pragma unroll
    for (int j = 0; j < F::N; ++j) d4[j] = make_uint4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
#else
    for (int j = 0; j < 4 * F::N; ++j) d[j] = s[j];
#endif
}
template <class F>
G16_HD XYZZ<F> load_xyzz(const uint32_t *src, size_t idx) {
    XYZZ<F> p;
    const uint32_t *s = src + idx * (4 * F::N);
    uint32_t *d = reinterpret_cast<uint32_t *>(&p);
#if G16_DEVICE_CODE
    const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
#pragma unroll
    for (int j = 0; j < F::N; ++j) {
        uint4 v = s4[j];
        d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
    }
#else
    for (int j = 0; j < 4 * F::N; ++j) d[j] = s[j];
#endif
    return p;
}

// One thread per work item.  A work item is (bucket, [begin, end)) -- a whole bucket, or a
// slice of an oversized bucket (see engine: buckets longer than the chunk limit are split so
// that skewed scalar distributions cannot serialise on one thread).
template <class F>
struct BucketAccumulate {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t, const uint32_t *pts, const uint32_t *entries, const uint32_t *offsets,
                           uint32_t *buckets) {
        uint32_t begin = offsets[t], end = offsets[t + 1];
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t e = begin; e < end; ++e) {
            uint32_t v = entries[e];
            Affine<F> p = load_affine<F>(pts, v & 0x7fffffffu);
            if (v >> 31) p.y = F::neg(p.y);
            xyzz_madd(acc, p.x, p.y);
        }
        store_xyzz<F>(buckets, t, acc);
    }
};

// One level of the parallel bucket reduction.  For every window, the level maps arrays
//   X[0..n_in) (to be weighted by index) and Y[0..n_in) (already weighted partial sums)
// to arrays of length n_out = ceil(n_in / L):
//   X'[g] = sum_j X[gL + j]
//   Y'[g] = sum_j Y[gL + j] + 2^shift * sum_j j * X[gL + j]
// With shift = log2(L) * level this telescopes to  sum_i i * X0[i] = Y_final[0]
// (derivation in DESIGN.md "Bucket reduction").
template <class F>
struct ReduceLevel {
    static constexpr int BLOCK = 64;
    G16_HD static void run(size_t t, const uint32_t *X, const uint32_t *Y, uint32_t n_in, uint32_t n_out, uint32_t L,
                           uint32_t shift, uint32_t *Xo, uint32_t *Yo) {
        uint32_t w = (uint32_t)(t / n_out), g = (uint32_t)(t % n_out);
        size_t base = (size_t)w * n_in;
        uint32_t lo = g * L;
        uint32_t hi = lo + L < n_in ? lo + L : n_in;
        XYZZ<F> running = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
        for (uint32_t i = hi; i-- > lo + 1;) {
            XYZZ<F> x = load_xyzz<F>(X, base + i);
            xyzz_add(running, x);
            xyzz_add(acc, running);
        }
        {
            XYZZ<F> x0 = load_xyzz<F>(X, base + lo);
            xyzz_add(running, x0);
        }
        for (uint32_t s = 0; s < shift; ++s) xyzz_dbl(acc);
        if (Y) {
            for (uint32_t i = lo; i < hi; ++i) {
                XYZZ<F> y = load_xyzz<F>(Y, base + i);
                xyzz_add(acc, y);
            }
        }
        store_xyzz<F>(Xo, (size_t)w * n_out + g, running);
        store_xyzz<F>(Yo, (size_t)w * n_out + g, acc);
    }
};

// Final fold: window sum S_w = X[w] + Y[w] (bucket b carries weight b + 1), then
// result = sum_w 2^(c w) S_w by Horner, optionally + `extra` partial sums, then to affine.
// out_xyzz (4*F::N words) receives the projective result; out_aff (2*F::N words + flag word).
template <class F>
struct WindowCombine {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *X, const uint32_t *Y, uint32_t nwin, uint32_t c,
                           uint32_t *out_xyzz, uint32_t *out_aff) {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t w = nwin; w-- > 0;) {
            for (uint32_t s = 0; s < c; ++s) xyzz_dbl(acc);
            XYZZ<F> x = load_xyzz<F>(X, w);
            xyzz_add(acc, x);
            if (Y) {
                XYZZ<F> y = load_xyzz<F>(Y, w);
                xyzz_add(acc, y);
            }
        }
        if (out_xyzz) store_xyzz<F>(out_xyzz, 0, acc);
        if (out_aff) {
            Affine<F> a = xyzz_to_affine(acc);
            const uint32_t *s = reinterpret_cast<const uint32_t *>(&a);
            for (int j = 0; j < 2 * F::N; ++j) out_aff[j] = s[j];
            out_aff[2 * F::N] = acc.is_inf() ? 1u : 0u;
        }
    }
};

// Sum of k projective partial results (multi-GPU combine), then to affine.
template <class F>
struct PartialCombine {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *partials, uint32_t k, uint32_t *out_xyzz, uint32_t *out_aff) {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t i = 0; i < k; ++i) {
            XYZZ<F> p = load_xyzz<F>(partials, i);
            xyzz_add(acc, p);
        }
        if (out_xyzz) store_xyzz<F>(out_xyzz, 0, acc);
        if (out_aff) {
            Affine<F> a = xyzz_to_affine(acc);
            const uint32_t *s = reinterpret_cast<const uint32_t *>(&a);
            for (int j = 0; j < 2 * F::N; ++j) out_aff[j] = s[j];
            out_aff[2 * F::N] = acc.is_inf() ? 1u : 0u;
        }
    }
};

// Host layout (ark in-memory: x, y Montgomery limbs + separate infinity byte) -> device layout
// ((0,0) encodes infinity).  One thread per point.
template <class F>
struct ImportBases {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t i, const uint32_t *xy, const uint8_t *inf, uint32_t *pts) {
        bool is_inf = inf && inf[i];
        for (int j = 0; j < 2 * F::N; ++j) pts[i * (2 * F::N) + j] = is_inf ? 0u : xy[i * (2 * F::N) + j];
    }
};

}  // namespace g16
