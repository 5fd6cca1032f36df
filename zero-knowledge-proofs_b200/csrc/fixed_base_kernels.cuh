// Fixed-base batch scalar multiplication for CRS generation.
//
// Replaces the `par_iter().map(|v| (g1_gen * fr).into_affine())` blocks of
// /root/reference/crates/groth16-setup/src/lib.rs:185-241 (and the six single muls at :166-171):
// ark does an independent double-and-add plus one inversion per element; here the base gets a
// window table (FB_WINDOWS windows of FB_BITS-bit signed digits, affine entries, L2 resident) and every scalar
// costs at most FB_WINDOWS mixed additions, followed by a to-affine step.
#pragma once
#include "msm_kernels.cuh"

namespace g16 {


// powers[j] = 2^(FB_BITS * j) * base  (XYZZ), one thread
template <class F>
struct FbPowers {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *base_xy, uint32_t *powers) {
        Affine<F> b = load_affine<F>(base_xy, 0);
        XYZZ<F> acc = XYZZ<F>::from_affine(b);
        for (uint32_t j = 0; j < FB_WINDOWS; ++j) {
            store_xyzz<F>(powers, j, acc);
            for (uint32_t s = 0; s < FB_BITS; ++s) xyzz_dbl(acc);
        }
    }
};

template <class F>
G16_HD void store_affine(uint32_t *dst, size_t idx, const Affine<F> &a) {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&a);
    uint32_t *d = dst + idx * (2 * F::N);
#if G16_DEVICE_CODE
    uint4 *d4 = reinterpret_cast<uint4 *>(d);
#pragma unroll
    for (int j = 0; j < F::N / 2; ++j) d4[j] = make_uint4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
#else
    for (int j = 0; j < 2 * F::N; ++j) d[j] = s[j];
#endif
}

// Shared inversion (Montgomery's trick) for up to G projective points of one thread: affine results, (0,0) = infinity.
template <class F, uint32_t G>
G16_HD void batch_to_affine(const XYZZ<F> *pt, uint32_t cnt, Affine<F> *out) {
    F prefix[G];
    F run = F::one();
    for (uint32_t e = 0; e < cnt; ++e) {
        prefix[e] = run;                                     // product of the zzz of the earlier finite points
        if (!pt[e].is_inf()) run = F::mul(run, pt[e].zzz);
    }
    F inv_all = field_inv_call(run);
    for (uint32_t e = cnt; e-- > 0;) {
        const XYZZ<F> &q = pt[e];
        Affine<F> r = Affine<F>::inf();
        if (!q.is_inf()) {
            F a = F::mul(inv_all, prefix[e]);                // 1 / zzz_e = Z^-3
            inv_all = F::mul(inv_all, q.zzz);
            F zi = F::mul(a, q.zz);                          // Z^-1
            r.x = F::mul(q.x, F::sqr(zi));
            r.y = F::mul(q.y, a);
        }
        out[e] = r;
    }
}

// table[j * FB_ENTRIES + (d - 1)] = d * powers[j] as affine points, d = 1 .. FB_ENTRIES.  One thread per
// FB_TABLE_GROUP consecutive d of one window: d0 * P by double-and-add, then P added FB_TABLE_GROUP - 1 times, one
// shared inversion for the group.
template <class F>
struct FbTable {
    static constexpr int BLOCK = 64;
    G16_HD static void run(size_t t, const uint32_t *powers, uint32_t *table) {
        constexpr uint32_t PER = FB_ENTRIES / FB_TABLE_GROUP;          // groups per window
        uint32_t j = (uint32_t)(t / PER), d0 = (uint32_t)(t % PER) * FB_TABLE_GROUP + 1u;
        XYZZ<F> p = load_xyzz<F>(powers, j);
        XYZZ<F> pt[FB_TABLE_GROUP];
        XYZZ<F> acc = XYZZ<F>::inf();
        for (int bit = (int)FB_BITS - 1; bit >= 0; --bit) {
            xyzz_dbl_call(acc);
            if ((d0 >> bit) & 1u) xyzz_add_call(acc, p);
        }
        pt[0] = acc;
        for (uint32_t e = 1; e < FB_TABLE_GROUP; ++e) {
            xyzz_add_call(acc, p);
            pt[e] = acc;
        }
        Affine<F> aff[FB_TABLE_GROUP];
        batch_to_affine<F, FB_TABLE_GROUP>(pt, FB_TABLE_GROUP, aff);
        for (uint32_t e = 0; e < FB_TABLE_GROUP; ++e)
            store_affine<F>(table, (size_t)j * FB_ENTRIES + (d0 - 1u) + e, aff[e]);
    }
};

// out[i] = scalar_i * base as a packed affine point ((0,0) = infinity).  One thread per FB_GROUP
// consecutive scalars: <= FB_WINDOWS mixed additions each (signed digits, see kernel_api.cuh), then ONE shared
// inversion for the group (Montgomery's trick) instead of ark's inversion per element.
constexpr uint32_t FB_GROUP = 16;
template <class F>
struct FbMul {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t, const uint32_t *scalars, bool mont, const uint32_t *table, size_t n, uint32_t *out) {
        XYZZ<F> pt[FB_GROUP];
        size_t first = t * FB_GROUP;
        uint32_t cnt = (uint32_t)(n - first < FB_GROUP ? n - first : FB_GROUP);
        for (uint32_t e = 0; e < cnt; ++e) {
            uint32_t k[8];
            load_scalar(scalars, first + e, mont, k);
            XYZZ<F> acc = XYZZ<F>::inf();
            uint32_t carry = 0;
            for (uint32_t j = 0; j < FB_WINDOWS; ++j) {
                uint32_t bit = j * FB_BITS;
                uint32_t raw = ((k[bit >> 5] >> (bit & 31u)) & ((1u << FB_BITS) - 1u)) + carry;
                // digits above 2^(FB_BITS-1) become raw - 2^FB_BITS with a carry; the top window never carries out
                // (scalars are below 2^255)
                carry = raw > FB_ENTRIES ? 1u : 0u;
                uint32_t mag = carry ? (1u << FB_BITS) - raw : raw;
                if (mag) {
                    Affine<F> p = load_affine<F>(table, (size_t)j * FB_ENTRIES + (mag - 1u));
                    if (carry) p.y = F::neg(p.y);
                    xyzz_madd_call(acc, p.x, p.y);
                }
            }
            pt[e] = acc;
        }
        Affine<F> aff[FB_GROUP];
        batch_to_affine<F, FB_GROUP>(pt, cnt, aff);
        for (uint32_t e = 0; e < cnt; ++e) store_affine<F>(out, first + e, aff[e]);
    }
};

// packed device points -> host layout helper: infinity flags from the (0,0) encoding
template <class F>
struct ExportFlags {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t i, const uint32_t *pts, uint8_t *inf) {
        uint32_t v = 0;
        for (int j = 0; j < 2 * F::N; ++j) v |= pts[i * (2 * F::N) + j];
        inf[i] = v == 0 ? 1 : 0;
    }
};

}  // namespace g16
