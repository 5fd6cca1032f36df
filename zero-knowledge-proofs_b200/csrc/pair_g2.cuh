// G2 bucket accumulation on lane pairs (device only).  NOT the shipped path: built only with -DG16_G2_ACC_THREAD=0
// (tools/lab_build.py g2_pair).  Bit-exact on B200 (all G2 parity tests), measured 16.9 ms at 168 registers / 15.8 ms at 255
// for the 2^20 accumulation against 14.1 ms for the one-thread kernel once that uses the same two-product multiplication
// (profiles/r02_run20_lab_g2_pair_and_dual.txt): kept because its arithmetic is the starting point for a
// lane-pair form of the G2 bucket REDUCTION.
//
// The one-thread form of the G2 mixed addition (ec.cuh with F = Fq2) needs more than the 255 registers a thread can
// have: it spills, runs two warps per scheduler and reaches 74 % of the Fq-multiply peak where the G1 kernel reaches
// 90 %.  Here two neighbouring lanes share one work item: lane h (= lane & 1) holds component c_h of every Fq2 value
// of the addition -- half the accumulator, half the point, half of every temporary.  Fq2 addition, subtraction,
// doubling and negation are component-wise, hence lane-local.  A multiplication needs the other component of both
// operands (one exchange of 12 words with the partner lane per operand, __shfl_xor) and is then ONE two-product
// Montgomery multiplication per lane (Fp::mul_dual, fp.cuh):
//     lane 0:  c0 = a0 b0 + a1 (-b1)        lane 1:  c1 = a1 b0 + a0 b1
// i.e. 2 x 444 wide MADs per Fq2 multiplication against Karatsuba's 900 in one thread, with no v0 / v1 / s
// temporaries.  A squaring is one plain multiplication per lane: (a0 + a1)(a0 - a1) on lane 0, 2 a0 a1 on lane 1.
//
// All 32 lanes of a warp walk their items in lockstep (the exchanges are full-warp shuffles): the loop runs to the
// longest item of the warp -- items are ordered by length, 16 items per warp differ by at most a few entries -- and a
// lane whose item has ended adds the point at infinity.  Exceptional cases are resolved by selects after the last
// exchange; P + P (equal bases meeting in one bucket: routine with this reference's CRS) runs the doubling for the
// whole warp when any pair needs it.
//
// Same formulas as xyzz_madd / xyzz_mdbl in ec.cuh (EFD madd-2008-s, mdbl-2008-s-1), replacing ark-ec's
// `Projective += &Affine` under `G2Projective::msm` (/root/reference/crates/groth16-core/src/lib.rs:296).
#pragma once
#include "msm_kernels.cuh"

#if !defined(G16_EMU) && defined(__CUDACC__)
namespace g16 {

#ifndef G16_G2_ACC_THREAD
#define G16_G2_ACC_THREAD 1   // 0: G2 accumulation on lane pairs (A/B builds, tools/lab_build.py); 1 (shipped): one thread per item
#endif
#ifndef G16_PAIR_BLOCK
#define G16_PAIR_BLOCK 64
#endif
#ifndef G16_PAIR_MIN_BLOCKS
#define G16_PAIR_MIN_BLOCKS 5
#endif

__device__ __forceinline__ Fq pair_xchg(const Fq &v) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 12; ++k) r.l[k] = __shfl_xor_sync(0xffffffffu, v.l[k], 1);
    return r;
}
__device__ __forceinline__ Fq pair_sel(bool c, const Fq &a, const Fq &b) {   // c ? a : b
    Fq r;
#pragma unroll
    for (int k = 0; k < 12; ++k) r.l[k] = c ? a.l[k] : b.l[k];
    return r;
}

// An Fq2 multiplication a * b on lane h is  a.own * x + a.oth * y  with (x, y) = (b0, -b1) on lane 0 and (b0, b1) on
// lane 1.  The left operand needs both components as they are (PairA), the right one in this prepared form (PairB);
// a value used in several products is prepared once.
struct PairA {
    Fq own, oth;
};
struct PairB {
    Fq x, y;
};
__device__ __forceinline__ PairA pair_a(const Fq &own) { return PairA{own, pair_xchg(own)}; }
__device__ __forceinline__ PairB pair_b(const Fq &own, int h) {
    Fq got = pair_xchg(pair_sel(h, Fq::neg(own), own));   // lane 0 receives -b1, lane 1 receives b0
    return PairB{pair_sel(h, got, own), pair_sel(h, own, got)};
}
// own component of a * b
__device__ __forceinline__ Fq pair_mul(const PairA &a, const PairB &b) { return Fq::mul_dual(a.own, b.x, a.oth, b.y); }
// own component of a^2
__device__ __forceinline__ Fq pair_sqr(const PairA &a, int h) {
    Fq l = pair_sel(h, a.own, Fq::add(a.own, a.oth));
    Fq r = pair_sel(h, a.oth, Fq::sub(a.own, a.oth));
    Fq m = Fq::mul(l, r);
    return pair_sel(h, Fq::dbl(m), m);
}
// bits set on both lanes of the pair (zero tests of Fq2 values: each lane contributes its component)
__device__ __forceinline__ uint32_t pair_flags(uint32_t own_bits) {
    return own_bits & __shfl_xor_sync(0xffffffffu, own_bits, 1);
}

// half of an XYZZ<Fq2> accumulator: component h of X, Y, ZZ, ZZZ
struct PairXYZZ {
    Fq x, y, zz, zzz;
};
__device__ __forceinline__ Fq pair_one(int h) { return pair_sel(h, Fq::zero(), Fq::one()); }   // Fq2::one() = (R, 0)

// 2 (px, py) as half an XYZZ; (px, py) finite with py != 0 where the result is used.  Every lane of the warp runs it.
// (Operands and result by value: a reference would pin the caller's accumulator to local memory.)
static __device__ __noinline__ PairXYZZ pair_mdbl(Fq px, Fq py, int h) {
    PairA u = pair_a(Fq::dbl(py));
    Fq v = pair_sqr(u, h);
    PairB vb = pair_b(v, h);
    Fq w = pair_mul(u, vb);
    PairA x2 = pair_a(px);
    Fq s = pair_mul(x2, vb);
    Fq xx = pair_sqr(x2, h);
    PairA m = pair_a(Fq::add(Fq::dbl(xx), xx));
    Fq x3 = Fq::sub(pair_sqr(m, h), Fq::dbl(s));
    Fq t1 = pair_mul(m, pair_b(Fq::sub(s, x3), h));
    Fq t2 = pair_mul(pair_a(w), pair_b(py, h));
    return PairXYZZ{x3, Fq::sub(t1, t2), v, w};
}

// acc += (px, py) on every pair with `live` set (EFD madd-2008-s); every lane of the warp must call it.  The order
// below keeps few values alive (every primitive is a volatile asm, so the source order is the order ptxas sees):
// exceptional cases are settled as soon as p and r exist, and every coordinate of the sum replaces the old one --
// under the pair's `normal` predicate -- as soon as nothing else reads it.
__device__ __forceinline__ void pair_madd(PairXYZZ &acc, const Fq &px, const Fq &py, bool live, int h) {
    PairA p, r;
    {
        Fq u2 = pair_mul(pair_a(acc.zz), pair_b(px, h));
        p.own = Fq::sub(u2, acc.x);
        Fq s2 = pair_mul(pair_a(acc.zzz), pair_b(py, h));
        r.own = Fq::sub(s2, acc.y);
    }
    // bit 0: point at infinity, bit 1: accumulator at infinity, bit 2: p == 0, bit 3: r == 0 (each over both components)
    const uint32_t f = pair_flags(((px.is_zero() && py.is_zero()) ? 1u : 0u) | (acc.zz.is_zero() ? 2u : 0u) |
                                  (p.own.is_zero() ? 4u : 0u) | (r.own.is_zero() ? 8u : 0u));
    const bool add = live && !(f & 1u);   // something to add
    const bool ainf = (f & 2u) != 0, pz = (f & 4u) != 0, rz = (f & 8u) != 0;
    const bool normal = add && !ainf && !pz;
    const bool need_dbl = add && !ainf && pz && rz;
    if (add && ainf) {
        acc.x = px; acc.y = py; acc.zz = pair_one(h); acc.zzz = pair_one(h);
    } else if (add && pz && !rz) {   // P + (-P)
        acc.x = pair_one(h); acc.y = pair_one(h); acc.zz = Fq::zero(); acc.zzz = Fq::zero();
    }
    if (__any_sync(0xffffffffu, need_dbl)) {   // P + P somewhere in this warp: everybody doubles, the pair concerned keeps it
        PairXYZZ d = pair_mdbl(px, py, h);
        if (need_dbl) acc = d;
    }
    // from here on the lanes that are not `normal` compute on whatever they hold and discard it
    p.oth = pair_xchg(p.own);
    PairB pp = pair_b(pair_sqr(p, h), h);
    Fq ppp = pair_mul(p, pp);
    {
        Fq zz3 = pair_mul(pair_a(acc.zz), pp);
        acc.zz = pair_sel(normal, zz3, acc.zz);
    }
    Fq q = pair_mul(pair_a(acc.x), pp);
    PairB pppb = pair_b(ppp, h);
    {
        Fq zzz3 = pair_mul(pair_a(acc.zzz), pppb);
        acc.zzz = pair_sel(normal, zzz3, acc.zzz);
    }
    Fq t2 = pair_mul(pair_a(acc.y), pppb);
    r.oth = pair_xchg(r.own);
    Fq x3 = Fq::sub(Fq::sub(pair_sqr(r, h), ppp), Fq::dbl(q));
    Fq y3 = Fq::sub(pair_mul(r, pair_b(Fq::sub(q, x3), h)), t2);
    acc.x = pair_sel(normal, x3, acc.x);
    acc.y = pair_sel(normal, y3, acc.y);
}

__device__ __forceinline__ Fq pair_load_fq(const uint32_t *src) {   // 12 words, 16-byte aligned
    Fq r;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint4 v = s4[j];
        r.l[4 * j] = v.x; r.l[4 * j + 1] = v.y; r.l[4 * j + 2] = v.z; r.l[4 * j + 3] = v.w;
    }
    return r;
}
__device__ __forceinline__ Fq pair_ldg_fq(const uint32_t *src) {
    Fq r;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint4 v = __ldg(s4 + j);
        r.l[4 * j] = v.x; r.l[4 * j + 1] = v.y; r.l[4 * j + 2] = v.z; r.l[4 * j + 3] = v.w;
    }
    return r;
}
__device__ __forceinline__ void pair_store_fq(uint32_t *dst, const Fq &v) {
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int j = 0; j < 3; ++j) d4[j] = make_uint4(v.l[4 * j], v.l[4 * j + 1], v.l[4 * j + 2], v.l[4 * j + 3]);
}
// component h of the XYZZ<Fq2> stored at index idx (X.c0 X.c1 Y.c0 Y.c1 ZZ.c0 ZZ.c1 ZZZ.c0 ZZZ.c1, 12 words each)
__device__ __forceinline__ PairXYZZ pair_load_xyzz(const uint32_t *src, size_t idx, int h) {
    const uint32_t *s = src + idx * 96 + h * 12;
    return PairXYZZ{pair_load_fq(s), pair_load_fq(s + 24), pair_load_fq(s + 48), pair_load_fq(s + 72)};
}
__device__ __forceinline__ void pair_store_xyzz(uint32_t *dst, size_t idx, int h, const PairXYZZ &p) {
    uint32_t *d = dst + idx * 96 + h * 12;
    pair_store_fq(d, p.x); pair_store_fq(d + 24, p.y); pair_store_fq(d + 48, p.zz); pair_store_fq(d + 72, p.zzz);
}

// The G2 hot kernel: BucketAccumulate<Fq2, ADD_TO> (msm_kernels.cuh) with one lane pair per work item.
template <bool ADD_TO>
__global__ void __launch_bounds__(G16_PAIR_BLOCK, G16_PAIR_MIN_BLOCKS)
accumulate_pair_g2_kernel(const uint32_t *pts, const uint32_t *entries, const WorkItem *items, const uint32_t *n_items,
                          uint32_t *buckets, uint32_t *chunk_out) {
    const size_t t = ((size_t)blockIdx.x * G16_PAIR_BLOCK + threadIdx.x) >> 1;
    const int h = threadIdx.x & 1;
    bool active = t < *n_items;   // the launch covers an upper bound; the exact count lives on the device
    WorkItem it = active ? items[t] : WorkItem{0u, 0u, 0u};
    const bool split = (it.bucket & SPLIT_FLAG) != 0;
    PairXYZZ acc{pair_one(h), pair_one(h), Fq::zero(), Fq::zero()};
    if (ADD_TO && active && !split) {
        if (it.begin == it.end) active = false;
        else acc = pair_load_xyzz(buckets, it.bucket, h);
    }
    const uint32_t len = active ? it.end - it.begin : 0u;
    const uint32_t steps = __reduce_max_sync(0xffffffffu, len);
    for (uint32_t k = 0; k < steps; ++k) {
        const bool live = k < len;
        Fq px = Fq::zero(), py = Fq::zero();
        if (live) {
            uint32_t v = entries[it.begin + k];
            const uint32_t *src = pts + (size_t)(v & 0x7fffffffu) * 48 + h * 12;
            px = pair_ldg_fq(src);
            py = pair_ldg_fq(src + 24);
            if (v >> 31) py = Fq::neg(py);
        }
        pair_madd(acc, px, py, live, h);
    }
    if (!active) return;
    if (split) pair_store_xyzz(chunk_out, t, h, acc);
    else pair_store_xyzz(buckets, it.bucket, h, acc);
}

#ifdef G16_PAIR_DEBUG_KERNEL   // k_debug.cu only
// test hook: out = affine(P + Q), P through from_affine and Q through the lane-pair mixed addition
static __global__ void __launch_bounds__(64) debug_pair_add_kernel(size_t n, const uint32_t *p, const uint32_t *q, uint32_t *out) {
    const size_t i = ((size_t)blockIdx.x * 64 + threadIdx.x) >> 1;
    const int h = threadIdx.x & 1;
    const bool live = i < n;
    const size_t j = live ? i : 0;
    Fq ax = pair_load_fq(p + j * 48 + h * 12), ay = pair_load_fq(p + j * 48 + 24 + h * 12);
    Fq bx = pair_load_fq(q + j * 48 + h * 12), by = pair_load_fq(q + j * 48 + 24 + h * 12);
    PairXYZZ acc{pair_one(h), pair_one(h), Fq::zero(), Fq::zero()};
    pair_madd(acc, ax, ay, live, h);   // infinity + P: the from_affine of the one-thread hook
    pair_madd(acc, bx, by, live, h);
    // to affine on lane 0 of the pair: gather the other component
    PairA X = pair_a(acc.x), Y = pair_a(acc.y), ZZ = pair_a(acc.zz), ZZZ = pair_a(acc.zzz);
    if (live && h == 0) {
        XYZZ<Fq2> full{Fq2{X.own, X.oth}, Fq2{Y.own, Y.oth}, Fq2{ZZ.own, ZZ.oth}, Fq2{ZZZ.own, ZZZ.oth}};
        store_affine<Fq2>(out, i, xyzz_to_affine(full));
    }
}
#endif

}  // namespace g16
#endif
