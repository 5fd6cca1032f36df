// Bucket accumulation in AFFINE coordinates with shared inversions (Montgomery's trick).
//
// Replaces the same N * W mixed additions of ark-ec 0.4.2 `msm_bigint_wnaf`
// (/root/reference/crates/groth16-core/src/lib.rs:282,296) as BucketAccumulate (msm_kernels.cuh); the
// bucket sums are the same group elements, so everything downstream is unchanged.
//
// An affine addition costs one inversion, two multiplications and a squaring.  When the inversions of many
// independent additions are shared (3 multiplications each plus ONE inversion per batch) the price is
// 5M + 1S = 6 field multiplications instead of the 8M + 2S = 10 of the XYZZ mixed addition the hot kernel is
// bound by.  Independent additions come from summing a bucket as a TREE:
//
//   round 0   entries (e0 e1)(e2 e3)...    gathered from the base table, signs applied  -> buf[0]
//   round r   points of buf[r-1] pairwise                                                -> buf[r]
//   tail      what is left after `rounds` rounds (len / 2^rounds points) joins an XYZZ accumulator
//
// One thread owns one work item (a whole bucket; the chunks of split buckets stay on the XYZZ kernel).  A round
// is three launches of independent threads -- no barriers, small kernels, registers per phase:
//   AffinePhase1   d_j = x2 - x1 of every pair of the item, running product; the product *before* d_j is parked
//                  in the first half of output slot j; the item's total goes to totals[item]
//   BatchInverse   totals[] -> 1 / totals[] in groups of AFF_INV_GROUP (Montgomery's trick again: one
//                  inversion per group = per several thousand additions)
//   AffinePhase2   backwards: 1/d_j = inv_run * prefix_j, inv_run *= d_j, lambda = (y2 - y1) / d_j,
//                  x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1   -> output slot j
// Exceptional pairs (an operand at infinity, P + P, P - P) contribute d = 1 (or 2 y for a doubling) so no
// product ever vanishes, and are resolved in phase 2.
//
// Scratch layout: round r of the item with bucket number g writes slots [o_r, o_r + ceil(len_r / 2)) of
// buf[r], o_0 = begin / 2 + g, o_r = o_(r-1) / 2 + g -- monotone in g with gaps >= the slot count, so no
// offsets have to be scanned; buf[r] holds entries / 2^(r+1) + (2 - 2^-r) * buckets + 1 slots.
//
// (A first version ran all rounds in ONE kernel with a block-wide product tree and one inverting thread per
// block: bit-exact but slower than the XYZZ walk -- barrier waits, instruction-cache misses of the large
// kernel and 1.25 independent multiplications per thread left the multiplier 59 % busy;
// profiles/r01_run13_affine_block_kernel.md.)
#pragma once
#include "msm_kernels.cuh"

namespace g16 {

constexpr uint32_t AFF_MAX_ROUNDS = 8;

// slots of round-r scratch buffers, and their sum, for `entries` sorted entries over `buckets` buckets
inline size_t affine_round_slots(size_t entries, size_t buckets, uint32_t r) {
    size_t s = entries / 2 + buckets + 1;
    for (uint32_t k = 0; k < r; ++k) s = s / 2 + buckets + 1;
    return s;
}

template <class F>
struct AffineAcc {
    // source of a round: the base table through the sorted entries (round 0) or the previous round's slots
    struct Src {
        const uint32_t *pts;       // table (gather) or previous buffer
        const uint32_t *entries;   // non-null: gather
    };
    G16_HD static Affine<F> load(const Src &s, size_t pos) {
        if (s.entries) {
            uint32_t v = s.entries[pos];
            Affine<F> p = load_affine<F>(s.pts, v & 0x7fffffffu);
            if (v >> 31) p.y = F::neg(p.y);
            return p;
        }
        // slots written earlier in this kernel: coherent loads (the read-only path of load_affine is only
        // safe for data no thread of the running kernel writes)
        Affine<F> p;
        p.x = load_f(s.pts, pos);
        p.y = load_f(s.pts + F::N, pos);
        return p;
    }
    G16_HD static F load_x(const Src &s, size_t pos) {
        if (!s.entries) return load_f(s.pts, pos);
        F x;
        const uint32_t *src = s.pts + (size_t)(s.entries[pos] & 0x7fffffffu) * (2 * F::N);
        uint32_t *d = limbs(x);
#if G16_DEVICE_CODE
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll
        for (int j = 0; j < F::N / 4; ++j) {
            uint4 v = __ldg(s4 + j);
            d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
        }
#else
        for (int j = 0; j < F::N; ++j) d[j] = src[j];
#endif
        return x;
    }
    G16_HD static void store_f(uint32_t *dst, size_t slot, const F &v) {   // first half of a slot
        uint32_t *d = dst + slot * (2 * F::N);
        const uint32_t *s = limbs(v);
#if G16_DEVICE_CODE
        uint4 *d4 = reinterpret_cast<uint4 *>(d);
#pragma unroll
        for (int j = 0; j < F::N / 4; ++j) d4[j] = make_uint4(s[4 * j], s[4 * j + 1], s[4 * j + 2], s[4 * j + 3]);
#else
        for (int j = 0; j < F::N; ++j) d[j] = s[j];
#endif
    }
    G16_HD static F load_f(const uint32_t *src, size_t slot) {
        F x;
        const uint32_t *s = src + slot * (2 * F::N);
        uint32_t *d = limbs(x);
#if G16_DEVICE_CODE
        const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
#pragma unroll
        for (int j = 0; j < F::N / 4; ++j) {
            uint4 v = s4[j];   // written by this thread in phase 1: a plain (coherent) load
            d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
        }
#else
        for (int j = 0; j < F::N; ++j) d[j] = s[j];
#endif
        return x;
    }

    // Classification of one pair; the denominator both phases agree on.
    enum { NORMAL = 0, TAKE_Q = 1, TAKE_P = 2, DOUBLE = 3, CANCEL = 4 };
    G16_HD static int pair_case(const Affine<F> &p, const Affine<F> &q, F &d) {
        if (p.is_inf()) { d = F::one(); return TAKE_Q; }
        if (q.is_inf()) { d = F::one(); return TAKE_P; }
        d = F::sub(q.x, p.x);
        if (!d.is_zero()) return NORMAL;
        if (p.y == q.y && !p.y.is_zero()) { d = F::dbl(p.y); return DOUBLE; }
        d = F::one();
        return CANCEL;
    }
    // phase 1: product of the m denominators of pairs (src + 2 j, src + 2 j + 1); the product before pair j is
    // parked in slot dst_off + j
    G16_HD static F phase1(const Src &s, size_t src_off, uint32_t m, uint32_t *dst, size_t dst_off) {
        F run = F::one();
        if (m == 0) return run;
        // the x coordinates of the next pair are requested before the multiplication of this one
        F x1 = load_x(s, src_off), x2 = load_x(s, src_off + 1);
        for (uint32_t j = 0; j < m; ++j) {
            size_t pos = src_off + 2 * (size_t)j;
            F nx1 = x1, nx2 = x2;
            if (j + 1 < m) { nx1 = load_x(s, pos + 2); nx2 = load_x(s, pos + 3); }
            F d = F::sub(x2, x1);
            if (d.is_zero() || x1.is_zero() || x2.is_zero()) {
                Affine<F> p = load(s, pos), q = load(s, pos + 1);
                pair_case(p, q, d);
            }
            store_f(dst, dst_off + j, run);
            run = F::mul(run, d);
            x1 = nx1; x2 = nx2;
        }
        return run;
    }
    // phase 2: inv_run = 1 / (product of all m denominators); writes the m sums, and the unpaired last point
    // (len odd) behind them
    G16_HD static void phase2(const Src &s, size_t src_off, uint32_t len, uint32_t *dst, size_t dst_off, F inv_run) {
        uint32_t m = len >> 1;
        if (len & 1u) {
            Affine<F> p = load(s, src_off + len - 1);
            store_affine_pt<F>(dst, dst_off + m, p);
        }
        for (uint32_t j = m; j-- > 0;) {
            Affine<F> p = load(s, src_off + 2 * (size_t)j), q = load(s, src_off + 2 * (size_t)j + 1);
            F d;
            int c = pair_case(p, q, d);
            F inv_d = F::mul(inv_run, load_f(dst, dst_off + j));
            inv_run = F::mul(inv_run, d);
            Affine<F> r;
            if (c == NORMAL || c == DOUBLE) {
                F num;
                if (c == NORMAL) num = F::sub(q.y, p.y);
                else { F xx = F::sqr(p.x); num = F::add(F::dbl(xx), xx); }
                F lam = F::mul(num, inv_d);
                r.x = F::sub(F::sub(F::sqr(lam), p.x), q.x);
                r.y = F::sub(F::mul(lam, F::sub(p.x, r.x)), p.y);
            } else if (c == TAKE_Q) r = q;
            else if (c == TAKE_P) r = p;
            else r = Affine<F>::inf();
            store_affine_pt<F>(dst, dst_off + j, r);
        }
    }
    // what is left of the item joins an XYZZ accumulator (the form the bucket reduction consumes)
    G16_HD static XYZZ<F> tail(const Src &s, size_t src_off, uint32_t len) {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t j = 0; j < len; ++j) {
            Affine<F> p = load(s, src_off + j);
            xyzz_madd(acc, p.x, p.y);
        }
        return acc;
    }
};

// Position of an item in round r: source (len, off) and destination slot o of that round.
template <class F>
struct AffGeom {
    uint32_t len;
    size_t off, o;
    typename AffineAcc<F>::Src src;
    uint32_t *dst;
};
template <class F>
G16_HD AffGeom<F> affine_geometry(const WorkItem &it, uint32_t r, const uint32_t *pts, const uint32_t *entries,
                                  uint32_t *scratch, size_t n_entries, size_t n_buckets) {
    AffGeom<F> g;
    g.len = it.end - it.begin;
    g.off = it.begin;
    g.src = typename AffineAcc<F>::Src{pts, entries};
    g.dst = scratch;
    size_t slots = n_entries / 2 + n_buckets + 1;
    for (uint32_t k = 0; k < r; ++k) {
        g.off = g.off / 2 + it.bucket;
        g.len = (g.len + 1) >> 1;
        g.src = typename AffineAcc<F>::Src{g.dst, nullptr};
        g.dst += slots * (2 * F::N);
        slots = slots / 2 + n_buckets + 1;
    }
    g.o = g.off / 2 + it.bucket;
    return g;
}

constexpr uint32_t AFF_INV_GROUP = 32;

// one thread per slot of the item array behind *first_item (dead slots contribute 1 to their inversion group)
template <class F>
struct AffinePhase1 {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t0, const uint32_t *pts, const uint32_t *entries, const WorkItem *items,
                           const uint32_t *first_item, const uint32_t *n_items, uint32_t r, uint32_t *scratch,
                           size_t n_entries, size_t n_buckets, uint32_t *totals) {
        using A = AffineAcc<F>;
        size_t t = t0 + *first_item;
        F run = F::one();
        if (t < *n_items) {
            AffGeom<F> g = affine_geometry<F>(items[t], r, pts, entries, scratch, n_entries, n_buckets);
            run = A::phase1(g.src, g.off, g.len >> 1, g.dst, g.o);
        }
        A::store_f(totals, t0, run);   // totals: one field element per 2 * F::N words (slot layout)
    }
};

// totals[i] <- 1 / totals[i] for i < n.  Every thread owns AFF_INV_GROUP consecutive elements (Montgomery's trick);
// on the device the 32 running products of a warp are combined with shuffles (exclusive prefix and suffix products)
// and all lanes invert the SAME warp total: one inversion per 1024 elements, and -- the operands being identical
// across the warp -- the data-dependent loops of the binary Euclid inversion do not diverge.  (One inversion per
// thread cost 1.76 ms per round at 2^21 items: 32 lanes, 32 different branch histories.)
template <class F>
struct BatchInverse {
    static constexpr int BLOCK = 64;
    G16_HD static F forward(uint32_t *totals, size_t lo, size_t hi) {
        using A = AffineAcc<F>;
        F run = F::one();
        for (size_t i = lo; i < hi; ++i) {
            F v = A::load_f(totals, i);
            A::store_f(totals + F::N, i, run);   // second half of the slot: product of the earlier elements
            run = F::mul(run, v);
        }
        return run;
    }
    G16_HD static void backward(uint32_t *totals, size_t lo, size_t hi, F inv) {
        using A = AffineAcc<F>;
        for (size_t i = hi; i-- > lo;) {
            F v = A::load_f(totals, i);
            A::store_f(totals, i, F::mul(inv, A::load_f(totals + F::N, i)));
            inv = F::mul(inv, v);
        }
    }
    // serial statement (host emulation build): one inversion per thread
    G16_HD static void run(size_t t, uint32_t *totals, size_t n) {
        size_t lo = t * AFF_INV_GROUP, hi = lo + AFF_INV_GROUP < n ? lo + AFF_INV_GROUP : n;
        F run = forward(totals, lo, hi);
        backward(totals, lo, hi, field_inv_call(run));
    }
};
#if !defined(G16_EMU) && defined(__CUDACC__)
template <class F>
__device__ __forceinline__ F warp_shift(const F &v, int delta, bool up) {
    F r;
    const uint32_t *s = limbs(v);
    uint32_t *d = limbs(r);
#pragma unroll
    for (int k = 0; k < F::N; ++k)
        d[k] = up ? __shfl_up_sync(0xffffffffu, s[k], delta) : __shfl_down_sync(0xffffffffu, s[k], delta);
    return r;
}
template <class F>
__global__ void __launch_bounds__(64) batch_inverse_kernel(uint32_t *totals, size_t n) {
    const size_t t = (size_t)blockIdx.x * 64 + threadIdx.x;   // the grid covers whole warps; idle lanes carry 1
    const int lane = threadIdx.x & 31;
    size_t lo = t * AFF_INV_GROUP, hi = lo + AFF_INV_GROUP;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    F run = BatchInverse<F>::forward(totals, lo, hi);
    F pre = run, suf = run;   // inclusive prefix / suffix products over the warp
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        F a = F::mul(pre, warp_shift(pre, d, true));
        F b = F::mul(suf, warp_shift(suf, d, false));
        pre = fsel(lane >= d, a, pre);
        suf = fsel(lane + d < 32, b, suf);
    }
    F total;
    {
        const uint32_t *s = limbs(pre);
        uint32_t *dd = limbs(total);
#pragma unroll
        for (int k = 0; k < F::N; ++k) dd[k] = __shfl_sync(0xffffffffu, s[k], 31);
    }
    F inv_total = field_inv_call(total);   // same operand in every lane
    F ep = warp_shift(pre, 1, true), es = warp_shift(suf, 1, false);
    ep = fsel(lane >= 1, ep, F::one());
    es = fsel(lane < 31, es, F::one());
    BatchInverse<F>::backward(totals, lo, hi, F::mul(inv_total, F::mul(ep, es)));
}
#endif

template <class F>
struct AffinePhase2 {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t0, const uint32_t *pts, const uint32_t *entries, const WorkItem *items,
                           const uint32_t *first_item, const uint32_t *n_items, uint32_t r, uint32_t *scratch,
                           size_t n_entries, size_t n_buckets, const uint32_t *totals) {
        using A = AffineAcc<F>;
        size_t t = t0 + *first_item;
        if (t >= *n_items) return;
        AffGeom<F> g = affine_geometry<F>(items[t], r, pts, entries, scratch, n_entries, n_buckets);
        A::phase2(g.src, g.off, g.len, g.dst, g.o, A::load_f(totals, t0));
    }
};

// what the rounds left of every item joins an XYZZ accumulator (the form the bucket reduction consumes)
template <class F>
struct AffineTail {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t0, const uint32_t *pts, const uint32_t *entries, const WorkItem *items,
                           const uint32_t *first_item, const uint32_t *n_items, uint32_t rounds, uint32_t *scratch,
                           size_t n_entries, size_t n_buckets, uint32_t *buckets) {
        size_t t = t0 + *first_item;
        if (t >= *n_items) return;
        WorkItem it = items[t];
        AffGeom<F> g = affine_geometry<F>(it, rounds, pts, entries, scratch, n_entries, n_buckets);
        XYZZ<F> acc = AffineAcc<F>::tail(g.src, g.off, g.len);
        store_xyzz<F>(buckets, it.bucket, acc);
    }
};

}  // namespace g16
