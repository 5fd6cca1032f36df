// Bucket accumulation in AFFINE coordinates with block-shared inversions (Montgomery's trick).
//
// Replaces the same N * W mixed additions of ark-ec 0.4.2 `msm_bigint_wnaf`
// (/root/reference/crates/groth16-core/src/lib.rs:282,296) as BucketAccumulate (msm_kernels.cuh); the
// bucket sums are the same group elements, so everything downstream is unchanged.
//
// An affine addition costs one inversion, two multiplications and a squaring.  When the inversions of many
// independent additions are shared (3 multiplications each plus ONE inversion for the whole batch) the
// price is 5M + 1S = 6 field multiplications instead of the 8M + 2S = 10 of the XYZZ mixed addition the
// hot kernel is bound by.  Independent additions come from summing a bucket as a TREE:
//
//   round 0   entries (e0 e1)(e2 e3)...    gathered from the base table, signs applied  -> buf[0]
//   round r   points of buf[r-1] pairwise                                                -> buf[r]
//   tail      what is left after `rounds` rounds (len / 2^rounds points) joins an XYZZ accumulator
//
// One thread owns one work item (a whole bucket; the chunks of split buckets stay on the XYZZ kernel) and
// walks its pairs twice per round:
//   phase 1   d_j = x2 - x1 of every pair, running product; the product *before* d_j is parked in the
//             first half of output slot j
//   block     the 128 running products of the block are multiplied up a shared-memory tree, thread 0
//             inverts the root (binary extended Euclid, fp.cuh), the inverses come back down the tree:
//             one inversion per (block, round) = per several thousand additions.  Other resident blocks
//             keep the multiplier busy meanwhile.
//   phase 2   backwards: 1/d_j = inv_run * prefix_j, inv_run *= d_j, lambda = (y2 - y1) / d_j,
//             x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1   -> output slot j
// Exceptional pairs (an operand at infinity, P + P, P - P) contribute d = 1 (or 2 y for a doubling) so the
// shared product never vanishes, and are resolved in phase 2.
//
// Scratch layout: round r of the item with bucket number g writes slots [o_r, o_r + ceil(len_r / 2)) of
// buf[r], o_0 = begin / 2 + g, o_r = o_(r-1) / 2 + g -- monotone in g with gaps >= the slot count, so no
// offsets have to be scanned; buf[r] holds entries / 2^(r+1) + (2 - 2^-r) * buckets + 1 slots.
#pragma once
#include "msm_kernels.cuh"

namespace g16 {

constexpr int AFF_BLOCK = 128;
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
    // denominator of pair (pos, pos + 1) from the x coordinates alone whenever that decides the case
    G16_HD static F pair_denominator(const Src &s, size_t pos) {
        F x1 = load_x(s, pos), x2 = load_x(s, pos + 1);
        F d = F::sub(x2, x1);
        if (d.is_zero() || x1.is_zero() || x2.is_zero()) {
            Affine<F> p = load(s, pos), q = load(s, pos + 1);
            pair_case(p, q, d);
        }
        return d;
    }

    // phase 1: product of the m denominators of pairs (src + 2 j, src + 2 j + 1); the product before pair j is
    // parked in slot dst_off + j
    G16_HD static F phase1(const Src &s, size_t src_off, uint32_t m, uint32_t *dst, size_t dst_off) {
        F run = F::one();
        for (uint32_t j = 0; j < m; ++j) {
            F d = pair_denominator(s, src_off + 2 * (size_t)j);
            store_f(dst, dst_off + j, run);
            run = F::mul(run, d);
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

#if !defined(G16_EMU) && defined(__CUDACC__)
// 1 / v for every thread of the block (v != 0).  tree: 2 * AFF_BLOCK field elements of shared memory.
template <class F>
__device__ __forceinline__ F block_inverse(uint32_t *tree, const F &v) {
    constexpr int B = AFF_BLOCK, W = F::N;
    const int t = threadIdx.x;
    auto put = [&](int node, const F &x) {
        const uint32_t *s = limbs(x);
#pragma unroll
        for (int k = 0; k < W; ++k) tree[(size_t)node * W + k] = s[k];
    };
    auto get = [&](int node) {
        F x;
        uint32_t *d = limbs(x);
#pragma unroll
        for (int k = 0; k < W; ++k) d[k] = tree[(size_t)node * W + k];
        return x;
    };
    put(B + t, v);
    __syncthreads();
#pragma unroll 1
    for (int s = B >> 1; s >= 1; s >>= 1) {
        if (t < s) put(s + t, F::mul(get(2 * (s + t)), get(2 * (s + t) + 1)));
        __syncthreads();
    }
    if (t == 0) put(1, field_inv_call(get(1)));
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s < B; s <<= 1) {
        if (t < s) {
            int i = s + t;
            F inv = get(i), l = get(2 * i), r = get(2 * i + 1);
            put(2 * i, F::mul(inv, r));
            put(2 * i + 1, F::mul(inv, l));
        }
        __syncthreads();
    }
    F out = get(B + t);
    __syncthreads();   // the tree is reused by the next round
    return out;
}

template <class F>
__global__ void __launch_bounds__(AFF_BLOCK) accumulate_affine_kernel(const uint32_t *pts, const uint32_t *entries,
                                                                      const WorkItem *items, const uint32_t *first_item,
                                                                      const uint32_t *n_items, uint32_t rounds,
                                                                      uint32_t *scratch, size_t n_entries, size_t n_buckets,
                                                                      uint32_t *buckets) {
    extern __shared__ uint32_t tree[];
    using A = AffineAcc<F>;
    const size_t t = (size_t)blockIdx.x * AFF_BLOCK + threadIdx.x + *first_item;
    const bool live = t < *n_items;
    WorkItem it = live ? items[t] : WorkItem{0u, 0u, 0u};
    uint32_t len = it.end - it.begin;
    size_t off = it.begin;
    typename A::Src src{pts, entries};
    uint32_t *buf = scratch;
    size_t slots = n_entries / 2 + n_buckets + 1;
#pragma unroll 1
    for (uint32_t r = 0; r < rounds; ++r) {
        if (!__syncthreads_or(len > 1)) break;
        size_t o = off / 2 + it.bucket;
        F run = A::phase1(src, off, len >> 1, buf, o);
        F inv = block_inverse<F>(tree, run);
        A::phase2(src, off, len, buf, o, inv);
        len = (len + 1) >> 1;
        off = o;
        src = typename A::Src{buf, nullptr};
        buf += slots * (2 * F::N);
        slots = slots / 2 + n_buckets + 1;
    }
    if (!live) return;
    XYZZ<F> acc = A::tail(src, off, len);
    store_xyzz<F>(buckets, it.bucket, acc);
}
#endif

// Serial statement of the same schedule (host emulation build): every item runs its rounds on its own, with
// its own inversion -- same slots, same sums.
template <class F>
struct AccumulateAffineSerial {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t t0, const uint32_t *pts, const uint32_t *entries, const WorkItem *items,
                           const uint32_t *first_item, const uint32_t *n_items, uint32_t rounds, uint32_t *scratch,
                           size_t n_entries, size_t n_buckets, uint32_t *buckets) {
        using A = AffineAcc<F>;
        size_t t = t0 + *first_item;
        if (t >= *n_items) return;
        WorkItem it = items[t];
        uint32_t len = it.end - it.begin;
        size_t off = it.begin;
        typename A::Src src{pts, entries};
        uint32_t *buf = scratch;
        size_t slots = n_entries / 2 + n_buckets + 1;
        for (uint32_t r = 0; r < rounds && len > 1; ++r) {
            size_t o = off / 2 + it.bucket;
            F run = A::phase1(src, off, len >> 1, buf, o);
            A::phase2(src, off, len, buf, o, F::inv(run));
            len = (len + 1) >> 1;
            off = o;
            src = typename A::Src{buf, nullptr};
            buf += slots * (2 * F::N);
            slots = slots / 2 + n_buckets + 1;
        }
        XYZZ<F> acc = A::tail(src, off, len);
        store_xyzz<F>(buckets, it.bucket, acc);
    }
};

}  // namespace g16
