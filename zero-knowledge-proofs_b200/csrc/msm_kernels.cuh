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
#include "quad.cuh"
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

// Bucket code of one (scalar, window) digit: bucket index | sign << 31, or NO_DIGIT for a zero digit.
constexpr uint32_t NO_DIGIT = 0xffffffffu;

// One thread per scalar: canonical form, signed digits, bucket histogram, and the per-window code
// array codes[w * n + i] (window-major so that the scatter pass below streams it coalesced).
struct DigitDecompose {
    static constexpr int BLOCK = 256;
    // the launch covers scalars [i0, i0 + threads) of the n of the call (host scalars arrive in chunks)
    G16_HD static void run(size_t t, const uint32_t *scalars, bool mont, MsmPlan plan, size_t n, uint32_t *counts,
                           uint32_t *codes, uint32_t *ranks, size_t i0) {
        const size_t i = t + i0;
        uint32_t k[8];
        load_scalar(scalars, i, mont, k);
        DigitIter it(k, plan.c);
        for (uint32_t w = 0; w < plan.nwin; ++w) {
            int32_t d = it.next();
            uint32_t code = NO_DIGIT;
            if (d != 0) {
                uint32_t b = (uint32_t)(d < 0 ? -d : d) - 1u;
                // the histogram atomic also hands out this digit's rank inside its bucket, so the scatter
                // pass needs no second round of atomics (position = bucket offset + rank)
                ranks[(size_t)w * n + i] = atomic_add_u32(&counts[(plan.bwin == 1 ? 0u : w) * plan.nb + b], 1u);
                code = b | (d < 0 ? 0x80000000u : 0u);
            }
            codes[(size_t)w * n + i] = code;
        }
    }
};

// Counting-sort scatter with the ranks recorded by DigitDecompose, one thread per (window, scalar), window-major:
// no atomics, position = bucket offset + rank, entries[position] = point index | sign << 31.  Used while the entry
// array fits L2 (the scattered 4-byte writes merge there); larger arrays take the two-pass kernel below.
struct ScatterRanked {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t t, const uint32_t *codes, const uint32_t *ranks, MsmPlan plan, size_t n,
                           const uint32_t *offsets, uint32_t *entries) {
        uint32_t code = codes[t];
        if (code == NO_DIGIT) return;
        uint32_t w = (uint32_t)(t / n), i = (uint32_t)(t % n);
        uint32_t pos = offsets[(plan.bwin == 1 ? 0u : w) * plan.nb + (code & 0x7fffffffu)] + ranks[t];
        entries[pos] = (w * plan.stride + plan.offset + i) | (code & 0x80000000u);
    }
};

// Two-pass scatter for entry arrays far larger than L2 (shared bucket set at 2^24: 800 MB).  The one-pass kernel
// above writes 4 bytes at a random position per digit, and DRAM turns each of them into a read-modify-write of a
// whole line (6.9 ms for 201 M digits, 32 ms at 2^26).  Here the destination range is cut into partitions of
// 2^log_part entries:
//   pass A  a block takes SCATTER_TILE digits, counts them per destination partition in shared memory, reserves a
//           run in every partition's staging area with ONE global atomic per (block, partition) and appends
//           (position, entry) pairs there -- a few hundred open write frontiers that stay in L2 and leave it as
//           full lines.  The staging area of partition p is exactly [p << log_part, (p + 1) << log_part): positions
//           are dense, so no offsets are needed.
//   pass B  streams the staging area and stores every entry at its position: all stores of a warp fall into one
//           partition (2^log_part * 4 bytes, L2 resident while it is being filled).
// The result is the same `entries` array as ScatterRanked's (position = bucket offset + rank).
constexpr int SCATTER_THREADS = 256;
constexpr int SCATTER_PER_THREAD = 16;
constexpr int SCATTER_TILE = SCATTER_THREADS * SCATTER_PER_THREAD;
constexpr uint32_t SCATTER_MAX_PARTS = 1024;   // shared-memory counters per block
#if !defined(G16_EMU) && defined(__CUDACC__)
// shared memory: SCATTER_TILE (position, entry) pairs sorted by partition, then 3 x n_parts words
static __global__ void __launch_bounds__(SCATTER_THREADS) scatter_partition_kernel(const uint32_t *codes, const uint32_t *ranks,
                                                                            MsmPlan plan, size_t n, const uint32_t *offsets,
                                                                            uint32_t log_part, uint32_t n_parts,
                                                                            uint32_t *part_cursor, uint2 *staging) {
    extern __shared__ uint2 tile_buf[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(tile_buf + SCATTER_TILE);   // digits per partition, then fill cursor
    uint32_t *first = cnt + n_parts;                                        // first tile slot of the partition
    uint32_t *gbase = first + n_parts;                                      // first staging slot reserved for this block
    const size_t total = n * plan.nwin;
    const size_t tile0 = (size_t)blockIdx.x * SCATTER_TILE;
    for (uint32_t p = threadIdx.x; p < n_parts; p += SCATTER_THREADS) cnt[p] = 0;
    __syncthreads();
    // positions of this thread's digits stay in registers (0xffffffff: no digit), the sign bits in one word
    uint32_t pos[SCATTER_PER_THREAD], signs = 0;
    // (window, index) of a flattened digit number without a 64-bit division per digit: one per thread, then steps
    const uint32_t w0 = (uint32_t)(tile0 / n);
    const size_t r0 = tile0 - (size_t)w0 * n;
    auto window_index = [&](int k, uint32_t &w, uint32_t &i) {
        size_t r = r0 + (size_t)k * SCATTER_THREADS + threadIdx.x;
        w = w0;
        while (r >= n) { r -= n; ++w; }
        i = (uint32_t)r;
    };
#pragma unroll
    for (int k = 0; k < SCATTER_PER_THREAD; ++k) {
        size_t t = tile0 + (size_t)k * SCATTER_THREADS + threadIdx.x;
        uint32_t code = t < total ? codes[t] : NO_DIGIT;
        pos[k] = 0xffffffffu;
        if (code != NO_DIGIT) {
            uint32_t w, i;
            window_index(k, w, i);
            pos[k] = offsets[(plan.bwin == 1 ? 0u : w) * plan.nb + (code & 0x7fffffffu)] + ranks[t];
            signs |= (code >> 31) << k;
        }
    }
#pragma unroll
    for (int k = 0; k < SCATTER_PER_THREAD; ++k)
        if (pos[k] != 0xffffffffu) atomicAdd(&cnt[pos[k] >> log_part], 1u);
    __syncthreads();
    // exclusive prefix of the counters (n_parts <= 1024: four per thread, warp scan, warp totals through `gbase`)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t p = threadIdx.x * 4 + j;
            v[j] = p < n_parts ? cnt[p] : 0u;
            sum += v[j];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        __shared__ uint32_t warp_tot[SCATTER_THREADS / 32];
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int q = 0; q < warp; ++q) base += warp_tot[q];
        uint32_t run = base + incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t p = threadIdx.x * 4 + j;
            if (p < n_parts) {
                first[p] = run;
                gbase[p] = v[j] ? (p << log_part) + atomicAdd(&part_cursor[p], v[j]) : 0u;
                cnt[p] = run;   // fill cursor
            }
            run += v[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCATTER_PER_THREAD; ++k) {
        if (pos[k] == 0xffffffffu) continue;
        uint32_t w, i;
        window_index(k, w, i);
        uint32_t slot = atomicAdd(&cnt[pos[k] >> log_part], 1u);
        tile_buf[slot] = make_uint2(pos[k], (w * plan.stride + plan.offset + i) | (((signs >> k) & 1u) << 31));
    }
    __syncthreads();
    // linear write-out: consecutive threads carry consecutive members of a partition's run
    const uint32_t filled = first[n_parts - 1] + (cnt[n_parts - 1] - first[n_parts - 1]);
    for (uint32_t j = threadIdx.x; j < filled; j += SCATTER_THREADS) {
        uint2 e = tile_buf[j];
        uint32_t p = e.x >> log_part;
        staging[gbase[p] + (j - first[p])] = e;
    }
}
#endif
// pass B: one thread per staging slot below the entry count (offsets[total buckets])
struct ScatterFinal {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t s, const uint32_t *staging, const uint32_t *n_entries, uint32_t *entries) {
        if (s >= *n_entries) return;
        entries[staging[2 * s]] = staging[2 * s + 1];
    }
};

// ---- work items ------------------------------------------------------------------------------------
// A work item is a slice [begin, end) of one bucket's entries.  Buckets of up to ITEM_MAX entries are
// one item; longer ones are split into chunks of about max(ITEM_MAX, sqrt(size)) entries so that a
// skewed scalar distribution (boolean witnesses, repeated values) cannot serialise on one thread.
// Items are ordered by length, longest first, so the 32 lanes of a warp run loops of (nearly) equal
// length -- with Poisson-distributed bucket sizes this removes most of the divergence loss.
constexpr uint32_t ITEM_MAX = 256;             // upper limit of the per-call item length
constexpr uint32_t ITEM_BINS = ITEM_MAX + 2;   // bin 0: chunks of split buckets, bin 1 + (ITEM_MAX - len): whole buckets
struct WorkItem { uint32_t begin, end, bucket; };  // bucket | SPLIT_FLAG when it is a chunk
constexpr uint32_t SPLIT_FLAG = 0x80000000u;

// item_max (<= ITEM_MAX) is chosen per call by the host: long items only when there is so much work that a
// serial walk of item_max additions is negligible, short ones when the call is small and the longest item
// would be the critical path (a mixed addition has ~10 us of latency when few warps are resident).
G16_HD void item_shape(uint32_t size, uint32_t item_max, uint32_t &nch, uint32_t &len, uint32_t &bin) {
    if (size <= item_max) { nch = 1; len = size; bin = 1u + (ITEM_MAX - size); return; }
    nch = (size + item_max - 1) / item_max;
    len = (size + nch - 1) / nch;
    bin = 0;
}

struct ItemCount {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t g, const uint32_t *offsets, uint32_t item_max, uint32_t *bin_counts) {
        uint32_t nch, len, bin;
        item_shape(offsets[g + 1] - offsets[g], item_max, nch, len, bin);
        atomic_add_u32(&bin_counts[bin], nch);
    }
};

struct ItemScatter {
    static constexpr int BLOCK = 256;
    // split_list[0] counts the split buckets, split_list[1 + k] = {bucket, first item, chunks} (3 words each)
    G16_HD static void run(size_t g, const uint32_t *offsets, uint32_t item_max, uint32_t *bin_cursor, WorkItem *items,
                           uint32_t *split_list) {
        uint32_t begin = offsets[g], size = offsets[g + 1] - begin;
        uint32_t nch, len, bin;
        item_shape(size, item_max, nch, len, bin);
        uint32_t pos = atomic_add_u32(&bin_cursor[bin], nch);
        if (nch > 1) {
            uint32_t k = atomic_add_u32(&split_list[0], 1u);
            split_list[1 + 3 * k] = (uint32_t)g;
            split_list[2 + 3 * k] = pos;
            split_list[3 + 3 * k] = nch;
        }
        for (uint32_t j = 0; j < nch; ++j) {
            uint32_t b = begin + j * len;
            uint32_t e = b + len < begin + size ? b + len : begin + size;
            if (b > begin + size) b = begin + size;
            items[pos + j] = WorkItem{b, e, (uint32_t)g | (nch > 1 ? SPLIT_FLAG : 0u)};
        }
    }
};

// Device versions of the two kernels above: the 258 length bins are hot addresses (2^21 buckets hit them at 2^24),
// so a block first bins its 256 buckets in shared memory and then touches every non-empty global bin once.
#if !defined(G16_EMU) && defined(__CUDACC__)
static __global__ void __launch_bounds__(256) item_count_kernel(size_t buckets, const uint32_t *offsets, uint32_t item_max,
                                                                uint32_t *bin_counts) {
    __shared__ uint32_t cnt[ITEM_BINS];
    for (uint32_t b = threadIdx.x; b < ITEM_BINS; b += 256) cnt[b] = 0;
    __syncthreads();
    size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (g < buckets) {
        uint32_t nch, len, bin;
        item_shape(offsets[g + 1] - offsets[g], item_max, nch, len, bin);
        atomicAdd(&cnt[bin], nch);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < ITEM_BINS; b += 256)
        if (cnt[b]) atomicAdd(&bin_counts[b], cnt[b]);
}
static __global__ void __launch_bounds__(256) item_scatter_kernel(size_t buckets, const uint32_t *offsets, uint32_t item_max,
                                                                  uint32_t *bin_cursor, WorkItem *items, uint32_t *split_list) {
    __shared__ uint32_t cnt[ITEM_BINS];
    for (uint32_t b = threadIdx.x; b < ITEM_BINS; b += 256) cnt[b] = 0;
    __syncthreads();
    size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t begin = 0, size = 0, nch = 0, len = 0, bin = 0, local = 0;
    if (g < buckets) {
        begin = offsets[g];
        size = offsets[g + 1] - begin;
        item_shape(size, item_max, nch, len, bin);
        local = atomicAdd(&cnt[bin], nch);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < ITEM_BINS; b += 256)
        if (cnt[b]) cnt[b] = atomicAdd(&bin_cursor[b], cnt[b]);   // becomes the block's base in that bin
    __syncthreads();
    if (g >= buckets) return;
    uint32_t pos = cnt[bin] + local;
    if (nch > 1) {
        uint32_t k = atomicAdd(&split_list[0], 1u);
        split_list[1 + 3 * k] = (uint32_t)g;
        split_list[2 + 3 * k] = pos;
        split_list[3 + 3 * k] = nch;
    }
    for (uint32_t j = 0; j < nch; ++j) {
        uint32_t b = begin + j * len;
        uint32_t e = b + len < begin + size ? b + len : begin + size;
        if (b > begin + size) b = begin + size;
        items[pos + j] = WorkItem{b, e, (uint32_t)g | (nch > 1 ? SPLIT_FLAG : 0u)};
    }
}
#endif

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
G16_HD void store_affine_pt(uint32_t *dst, size_t idx, const Affine<F> &a) {
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

// The hot kernel: one thread per work item.  Whole buckets are written straight to `buckets`; chunks of
// split buckets go to chunk_out[item index] (split items occupy the front of the item array) and are
// folded by ChunkMerge.  ADD_TO: the call continues an MSM whose earlier scalar chunks already left their sums in
// `buckets` (host scalars arrive in pieces, engine.cuh) -- start from the stored sum, skip empty slices.  (A
// compile-time flag: the one-chunk kernel is exactly the plain walk.)
// Launch shape of the hot kernel.  Shipped values: 64 threads per block; G1 six blocks per SM (166 registers, no spill),
// G2 left to ptxas (255 registers = 4 blocks of 64).  Measured on B200 at 2^24 (profiles/r02_run2_lab_g1_acc_launch_shape_2p24.txt):
// 128 threads 76.1 ms, 128 x 4 blocks (128 registers) 75.3, 64 x 6 blocks 73.1, 64 x 8 blocks (128 registers) 75.8 -- the
// smaller block lets the length-sorted items of a block finish closer together.  G2 at 2^20: 16.6 ms either way, and
// forcing three blocks of 128 (168 registers, 1.8 KB of spills) costs 22.7 ms.  An L2 prefetch of the next entry's point
// changed nothing (73.2 vs 73.1 ms, profiles/r02_run7_lab_prefetch.txt).  tools/lab_build.py builds A/B variants of
// the library with other values (-DG16_ACC_BLOCK=.. -DG16_ACC_MIN_BLOCKS_G1=.. -DG16_ACC_MIN_BLOCKS_G2=..) for
// tools/bench_stages.py --lib.
#ifndef G16_ACC_BLOCK
#define G16_ACC_BLOCK 64
#endif
#ifndef G16_ACC_MIN_BLOCKS_G1
#define G16_ACC_MIN_BLOCKS_G1 6
#endif
#ifndef G16_ACC_MIN_BLOCKS_G2
#define G16_ACC_MIN_BLOCKS_G2 1
#endif
template <class F, bool ADD_TO>
struct BucketAccumulate {
    static constexpr int BLOCK = G16_ACC_BLOCK;
    static constexpr int MIN_BLOCKS = F::N == 12 ? G16_ACC_MIN_BLOCKS_G1 : G16_ACC_MIN_BLOCKS_G2;
    G16_HD static void run(size_t t, const uint32_t *pts, const uint32_t *entries, const WorkItem *items,
                           const uint32_t *n_items, uint32_t *buckets, uint32_t *chunk_out) {
        if (t >= *n_items) return;   // the launch covers an upper bound; the exact count lives on the device
        WorkItem it = items[t];
        XYZZ<F> acc = XYZZ<F>::inf();
        if (ADD_TO && !(it.bucket & SPLIT_FLAG)) {
            if (it.begin == it.end) return;
            acc = load_xyzz<F>(buckets, it.bucket);
        }
        for (uint32_t e = it.begin; e < it.end; ++e) {
            uint32_t v = entries[e];
            Affine<F> p = load_affine<F>(pts, v & 0x7fffffffu);
            if (v >> 31) p.y = F::neg(p.y);
            xyzz_madd(acc, p.x, p.y);
        }
        if (it.bucket & SPLIT_FLAG) store_xyzz<F>(chunk_out, t, acc);
        else store_xyzz<F>(buckets, it.bucket, acc);
    }
};

constexpr uint32_t MERGE_SERIAL_MAX = 16;   // buckets with more chunks than this are folded by a whole block
// Split buckets get the sum of their chunk partials.  GPU: one block per split bucket (grid-stride over
// the split list): every thread folds a strided subset of the chunks, then a shared-memory tree.
#if !defined(G16_EMU) && defined(__CUDACC__)
constexpr int MERGE_THREADS = 64;
template <class F>
__global__ void __launch_bounds__(MERGE_THREADS) chunk_merge_kernel(const uint32_t *split_list, const uint32_t *chunk_out,
                                                                    uint32_t *buckets, uint32_t add_to) {
    extern __shared__ uint32_t sm[];
    const int T = MERGE_THREADS, j = threadIdx.x;
    const uint32_t n_split = split_list[0];
    for (uint32_t k = blockIdx.x; k < n_split; k += gridDim.x) {
        uint32_t g = split_list[1 + 3 * k], first = split_list[2 + 3 * k], nch = split_list[3 + 3 * k];
        if (nch <= MERGE_SERIAL_MAX) continue;   // folded by ChunkMergeSerial (block-uniform branch)
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t c = j; c < nch; c += T) {
            XYZZ<F> p = load_xyzz<F>(chunk_out, first + c);
            xyzz_add_call(acc, p);
        }
        for (int d = T >> 1; d >= 1; d >>= 1) {
            const uint32_t *s = reinterpret_cast<const uint32_t *>(&acc);
#pragma unroll
            for (int w = 0; w < 4 * F::N; ++w) sm[w * T + j] = s[w];
            __syncthreads();
            if (j < d) {
                XYZZ<F> q;
                uint32_t *qd = reinterpret_cast<uint32_t *>(&q);
#pragma unroll
                for (int w = 0; w < 4 * F::N; ++w) qd[w] = sm[w * T + j + d];
                xyzz_add_call(acc, q);
            }
            __syncthreads();
        }
        if (j == 0) {
            if (add_to) {
                XYZZ<F> prev = load_xyzz<F>(buckets, g);
                xyzz_add_call(acc, prev);
            }
            store_xyzz<F>(buckets, g, acc);
        }
    }
}
#endif
// one thread per entry of the split list: buckets cut into at most `serial_max` chunks (all of them in the
// host emulation build, where serial_max = ~0)
template <class F>
struct ChunkMergeSerial {
    static constexpr int BLOCK = 64;
    G16_HD static void run(size_t k, const uint32_t *split_list, const uint32_t *chunk_out, uint32_t serial_max,
                           uint32_t *buckets, uint32_t add_to) {
        if (k >= split_list[0]) return;
        uint32_t g = split_list[1 + 3 * k], first = split_list[2 + 3 * k], nch = split_list[3 + 3 * k];
        if (nch > serial_max) return;
        XYZZ<F> acc = add_to ? load_xyzz<F>(buckets, g) : XYZZ<F>::inf();
        for (uint32_t c = 0; c < nch; ++c) {
            XYZZ<F> p = load_xyzz<F>(chunk_out, first + c);
            xyzz_add_call(acc, p);
        }
        store_xyzz<F>(buckets, g, acc);
    }
};

// One level of the parallel bucket reduction.  For every window, the level maps arrays
//   X[0..n_in) (to be weighted by index) and Y[0..n_in) (already weighted partial sums)
// to arrays of length n_out = ceil(n_in / L):
//   X'[g] = sum_j X[gL + j]
//   Y'[g] = sum_j Y[gL + j] + 2^shift * sum_j j * X[gL + j]
// With shift = log2(L) * level this telescopes to  sum_i i * X0[i] = Y_final[0]
// (derivation in DESIGN.md "Bucket reduction").
#ifndef G16_RED_TWO_LOOPS_G2
#define G16_RED_TWO_LOOPS_G2 1   // see ReduceLevel::run; 0 builds the one-loop form for A/B runs
#endif
#ifndef G16_RED_TWO_LOOPS_G1
#define G16_RED_TWO_LOOPS_G1 0   // G1 keeps one loop: 0.97 ms against 1.00 (two loops, 216 registers) / 1.06 (two loops at 168
#endif                           // registers, 6 blocks per SM) at 2^19 buckets, 2.96 against 3.24 ms at 2^21 (profiles/r02_run34_*)
#ifndef G16_RED_MIN_BLOCKS_G1
#define G16_RED_MIN_BLOCKS_G1 1
#endif
#ifndef G16_RED_MIN_BLOCKS_G2
#define G16_RED_MIN_BLOCKS_G2 1   // left to ptxas (255 registers); A/B builds: tools/lab_build.py
#endif
template <class F>
struct ReduceLevel {
    static constexpr int BLOCK = 64;
    static constexpr int MIN_BLOCKS = F::N == 12 ? G16_RED_MIN_BLOCKS_G1 : G16_RED_MIN_BLOCKS_G2;
    static constexpr bool TWO_LOOPS = F::N == 24 ? G16_RED_TWO_LOOPS_G2 : G16_RED_TWO_LOOPS_G1;
    G16_HD static void run(size_t t, const uint32_t *X, const uint32_t *Y, uint32_t n_in, uint32_t n_out, uint32_t L,
                           uint32_t shift, uint32_t *Xo, uint32_t *Yo) {
        uint32_t w = (uint32_t)(t / n_out), g = (uint32_t)(t % n_out);
        size_t base = (size_t)w * n_in;
        uint32_t lo = g * L;
        uint32_t hi = lo + L < n_in ? lo + L : n_in;
        XYZZ<F> running = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
        if (TWO_LOOPS) {
            // G2: three projective points alive at once (x, running, acc = 288 words) is what makes this level spill;
            // so first the suffix sums alone, parked IN PLACE of the entries they replace (the input of a level is
            // dead after it: bucket array or a workspace buffer of the previous level), then their plain sum.  Same
            // additions in the same order; 768 bytes of extra traffic per entry.
            uint32_t *Xm = const_cast<uint32_t *>(X);
            for (uint32_t i = hi; i-- > lo + 1;) {
                XYZZ<F> x = load_xyzz<F>(X, base + i);
                xyzz_add(running, x);
                store_xyzz<F>(Xm, base + i, running);
            }
            {
                XYZZ<F> x0 = load_xyzz<F>(X, base + lo);
                xyzz_add(running, x0);
            }
            store_xyzz<F>(Xo, (size_t)w * n_out + g, running);
            for (uint32_t i = hi; i-- > lo + 1;) {
                XYZZ<F> sfx = load_xyzz<F>(X, base + i);
                xyzz_add(acc, sfx);
            }
        } else {
            for (uint32_t i = hi; i-- > lo + 1;) {
                XYZZ<F> x = load_xyzz<F>(X, base + i);
                xyzz_add(running, x);
                xyzz_add(acc, running);
            }
            {
                XYZZ<F> x0 = load_xyzz<F>(X, base + lo);
                xyzz_add(running, x0);
            }
        }
        for (uint32_t s = 0; s < shift; ++s) xyzz_dbl(acc);
        if (Y) {
            for (uint32_t i = lo; i < hi; ++i) {
                XYZZ<F> y = load_xyzz<F>(Y, base + i);
                xyzz_add(acc, y);
            }
        }
        if (!TWO_LOOPS) store_xyzz<F>(Xo, (size_t)w * n_out + g, running);
        store_xyzz<F>(Yo, (size_t)w * n_out + g, acc);
    }
};

// (A variant of this level with one quad of lanes per group -- every addition the 4-lane cooperative one of quad.cuh --
// was measured for G2 and lost: bucket reduction 5.79 ms against 5.31 ms at 2^19 buckets,
// profiles/r02_run3_lab_g2_quad_reduce_and_window_sweep_2p20.txt.  Not in the library.)

// Block-cooperative level of the same reduction for the upper, latency-bound part of the tree.  From here
// on a level carries THREE arrays: X (to be weighted by index), Y1 (weighted partials produced by the X
// blocks of the previous level) and Y2 (plain sums of the previous level's Y1 + Y2); the total is
// sum Y1 + sum Y2 + 2^shift * sum_i i X_i.  Splitting Y keeps the two halves of a level independent:
//   X blocks (blockIdx.z == 0): every element (a quad of lanes, see quad.cuh) folds TILE_K consecutive
//       entries, a suffix scan over the block's elements gives P_j = sum_{k >= j} S_k, so that
//       X' = P_0 and sum_j j X_j = sum_j lw_j + TILE_K sum_{j >= 1} P_j (one tree), then `shift` doublings;
//   Y blocks (blockIdx.z == 1): plain sum of the tile's Y1 and Y2 entries (fold + tree).
// Every addition is the 4-lane cooperative one: a level costs ~2 log2(T) + 2 TILE_K dependent additions of
// 4 multiplication latencies each.  Shared memory holds one XYZZ per element, element-major with an odd
// stride (conflict free for the quad-broadcast reads).
constexpr int TILE_K = 4;        // consecutive entries folded serially by each element before the block scan
constexpr int TILE_ELEMS = 64;   // elements (quads) per block -> 256 threads, TILE_K * TILE_ELEMS entries per tile
// Per group: the same 256 entries per tile as K entries x 256 / K elements.  G2 runs 8 x 32: blocks of 128 threads may
// use 254 registers (two per SM) where 256 threads were held to 128, and a tile needs a third fewer cooperative
// additions (fold 15 x 32 + scan / tree 10 x 32 against 7 x 64 + 12 x 64): bucket reduction 4.53 -> 4.28 ms at 2^19 buckets
// (profiles/r02_run36_lab_g2_tile_k8.txt).  (G16_TILE_K_G1 / _G2: A/B builds, tools/lab_build.py)
#ifndef G16_TILE_K_G2
#define G16_TILE_K_G2 8
#endif
#ifndef G16_TILE_K_G1
#define G16_TILE_K_G1 4
#endif
template <class F>
struct TileShape {
    static constexpr int K = F::N == 24 ? G16_TILE_K_G2 : G16_TILE_K_G1;
    static constexpr int ELEMS = TILE_K * TILE_ELEMS / K;
};
#if !defined(G16_EMU) && defined(__CUDACC__)
template <class F>
__device__ __forceinline__ void tile_put(uint32_t *sm, int e, int q, const XYZZ<F> &p) {
    constexpr int W = 4 * F::N, QW = W / 4;
    const uint32_t *s = reinterpret_cast<const uint32_t *>(&p);
    uint32_t *d = sm + (size_t)e * (W + 1);
#pragma unroll
    for (int k = 0; k < W; ++k)
        if (k / QW == q) d[k] = s[k];   // every lane of the quad writes its quarter
}
template <class F>
__device__ __forceinline__ XYZZ<F> tile_get(const uint32_t *sm, int e) {
    constexpr int W = 4 * F::N;
    XYZZ<F> p;
    uint32_t *d = reinterpret_cast<uint32_t *>(&p);
    const uint32_t *s = sm + (size_t)e * (W + 1);
#pragma unroll
    for (int k = 0; k < W; ++k) d[k] = s[k];
    return p;
}
template <class F>
__global__ void __launch_bounds__(4 * TileShape<F>::ELEMS, 2) tile_reduce_kernel(const uint32_t *X, const uint32_t *Y1, const uint32_t *Y2,
                                                                     uint32_t n_in, uint32_t n_out, uint32_t tile_entries,
                                                                     uint32_t shift, uint32_t *Xo, uint32_t *Y1o, uint32_t *Y2o) {
    extern __shared__ uint32_t sm[];
    constexpr int TILE_K = TileShape<F>::K;   // (shadows the global default)
    const int T = blockDim.x >> 2;                 // elements in this block; T * TILE_K >= tile_entries
    const int e = threadIdx.x >> 2, q = threadIdx.x & 3;
    const int warp_e0 = (threadIdx.x & ~31) >> 2;  // first element of this warp (8 elements per warp)
    const uint32_t w = blockIdx.y, g = blockIdx.x;
    const size_t base = (size_t)w * n_in;
    const uint32_t local = (uint32_t)e * TILE_K;   // offset of this element's entries inside the tile
    const uint32_t first = g * tile_entries + local;
    if (blockIdx.z == 1) {
        // ---- plain sum of the already weighted partials
        XYZZ<F> y = XYZZ<F>::inf();
#pragma unroll 1
        for (int k = 0; k < TILE_K; ++k) {
            uint32_t i = first + k;
            bool live = local + k < tile_entries && i < n_in;
            if (Y1) { XYZZ<F> t = live ? load_xyzz<F>(Y1, base + i) : XYZZ<F>::inf(); xyzz_add_quad(y, t, q); }
            if (Y2) { XYZZ<F> t = live ? load_xyzz<F>(Y2, base + i) : XYZZ<F>::inf(); xyzz_add_quad(y, t, q); }
        }
#pragma unroll 1
        for (int d = T >> 1; d >= 1; d >>= 1) {
            tile_put<F>(sm, e, q, y);
            __syncthreads();
            if (warp_e0 < d) {
                XYZZ<F> t = e < d ? tile_get<F>(sm, e + d) : XYZZ<F>::inf();
                xyzz_add_quad(y, t, q);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) store_xyzz<F>(Y2o, (size_t)w * n_out + g, y);
        return;
    }
    // ---- X block.  Per-element fold of TILE_K entries: p = their sum, lw = sum_k k * x_k (local weights)
    XYZZ<F> p = XYZZ<F>::inf(), lw = XYZZ<F>::inf();
#pragma unroll 1
    for (int k = TILE_K - 1; k >= 0; --k) {
        uint32_t i = first + k;
        bool live = local + k < tile_entries && i < n_in;
        XYZZ<F> x = live ? load_xyzz<F>(X, base + i) : XYZZ<F>::inf();
        xyzz_add_quad(p, x, q);
        if (k >= 1) xyzz_add_quad(lw, p, q);   // after the loop: lw = sum_{k>=1} (suffix sum from k) = sum_k k x_k
    }
    // suffix scan of the element sums (Hillis-Steele): p = P_e = sum_{t >= e} S_t
#pragma unroll 1
    for (int d = 1; d < T; d <<= 1) {
        tile_put<F>(sm, e, q, p);
        __syncthreads();
        XYZZ<F> t = e + d < T ? tile_get<F>(sm, e + d) : XYZZ<F>::inf();
        xyzz_add_quad(p, t, q);
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz<F>(Xo, (size_t)w * n_out + g, p);
    // sum_i i x_i over the tile = sum_e lw_e + TILE_K * sum_{e >= 1} P_e
    XYZZ<F> v = e >= 1 ? p : XYZZ<F>::inf();
#pragma unroll 1
    for (int k = 1; k < TILE_K; k <<= 1) xyzz_dbl_quad(v, q);
    xyzz_add_quad(v, lw, q);
#pragma unroll 1
    for (int d = T >> 1; d >= 1; d >>= 1) {
        tile_put<F>(sm, e, q, v);
        __syncthreads();
        if (warp_e0 < d) {
            XYZZ<F> t = e < d ? tile_get<F>(sm, e + d) : XYZZ<F>::inf();
            xyzz_add_quad(v, t, q);
        }
        __syncthreads();
    }
    if (threadIdx.x < 32) {
#pragma unroll 1
        for (uint32_t s = 0; s < shift; ++s) xyzz_dbl_quad(v, q);
        if (threadIdx.x == 0) store_xyzz<F>(Y1o, (size_t)w * n_out + g, v);
    }
}
#endif
// Serial statement of one tile (host emulation build; also documents what the kernel above computes).
template <class F>
struct TileReduceSerial {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t t, const uint32_t *X, const uint32_t *Y1, const uint32_t *Y2, uint32_t n_in, uint32_t n_out,
                           uint32_t T, uint32_t shift, uint32_t *Xo, uint32_t *Y1o, uint32_t *Y2o) {
        uint32_t w = (uint32_t)(t / n_out), g = (uint32_t)(t % n_out);
        size_t base = (size_t)w * n_in;
        uint32_t lo = g * T, hi = lo + T < n_in ? lo + T : n_in;
        XYZZ<F> running = XYZZ<F>::inf(), acc = XYZZ<F>::inf(), ysum = XYZZ<F>::inf();
        for (uint32_t i = hi; i-- > lo + 1;) {
            XYZZ<F> x = load_xyzz<F>(X, base + i);
            xyzz_add_call(running, x);
            xyzz_add_call(acc, running);
        }
        XYZZ<F> x0 = load_xyzz<F>(X, base + lo);
        xyzz_add_call(running, x0);
        for (uint32_t s = 0; s < shift; ++s) xyzz_dbl_call(acc);
        if (Y1) for (uint32_t i = lo; i < hi; ++i) { XYZZ<F> y = load_xyzz<F>(Y1, base + i); xyzz_add_call(ysum, y); }
        if (Y2) for (uint32_t i = lo; i < hi; ++i) { XYZZ<F> y = load_xyzz<F>(Y2, base + i); xyzz_add_call(ysum, y); }
        store_xyzz<F>(Xo, (size_t)w * n_out + g, running);
        store_xyzz<F>(Y1o, (size_t)w * n_out + g, acc);
        if (Y1 || Y2) store_xyzz<F>(Y2o, (size_t)w * n_out + g, ysum);
    }
};

// Final fold: window sum S_w = X[w] + Y[w] (bucket b carries weight b + 1), then
// result = sum_w 2^(c w) S_w by Horner, optionally + `extra` partial sums, then to affine.
// out_xyzz (4*F::N words) receives the projective result; out_aff (2*F::N words + flag word).
template <class F>
struct WindowCombine {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *X, const uint32_t *Y, const uint32_t *Y2, uint32_t nwin, uint32_t c,
                           uint32_t *out_xyzz, uint32_t *out_aff) {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t w = nwin; w-- > 0;) {
            if (!acc.is_inf()) for (uint32_t s = 0; s < c; ++s) xyzz_dbl_call(acc);
            XYZZ<F> x = load_xyzz<F>(X, w);
            xyzz_add_call(acc, x);
            if (Y) {
                XYZZ<F> y = load_xyzz<F>(Y, w);
                xyzz_add_call(acc, y);
            }
            if (Y2) {
                XYZZ<F> y = load_xyzz<F>(Y2, w);
                xyzz_add_call(acc, y);
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

// out[j] = scalar_j * P_j for a handful of points (prove: s * pi_A and r * pi_B', crates/groth16-core/src/lib.rs:235,259),
// one thread per point, double-and-add from the top bit with mixed additions.  scalars: Montgomery Fr (8 words each);
// aff: affine result records as WindowCombine / PartialCombine write them (x, y, infinity word), `stride` words apart.
// ~255 doublings + ~128 mixed additions in a chain: 2 ms of latency, which the prove schedule hides under the bucket
// accumulation of the MSMs that are still running.
template <class F>
struct ScalarMulAffine {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t j, const uint32_t *scalars, const uint32_t *aff, uint32_t stride, uint32_t *out_xyzz) {
        uint32_t k[8];
        load_scalar(scalars, j, true, k);
        const uint32_t *src = aff + j * stride;
        Affine<F> p;
        uint32_t *dx = limbs(p.x), *dy = limbs(p.y);
        for (int w = 0; w < F::N; ++w) { dx[w] = src[w]; dy[w] = src[F::N + w]; }
        XYZZ<F> acc = XYZZ<F>::inf();
        if (!src[2 * F::N])
            for (int bit = 254; bit >= 0; --bit) {
                xyzz_dbl_call(acc);
                if ((k[bit >> 5] >> (bit & 31)) & 1u) xyzz_madd_call(acc, p.x, p.y);
            }
        store_xyzz<F>(out_xyzz, j, acc);
    }
};

// The same on one warp with the 4-lane cooperative group law of quad.cuh (device build): quad j runs the chain of point
// j (up to 8 points), two scalar bits per step from a table {P, 2P, 3P}: 128 steps of two doublings and one addition of
// ~5.6 us each -- a thread on its own needs ~12 us per doubling or addition (measured: 445 operations in 5.3 ms,
// profiles/r02_run15_prove_timeline_serial_chains.txt).
#if !defined(G16_EMU) && defined(__CUDACC__)
template <class F>
__global__ void __launch_bounds__(32) scalar_mul_quad_kernel(uint32_t n, const uint32_t *scalars, const uint32_t *aff, uint32_t stride,
                                                             uint32_t *out_xyzz) {
    const uint32_t j = threadIdx.x >> 2;
    const int q = threadIdx.x & 3;
    const bool live = j < n;
    uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    XYZZ<F> tab[3];
    tab[0] = XYZZ<F>::inf();
    if (live) {
        load_scalar(scalars, j, true, k);
        const uint32_t *src = aff + j * stride;
        if (!src[2 * F::N]) {
            uint32_t *dx = limbs(tab[0].x), *dy = limbs(tab[0].y);
            for (int w = 0; w < F::N; ++w) { dx[w] = src[w]; dy[w] = src[F::N + w]; }
            tab[0].zz = F::one(); tab[0].zzz = F::one();
        }
    }
    tab[1] = tab[0]; xyzz_dbl_quad(tab[1], q);                      // 2P
    tab[2] = tab[1]; xyzz_add_quad(tab[2], tab[0], q);              // 3P
    XYZZ<F> acc = XYZZ<F>::inf();
#pragma unroll 1
    for (int bit = 254; bit >= 0; bit -= 2) {                       // pairs (255, 254), (253, 252), ... ; bit 255 is zero
        xyzz_dbl_quad(acc, q);
        xyzz_dbl_quad(acc, q);
        uint32_t hi = bit + 1 < 256 ? (k[(bit + 1) >> 5] >> ((bit + 1) & 31)) & 1u : 0u;
        uint32_t d = ((k[bit >> 5] >> (bit & 31)) & 1u) | (hi << 1);
        XYZZ<F> t = XYZZ<F>::inf();
        if (d) t = tab[d - 1];
        xyzz_add_quad(acc, t, q);                                   // all quads step together; a zero digit adds infinity
    }
    if (live && q == 0) store_xyzz<F>(out_xyzz, j, acc);
}
#endif

// Sum of k projective partial results (multi-GPU combine), then to affine.
template <class F>
struct PartialCombine {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *partials, uint32_t k, uint32_t *out_xyzz, uint32_t *out_aff) {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t i = 0; i < k; ++i) {
            XYZZ<F> p = load_xyzz<F>(partials, i);
            xyzz_add_call(acc, p);
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

// The same fold on one warp (device build): eight quads take the partials round-robin with the 4-lane cooperative
// addition, a three-step tree through shared memory joins them, lane 0 converts to affine.  k = 8 (one box of GPUs):
// 1 + 3 cooperative additions instead of 8 serial ones in front of the inversion.
#if !defined(G16_EMU) && defined(__CUDACC__)
template <class F>
__global__ void __launch_bounds__(32) partial_combine_warp_kernel(const uint32_t *partials, uint32_t k, uint32_t *out_xyzz,
                                                                  uint32_t *out_aff) {
    extern __shared__ uint32_t sm[];
    const int e = threadIdx.x >> 2, q = threadIdx.x & 3;
    XYZZ<F> acc = XYZZ<F>::inf();
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < k; i0 += 8) {
        XYZZ<F> p = i0 + e < k ? load_xyzz<F>(partials, i0 + e) : XYZZ<F>::inf();
        xyzz_add_quad(acc, p, q);
    }
#pragma unroll 1
    for (int d = 4; d >= 1; d >>= 1) {
        tile_put<F>(sm, e, q, acc);
        __syncwarp();
        XYZZ<F> t = e < d ? tile_get<F>(sm, e + d) : XYZZ<F>::inf();
        xyzz_add_quad(acc, t, q);
        __syncwarp();
    }
    if (threadIdx.x == 0) {
        if (out_xyzz) store_xyzz<F>(out_xyzz, 0, acc);
        if (out_aff) {
            Affine<F> a = xyzz_to_affine(acc);
            const uint32_t *s = reinterpret_cast<const uint32_t *>(&a);
            for (int j = 0; j < 2 * F::N; ++j) out_aff[j] = s[j];
            out_aff[2 * F::N] = acc.is_inf() ? 1u : 0u;
        }
    }
}
#endif

// Precomputed multiples for resident bases: table[w * n + i] = 2^(c w) * P_i as affine points, so that
// digit w of scalar i adds table[w][i] into ONE shared bucket set -- the Horner fold over windows and
// (nwin - 1) of the nwin bucket reductions disappear.  One thread per base point: nwin - 1 runs of c
// doublings, then one shared inversion (Montgomery's trick) to bring all multiples back to affine.
constexpr uint32_t PRE_MAX_WIN = 32;
template <class F>
struct PrecomputeBases {
    static constexpr int BLOCK = 64;
    G16_HD static void run(size_t i, const uint32_t *pts, size_t n, uint32_t c, uint32_t nwin, uint32_t *table) {
        Affine<F> p = load_affine<F>(pts, i);
        store_affine_pt<F>(table, i, p);
        if (p.is_inf()) {
            for (uint32_t w = 1; w < nwin; ++w) store_affine_pt<F>(table, (size_t)w * n + i, p);
            return;
        }
        XYZZ<F> m[PRE_MAX_WIN - 1];
        F prefix[PRE_MAX_WIN - 1];
        XYZZ<F> acc = XYZZ<F>::from_affine(p);
        F run = F::one();
        for (uint32_t w = 1; w < nwin; ++w) {
            for (uint32_t s = 0; s < c; ++s) xyzz_dbl(acc);
            m[w - 1] = acc;
            prefix[w - 1] = run;            // product of zzz of the earlier multiples
            run = F::mul(run, acc.zzz);
        }
        F inv_all = F::inv(run);
        for (uint32_t w = nwin - 1; w >= 1; --w) {
            const XYZZ<F> &q = m[w - 1];
            F a = F::mul(inv_all, prefix[w - 1]);   // 1 / zzz_w = Z^-3
            inv_all = F::mul(inv_all, q.zzz);
            F zi = F::mul(a, q.zz);                 // Z^-1
            Affine<F> r;
            r.x = F::mul(q.x, F::sqr(zi));
            r.y = F::mul(q.y, a);
            store_affine_pt<F>(table, (size_t)w * n + i, r);
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
