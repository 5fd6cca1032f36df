// Host-side engine: contexts, resident bases, the MSM pipeline and the fixed-base pipeline.
// Everything here only allocates, copies and launches; all arithmetic runs in the kernels.
#pragma once
#include <algorithm>
#include <memory>
#include <vector>
#include "kernel_api.cuh"

namespace g16 {

enum Group { GROUP_G1 = 1, GROUP_G2 = 2 };

template <class F> struct GroupOf { static constexpr int id = FieldWords<F>::group; };

struct Workspace {
    DevBuf scalars, counts, codes, ranks, bins, items, item_start, chunk_out, cursor, entries, buckets, scatter_stage, red[6], scan_tmp, out, partials, staging;
    DevBuf prove_w, prove_misc, ntt_abc, ntt_tw, ntt_consts, ntt_out;
    uint32_t ntt_log_n = 0xffffffffu;   // size the cached twiddles / constants were built for (none yet)
    DevBuf fb_base, fb_powers, fb_table[3], fb_out, fb_flags;
    std::vector<uint32_t> fb_table_key[3];  // base limbs the cached table was built for
    void release() {
        scalars.release(); counts.release(); codes.release(); ranks.release(); bins.release(); items.release(); item_start.release(); chunk_out.release(); cursor.release(); entries.release(); buckets.release(); scatter_stage.release();
        for (auto &r : red) r.release();
        scan_tmp.release(); out.release(); partials.release(); staging.release();
        prove_w.release(); prove_misc.release(); ntt_abc.release(); ntt_tw.release(); ntt_consts.release(); ntt_out.release(); ntt_log_n = 0xffffffffu; fb_base.release(); fb_powers.release(); fb_out.release(); fb_flags.release();
        for (auto &t : fb_table) t.release();
    }
};

// optional per-stage CUDA-event timing of the last MSM on a device (bench.py reads it to put the
// dominant kernel's measured duration into the roofline line)
constexpr int N_STAGES = 6;  // digits, scan+items, scatter, accumulate, reduce, combine
struct StageTimer {
    bool enabled = false;
#ifndef G16_EMU
    cudaEvent_t ev[N_STAGES + 1] = {};
    bool created = false;
#endif
    bool valid = false;
    void mark(int i, stream_t s) {
#ifndef G16_EMU
        if (!enabled) return;
        if (!created) {
            for (auto &e : ev) G16_CUDA_CHECK(cudaEventCreate(&e));
            created = true;
        }
        G16_CUDA_CHECK(cudaEventRecord(ev[i], s));
        if (i == N_STAGES) valid = true;
#else
        (void)i; (void)s;
#endif
    }
    // elapsed ms from `epoch` (an event recorded earlier on any stream of the device) to every mark
    bool read_since(void *epoch, float *t) {
#ifndef G16_EMU
        if (!enabled || !valid || !epoch) return false;
        for (int i = 0; i <= N_STAGES; ++i) G16_CUDA_CHECK(cudaEventElapsedTime(&t[i], (cudaEvent_t)epoch, ev[i]));
        return true;
#else
        (void)epoch; (void)t; return false;
#endif
    }
    // call after the stream has been synchronised
    bool read(float *ms) {
#ifndef G16_EMU
        if (!enabled || !valid) return false;
        for (int i = 0; i < N_STAGES; ++i) G16_CUDA_CHECK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
        return true;
#else
        (void)ms; return false;
#endif
    }
    void destroy() {
#ifndef G16_EMU
        if (created) for (auto &e : ev) cudaEventDestroy(e);
        created = false;
#endif
    }
};

struct Device {
    int id = 0;
    uint32_t sm_count = 1;   // multiprocessors of this GPU (cudaDeviceProp), sizes the grid-stride launches
    stream_t stream = nullptr;
    bool own_stream = false;
    Workspace ws;
    StageTimer timer;
    MsmPlan last_plan{};
    uint32_t item_max_override = 0;   // g16_ctx_set_item_max: 0 = chosen per call from its size
    // extra lanes (own stream + workspace) on the same GPU: lets the latency-bound tail of one MSM
    // (reduction tree, inversion) overlap the bucket accumulation of another (prove schedule)
    std::vector<std::unique_ptr<Device>> extra;
};

inline void set_device(int id);
inline void set_device_nothrow(int id) {
#ifndef G16_EMU
    cudaSetDevice(id);
#else
    (void)id;
#endif
}
inline Device &lane_of(Device &dv, int k) {
    if (k == 0) return dv;
    while ((int)dv.extra.size() < k) {
        std::unique_ptr<Device> l(new Device);
        l->id = dv.id;
        l->sm_count = dv.sm_count;
#ifndef G16_EMU
        G16_CUDA_CHECK(cudaSetDevice(dv.id));
        G16_CUDA_CHECK(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking));
        l->own_stream = true;
#endif
        dv.extra.push_back(std::move(l));
    }
    return *dv.extra[k - 1];
}

inline void set_device(int id) {
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaSetDevice(id));
#else
    (void)id;
#endif
}
// Host-scalar MSMs of at least `h2d_pipe_min` scalars are cut into H2D_PIPE_PARTS index ranges of growing size
// (1/32, 5/32, 26/32 of the scalars).  The ranges are copied back to back on a second stream; each one is decomposed,
// sorted and accumulated INTO THE SAME bucket array as soon as it has arrived, and the buckets are reduced once at the
// end -- so only the first, small copy is exposed (0.3 ms of 9.7 ms at 2^24) and no range pays its own reduction.
#ifndef G16_EMU
constexpr size_t H2D_PIPE_MIN = (size_t)1 << 19;
#else
constexpr size_t H2D_PIPE_MIN = 64;   // the emulation tests walk the chunk logic on small inputs
#endif
constexpr size_t H2D_PIPE_PARTS = 3;
constexpr size_t H2D_PIPE_NUM[H2D_PIPE_PARTS + 1] = {0, 1, 6, 32};   // cumulative 32nds
constexpr int H2D_PIPE_LANE = 7;   // lanes 0-4: prove schedule

struct Context {
    std::vector<Device> devs;
    std::string err;
    void *prove_epoch = nullptr;   // cudaEvent_t recorded at the start of the last timed prove (test hook: lane timeline)
    unsigned c_override = 0;
    size_t h2d_pipe_min = H2D_PIPE_MIN;
};

struct BasesShard {
    int dev = 0;       // index into Context::devs
    int cuda_dev = 0;  // CUDA ordinal of that device (the destructor must not reach through the context)
    uint32_t *pts = nullptr;
    size_t begin = 0, n = 0;
    bool owned = true;
    // optional precomputed multiples: table[w * n + i] = 2^(pre_c w) P_i, w < ceil(256 / pre_c)
    uint32_t *table = nullptr;
    unsigned pre_c = 0;
};

struct Bases {
    Context *ctx = nullptr;   // identity check + launch-time lookup; never dereferenced by the destructor
    int group = 0;
    size_t n = 0;
    std::vector<BasesShard> shards;
    ~Bases() {
        for (auto &s : shards)
        {
            // destructors must not throw: select the device without the checking wrapper
            if ((s.owned && s.pts) || s.table) set_device_nothrow(s.cuda_dev);
            if (s.owned && s.pts) dev_free(s.pts);
            if (s.table) dev_free(s.table);
        }
    }
};

// ---------------------------------------------------------------------------------------
// Window size.  Cost model in field multiplications: every non-zero digit costs one mixed
// addition (10 M), every bucket costs two full additions in the reduction (2 x 14 M).
// ---------------------------------------------------------------------------------------
inline MsmPlan make_plan(size_t n, unsigned c_override, size_t point_words) {
    unsigned best_c = 1;
    double best = 1e300;
    unsigned lo = 2, hi = 22;
    // tiny inputs (verifier, ad-hoc prove terms): the cost is the serial fold over the windows, not the
    // additions -- use few, wide windows
    if (n < 1024) lo = hi = 8;
    if (c_override) lo = hi = std::min(std::max(c_override, 2u), 24u);
    for (unsigned c = lo; c <= hi; ++c) {
        double nwin = (256 + c - 1) / c;
        double nb = (double)(1ull << (c - 1));
        // bucket storage cap: 8 GiB
        if (!c_override && nwin * nb * 2.0 * point_words * 4.0 > 8.0 * 1073741824.0) break;
        double cost = nwin * ((double)n * 10.0 + nb * 30.0);
        if (cost < best) { best = cost; best_c = c; }
    }
    MsmPlan p;
    p.c = best_c;
    p.nwin = (256 + best_c - 1) / best_c;
    p.nb = 1u << (best_c - 1);
    p.total = p.nwin * p.nb;
    p.bwin = p.nwin;
    p.stride = 0;
    p.offset = 0;
    return p;
}

// plan for bases that carry precomputed multiples for window size c
inline MsmPlan make_shared_plan(unsigned c, size_t stride) {
    MsmPlan p;
    p.c = c;
    p.nwin = (256 + c - 1) / c;
    p.nb = 1u << (c - 1);
    p.total = p.nb;
    p.bwin = 1;
    p.stride = (uint32_t)stride;
    p.offset = 0;
    return p;
}

// window size for a precomputed table over n bases: every non-zero digit costs one mixed addition, the single shared
// bucket set two full additions per bucket (charged as three mixed additions: the reduction runs below the accumulation's
// efficiency), and the table (always ceil(256 / c) windows: any scalar stays valid) must fit `budget` bytes and 31 bits.
// Checked against sweeps on B200: 2^20 bases full width G2 c = 18 / 19 / 20: 22.2 / 20.8 / 19.6 ms; 64-bit scalars c = 16 /
// 17 / 18 / 20: G1 4.63 / 2.59 / 2.93 / 3.45 ms, G2 13.8 / 7.6 / 8.3 / 10.8 ms (profiles/r02_run27_lab_window_choice.txt).
// scalar_bits < 255: the caller knows its scalars are that short (the reference truncates the witness and the H
// coefficients to 64 bits, crates/groth16-core/src/lib.rs:156-161,203-208), so only ceil((bits + 1) / c) windows hold
// non-zero digits (the + 1: the carry of the signed recoding -- c = 16 puts half of all 64-bit scalars into a fifth
// window) and a smaller c -- fewer buckets to reduce -- wins.
inline unsigned choose_precompute_c(size_t n, size_t point_bytes, size_t budget, unsigned scalar_bits = 255) {
    unsigned best_c = 0;
    double best = 1e300;
    const double bucket_cost = 30.0;
    const unsigned bits = scalar_bits == 0 || scalar_bits > 255 ? 255 : scalar_bits;
    for (unsigned c = 8; c <= 24; ++c) {
        double nwin = (256 + c - 1) / c;
        if (nwin * (double)n * point_bytes > (double)budget) continue;
        if (nwin * (double)n >= 2147483648.0) continue;
        double used = std::min(nwin, (double)((bits + 1 + c - 1) / c));
        double cost = used * (double)n * 10.0 + (double)(1ull << (c - 1)) * bucket_cost;
        if (cost < best) { best = cost; best_c = c; }
    }
    return best_c;
}

constexpr size_t SCATTER_TWO_PASS_BYTES = (size_t)96 << 20;   // `entries` larger than this (~L2) are scattered in two passes
constexpr uint32_t REDUCE_LOG_L = 5;
// longest work item: 256 additions when the call is large (a serial walk of 256 is noise), shorter when it is small
// and the longest item would set the kernel's duration: max(ITEM_FLOOR, entries >> ITEM_SHIFT).  The floor only acts
// below ~2^17 pairs: at 2^16 pairs the 2^15 buckets (32 entries each) alone are too few threads, and slices of <= 11
// entries -- three per bucket, 1.7 waves of work items -- beat slices of 16 (accumulate 0.65 -> 0.60 ms for G1, 2.07 ->
// 1.68 ms for G2; sweep over 6 .. 128 in profiles/r02_run30_sweep_item_max.jsonl, results bit-identical throughout).
// g16_ctx_set_item_max overrides the choice per context.
#ifndef G16_ITEM_FLOOR
#define G16_ITEM_FLOOR 11
#endif
constexpr size_t ITEM_FLOOR = G16_ITEM_FLOOR;   // (tools/lab_build.py overrides the G16_* constants for A/B builds)
constexpr unsigned ITEM_SHIFT = 17;
// Levels with at most 2^tile_max_log2 entries run block-cooperatively; a thread level aims to leave 2^groups_log2 groups
// behind.  Measured on B200 (profiles/README.md run 19): 15 / 15 is best up to 2^20 buckets (reduce 1.14 -> 0.99 ms at 2^19
// buckets), 17 / 16 beyond (3.07 -> 3.00 ms at 2^21).
#ifndef G16_RED_GROUPS_LOG2
#define G16_RED_GROUPS_LOG2 15
#endif
#ifndef G16_RED_TILE_MAX_LOG2
#define G16_RED_TILE_MAX_LOG2 15
#endif
inline void reduce_split(size_t buckets, size_t &tile_level_max, size_t &thread_level_groups) {
    bool big = buckets > ((size_t)1 << 20);
    thread_level_groups = (size_t)1 << (big ? 17 : G16_RED_GROUPS_LOG2);
    tile_level_max = (size_t)1 << (big ? 16 : G16_RED_TILE_MAX_LOG2);
}

// One MSM on one lane as three steps, so that a schedule of several MSMs (prove) can order them: `front` = stages 1-3
// (digits, bucket offsets + work items, counting sort) for a range of the scalars, `accumulate` = stage 4 for that
// range (the hot kernel), `back` = stages 5-6 (bucket reduction, fold, to affine).  All asynchronous on dv.stream.
template <class F>
struct MsmJob {
    Device &dv;
    const uint32_t *pts = nullptr;   // resident bases (or their table of multiples)
    MsmPlan plan{};
    size_t n = 0, first = 0, n_max = 0, total = 0, nbins = 0;
    uint32_t *counts = nullptr, *codes = nullptr, *ranks = nullptr, *scan_tmp = nullptr, *bins = nullptr, *bin_cursor = nullptr;
    uint32_t *split_list = nullptr, *entries = nullptr, *staging = nullptr, *part_cursor = nullptr, *buckets = nullptr, *chunk_out = nullptr;
    WorkItem *items = nullptr;
    uint32_t item_max = 0;      // of the range `front` handled last
    size_t item_min_ = 1;       // shortest item limit the workspaces were sized for
    size_t n_entries = 0;

    // bases [first_, first_ + n_) of the shard; ranges of at most n_max_ scalars will be fed (0 = the whole call at once).
    // Takes every workspace buffer before anything is queued (growing a buffer frees the old one).
    MsmJob(Device &dv_, const BasesShard &sh, size_t n_, unsigned c_override, size_t first_, size_t n_max_ = 0) : dv(dv_) {
        n = n_; first = first_; n_max = n_max_ ? n_max_ : n_;
        pts = sh.table ? sh.table : sh.pts;
        if (n >= (1ull << 31)) throw Error{G16_ERR_INVALID, "MSM length must be < 2^31"};
        plan = sh.table ? make_shared_plan(sh.pre_c, sh.n) : make_plan(n, c_override, 2 * FieldWords<F>::N);
        total = plan.total;
        if ((double)n * plan.nwin >= 4294967295.0) throw Error{G16_ERR_INVALID, "n * windows exceeds 2^32 entries"};
        plan.offset = (uint32_t)first;
        dv.last_plan = plan;
        Workspace &ws = dv.ws;
        const size_t max_entries = n_max * plan.nwin;
        nbins = k_item_bins();
        counts = ws.counts.as<uint32_t>(total + 1);
        codes = ws.codes.as<uint32_t>(max_entries);
        ranks = ws.ranks.as<uint32_t>(max_entries);
        scan_tmp = ws.scan_tmp.as<uint32_t>(k_scan_tmp_words(total + 1));
        bins = ws.bins.as<uint32_t>(2 * (nbins + 1));
        bin_cursor = bins + nbins + 1;
        const size_t item_min = std::min<size_t>(k_item_max(), dv.item_max_override ? std::min<size_t>(ITEM_FLOOR, dv.item_max_override) : ITEM_FLOOR);
        item_min_ = item_min;
        const size_t max_split_buckets = max_entries / item_min + 1;     // buckets longer than the shortest item limit
        const size_t max_split = 2 * max_split_buckets + 16;             // chunks they are cut into
        items = (WorkItem *)ws.items.need((total + max_split) * k_item_bytes());
        split_list = ws.item_start.as<uint32_t>(1 + 3 * max_split_buckets);
        entries = ws.entries.as<uint32_t>(max_entries);
        if (max_entries * 4 > SCATTER_TWO_PASS_BYTES) {
            staging = ws.scatter_stage.as<uint32_t>(2 * max_entries + 2);
            part_cursor = ws.cursor.as<uint32_t>((max_entries >> k_scatter_log_part(max_entries)) + 2);
        }
        buckets = ws.buckets.as<uint32_t>(total * 4 * FieldWords<F>::N);
        chunk_out = ws.chunk_out.as<uint32_t>(max_split * 4 * FieldWords<F>::N);
    }

    // stages 1-3 for scalars [lo, lo + cnt) of the call; sc points at the first of them (device)
    void front(const uint32_t *sc, size_t lo, size_t cnt, bool mont, bool timed) {
        stream_t s = dv.stream;
        MsmPlan pl = plan;
        pl.offset = (uint32_t)(first + lo);
        n_entries = cnt * plan.nwin;
        if (timed) dv.timer.mark(0, s);
        // 1. canonical scalars -> signed digits: bucket histogram + per-window code array; the histogram atomic also
        //    hands out the digit's rank inside its bucket
        dev_memset(counts, 0, (total + 1) * sizeof(uint32_t), s);
        k_digit_decompose(s, cnt, sc, mont, pl, counts, codes, ranks);
        if (timed) dv.timer.mark(1, s);
        // 2. bucket offsets (exclusive scan; offsets[total] = number of entries) and the work-item list (bucket slices
        //    ordered by length, longest first).  Longest item: 256 additions when the range is large, shorter when it
        //    is small and the longest item would set the kernel's duration
        k_exclusive_scan(s, counts, counts, total + 1, scan_tmp);
        uint32_t *offsets = counts;
        item_max = (uint32_t)std::min<size_t>(k_item_max(), std::max<size_t>(ITEM_FLOOR, n_entries >> ITEM_SHIFT));
        if (dv.item_max_override) item_max = std::min<uint32_t>(k_item_max(), std::max<uint32_t>(dv.item_max_override, (uint32_t)item_min_));
        dev_memset(bins, 0, (nbins + 1) * sizeof(uint32_t), s);
        k_item_count(s, total, offsets, item_max, bins);
        k_exclusive_scan(s, bins, bins, nbins + 1, scan_tmp);
        copy_d2d(bin_cursor, bins, (nbins + 1) * sizeof(uint32_t), s);
        dev_memset(split_list, 0, sizeof(uint32_t), s);
        k_item_scatter(s, total, offsets, item_max, bin_cursor, items, split_list);
        if (timed) dv.timer.mark(2, s);
        // 3. counting-sort scatter of (point index, sign) into bucket order: position = bucket offset + rank.  Entry
        //    arrays beyond L2 go through the two-pass partitioned scatter (msm_kernels.cuh)
        if (n_entries * 4 > SCATTER_TWO_PASS_BYTES) k_scatter_partitioned(s, cnt, codes, ranks, pl, offsets, n_entries, part_cursor, staging, entries);
        else k_scatter_ranked(s, cnt, codes, ranks, pl, offsets, entries);
        if (timed) dv.timer.mark(3, s);
    }

    // stage 4 for the range `front` handled last: bucket accumulation (the hot kernel) + fold of split buckets.
    // add_to: continue the bucket sums earlier ranges of the same call left behind
    void accumulate(bool add_to) {
        stream_t s = dv.stream;
        size_t split_buckets = n_entries / item_max + 1, split_items = 2 * split_buckets + 16;
        k_accumulate<F>(s, total + split_items, pts, entries, items, bins + nbins, buckets, chunk_out, add_to);
        k_chunk_merge<F>(s, split_buckets, split_list, chunk_out, buckets, add_to, dv.sm_count);
    }

    // stages 5-6.  d_out_xyzz: 4 * FieldWords<F>::N words (may be null), d_out_aff: 2 * FieldWords<F>::N + 1 words (may be null)
    void back(uint32_t *d_out_xyzz, uint32_t *d_out_aff) {
        stream_t s = dv.stream;
        Workspace &ws = dv.ws;
        dv.timer.mark(4, s);
        // 5. parallel bucket reduction: thread levels while the level is work bound (every thread walks 2^log_l
        //    consecutive buckets), then block-cooperative levels (quad additions, scan + tree) for the latency
        //    bound top of the tree
        const uint32_t *X = buckets, *Y1 = nullptr, *Y2 = nullptr;
        uint32_t n_in = plan.nb, shift = 0;
        int flip = 0;
        constexpr size_t PWORDS = 4 * FieldWords<F>::N;
        size_t TILE_LEVEL_MAX, THREAD_LEVEL_GROUPS;
        reduce_split((size_t)plan.bwin * plan.nb, TILE_LEVEL_MAX, THREAD_LEVEL_GROUPS);
        while (n_in > 1) {
            uint32_t n_out, log_l;
            if ((size_t)plan.bwin * n_in > TILE_LEVEL_MAX) {
                // aim at ~2^15 groups: fewer and the serial walk (2 x 2^log_l dependent additions) is latency
                // bound, more and the tile levels above get more blocks than one wave
                log_l = 2;
                while (log_l < REDUCE_LOG_L && ((size_t)plan.bwin * n_in >> log_l) > THREAD_LEVEL_GROUPS) ++log_l;
                uint32_t L = 1u << log_l;
                n_out = (n_in + L - 1) / L;
                uint32_t *Xo = ws.red[flip].as<uint32_t>((size_t)plan.bwin * n_out * PWORDS);
                uint32_t *Yo = ws.red[flip + 1].as<uint32_t>((size_t)plan.bwin * n_out * PWORDS);
                k_reduce_level<F>(s, (size_t)plan.bwin * n_out, X, Y1, n_in, n_out, L, shift, Xo, Yo);
                X = Xo; Y1 = Yo;
            } else {
                uint32_t tile_max = k_tile_entries();
                log_l = 1;
                while ((1u << log_l) < n_in && (1u << log_l) < tile_max) ++log_l;
                uint32_t T = 1u << log_l;
                n_out = (n_in + T - 1) / T;
                uint32_t *Xo = ws.red[flip].as<uint32_t>((size_t)plan.bwin * n_out * PWORDS);
                uint32_t *Y1o = ws.red[flip + 1].as<uint32_t>((size_t)plan.bwin * n_out * PWORDS);
                uint32_t *Y2o = (Y1 || Y2) ? ws.red[flip + 2].as<uint32_t>((size_t)plan.bwin * n_out * PWORDS) : nullptr;
                k_tile_reduce<F>(s, plan.bwin, X, Y1, Y2, n_in, n_out, T, shift, Xo, Y1o, Y2o);
                X = Xo; Y1 = Y1o; Y2 = Y2o;
            }
            n_in = n_out; shift += log_l; flip ^= 3;
        }
        dv.timer.mark(5, s);
        // 6. window fold + to affine
        k_window_combine<F>(s, X, Y1, Y2, plan.bwin, plan.c, d_out_xyzz, d_out_aff);
        dv.timer.mark(6, s);
    }
};

// One MSM on one device, asynchronous on dv.stream.
//   d_scalars  : n x 8 u32 on this device
//   d_out_xyzz : 4*FieldWords<F>::N words (may be null), d_out_aff : 2*FieldWords<F>::N + 1 words (may be null)
// h_scalars (optional): the scalars still live on the host and are copied into d_scalars here.  From pipe_min scalars
// on, in H2D_PIPE_PARTS ranges on a second stream, each range going through stages 1-4 into the shared bucket array as
// soon as it has arrived (see H2D_PIPE_PARTS above); below that, one copy in front of the pipeline.
template <class F>
void msm_run(Device &dv, const BasesShard &sh, const uint32_t *d_scalars, size_t n, bool mont, unsigned c_override,
             uint32_t *d_out_xyzz, uint32_t *d_out_aff, size_t first = 0, const uint64_t *h_scalars = nullptr,
             size_t pipe_min = ~(size_t)0) {
    stream_t s = dv.stream;
    if (n == 0) {
        k_partial_combine<F>(s, nullptr, 0u, d_out_xyzz, d_out_aff);
        return;
    }
    // index ranges of the scalars that run stages 1-4 one after the other (one range unless the H2D copy is pipelined)
    size_t part_lo[H2D_PIPE_PARTS + 1] = {0, n, n, n}, parts = 1;
    if (h_scalars && n >= pipe_min && n >= 64) {
        parts = H2D_PIPE_PARTS;
        for (size_t k = 1; k < parts; ++k) part_lo[k] = n * H2D_PIPE_NUM[k] / H2D_PIPE_NUM[parts];
        part_lo[parts] = n;
    }
    size_t n_max = 0;
    for (size_t k = 0; k < parts; ++k) n_max = std::max(n_max, part_lo[k + 1] - part_lo[k]);
    MsmJob<F> job(dv, sh, n, c_override, first, n_max);

    uint32_t *d_stage = const_cast<uint32_t *>(d_scalars);   // written only when the scalars come from the host
    // host scalars: all copies are queued first, back to back on the copy lane; `arrived[k]` fires when range k is in
    struct Arrivals {   // events not yet consumed are released if a launch below throws
        event_t ev[H2D_PIPE_PARTS] = {};
        bool live[H2D_PIPE_PARTS] = {};
        ~Arrivals() {
#ifndef G16_EMU
            for (size_t k = 0; k < H2D_PIPE_PARTS; ++k)
                if (live[k]) cudaEventDestroy(ev[k]);
#endif
        }
    } arrived;
    if (h_scalars && parts > 1) {
        Device &cp = lane_of(dv, H2D_PIPE_LANE);
        stream_wait(cp.stream, s);   // earlier work on s may still read the scalar buffer
        for (size_t k = 0; k < parts; ++k) {
            size_t lo = part_lo[k], cnt = part_lo[k + 1] - lo;
            copy_h2d(d_stage + lo * 8, h_scalars + lo * 4, cnt * 32, cp.stream);
            arrived.ev[k] = event_record(cp.stream);
            arrived.live[k] = true;
        }
    } else if (h_scalars) {
        copy_h2d(d_stage, h_scalars, n * 32, s);
    }
    for (size_t k = 0; k < parts; ++k) {
        const size_t lo = part_lo[k], cnt = part_lo[k + 1] - lo;
        if (cnt == 0) continue;
        if (h_scalars && parts > 1) {
            arrived.live[k] = false;
            event_wait_and_release(s, arrived.ev[k]);
        }
        job.front(d_scalars + lo * 8, lo, cnt, mont, k == 0);
        job.accumulate(k > 0);   // later ranges continue the bucket sums of the earlier ones
    }
    job.back(d_out_xyzz, d_out_aff);
}

// Import host points (ark layout + infinity bytes) into a device shard.
template <class F>
uint32_t *import_points(Device &dv, const uint64_t *xy, const uint8_t *inf, size_t n) {
    set_device(dv.id);
    stream_t s = dv.stream;
    size_t words = n * 2 * FieldWords<F>::N;
    uint32_t *pts = (uint32_t *)dev_alloc(words * 4);
    try {
        if (!inf) {
            copy_h2d(pts, xy, words * 4, s);
        } else {
            uint32_t *stage = dv.ws.staging.as<uint32_t>(words + (n + 3) / 4 + 4);
            uint8_t *d_inf = (uint8_t *)(stage + words);
            copy_h2d(stage, xy, words * 4, s);
            copy_h2d(d_inf, inf, n, s);
            k_import_bases<F>(s, n, stage, d_inf, pts);
        }
        stream_sync(s);
    } catch (...) {
        dev_free(pts);
        throw;
    }
    return pts;
}

template <class F>
std::unique_ptr<Bases> bases_upload(Context *ctx, const uint64_t *xy, const uint8_t *inf, size_t n) {
    std::unique_ptr<Bases> b(new Bases);
    b->ctx = ctx; b->group = GroupOf<F>::id; b->n = n;
    size_t ndev = ctx->devs.size();
    for (size_t d = 0; d < ndev; ++d) {
        size_t begin = n * d / ndev, end = n * (d + 1) / ndev;
        BasesShard sh;
        sh.dev = (int)d; sh.cuda_dev = ctx->devs[d].id; sh.begin = begin; sh.n = end - begin; sh.owned = true;
        sh.pts = import_points<F>(ctx->devs[d], xy + begin * (FieldWords<F>::N), inf ? inf + begin : nullptr, sh.n);
        b->shards.push_back(sh);
    }
    return b;
}

// Build the table of multiples for every shard (one-time, at upload).  c = 0: choose from the shard size
// and `budget_bytes` of device memory per shard.  Returns the window size used (0 = not applicable).
template <class F>
unsigned bases_precompute(Context *ctx, Bases *bases, unsigned c, size_t budget_bytes, unsigned scalar_bits = 255) {
    unsigned used = 0;
    struct Pending { BasesShard *sh; uint32_t *table; unsigned cc; };
    std::vector<Pending> pending;
    auto drop = [&] {
        for (auto &p : pending) { set_device_nothrow(ctx->devs[p.sh->dev].id); dev_free(p.table); }
    };
    try {
        // launch on every device first (the shards are independent), then wait for all of them
        for (auto &sh : bases->shards) {
            if (sh.n == 0) continue;
            Device &dv = ctx->devs[sh.dev];
            set_device(dv.id);
            size_t point_bytes = 2 * FieldWords<F>::N * 4;
            unsigned cc = c ? c : choose_precompute_c(sh.n, point_bytes, budget_bytes, scalar_bits);
            if (!c && cc == 0) continue;   // nothing fits the budget: this shard stays plain
            if (cc < 8 || cc > 24) throw Error{G16_ERR_INVALID, "precompute window bits must be in [8, 24]"};
            uint32_t nwin = (256 + cc - 1) / cc;
            if ((double)nwin * (double)sh.n >= 2147483648.0) throw Error{G16_ERR_INVALID, "precomputed table exceeds 2^31 points"};
            if (sh.table) { dev_free(sh.table); sh.table = nullptr; sh.pre_c = 0; }
            uint32_t *table = (uint32_t *)dev_alloc((size_t)nwin * sh.n * point_bytes);
            pending.push_back(Pending{&sh, table, cc});
            k_precompute_bases<F>(dv.stream, sh.n, sh.pts, cc, nwin, table);
        }
        for (auto &p : pending) {
            Device &dv = ctx->devs[p.sh->dev];
            set_device(dv.id);
            stream_sync(dv.stream);
        }
    } catch (...) { drop(); throw; }
    for (auto &p : pending) { p.sh->table = p.table; p.sh->pre_c = p.cc; used = p.cc; }
    return used;
}

// Host scalars -> host affine result over all shards of `bases`, in two halves so that several MSMs
// can be in flight on different lanes: msm_launch issues the H2D copies and the whole pipeline of every
// shard asynchronously -- no call in it waits for a device, so the shards of a multi-device context really run
// side by side -- and every shard sends its 192 / 384-byte partial sum straight into device 0's `partials` buffer
// (peer copy on the shard's stream).  msm_finish makes device 0 wait for those copies, folds them and returns the point.
template <class F>
void msm_launch(Context *ctx, const Bases *bases, const uint64_t *scalars, size_t n, int lane,
                uint32_t *user_xyzz = nullptr, uint32_t *user_aff = nullptr) {
    if (bases->group != GroupOf<F>::id) throw Error{G16_ERR_INVALID, "bases belong to the other group"};
    if (n > bases->n) throw Error{G16_ERR_LENGTH, "more scalars than bases (ark: Err(min_len))"};
    constexpr size_t PW = 4 * FieldWords<F>::N, AW = 2 * FieldWords<F>::N + 1;
    size_t nsh = bases->shards.size();
    bool user_out = user_xyzz || user_aff;   // results stay on the device (single shard only)
    if (user_out && nsh != 1) throw Error{G16_ERR_INVALID, "device outputs need unsharded bases"};
    uint32_t *d0_parts = nullptr;
    if (nsh > 1) {
        Device &l0 = lane_of(ctx->devs[0], lane);
        set_device(l0.id);
        d0_parts = l0.ws.partials.as<uint32_t>(nsh * PW + AW);
    }
    for (size_t k = 0; k < nsh; ++k) {
        const BasesShard &sh = bases->shards[k];
        Device &dv = lane_of(ctx->devs[sh.dev], lane);
        set_device(dv.id);
        size_t lo = std::min(sh.begin, n), hi = std::min(sh.begin + sh.n, n);
        size_t cnt = hi - lo;
        uint32_t *d_out = dv.ws.out.as<uint32_t>(PW + AW);
        uint32_t *d_sc = dv.ws.scalars.as<uint32_t>(cnt * 8 + 8);
        const uint64_t *h_sc = scalars + lo * 4;   // copied inside msm_run, overlapped with the pipeline
        if (nsh == 1) {
            msm_run<F>(dv, sh, d_sc, cnt, true, ctx->c_override, user_xyzz, user_out ? user_aff : d_out + PW, lo - sh.begin, h_sc,
                       ctx->h2d_pipe_min);
        } else {
            msm_run<F>(dv, sh, d_sc, cnt, true, ctx->c_override, d_out, nullptr, lo - sh.begin, h_sc, ctx->h2d_pipe_min);
            copy_peer(d0_parts + k * PW, d_out, PW * 4, dv.stream);
        }
    }
}

template <class F>
void msm_finish(Context *ctx, const Bases *bases, int lane, uint64_t *out_xy, uint8_t *out_inf) {
    constexpr size_t PW = 4 * FieldWords<F>::N, AW = 2 * FieldWords<F>::N + 1;
    size_t nsh = bases->shards.size();
    Device &d0 = lane_of(ctx->devs[0], lane);
    uint32_t aff[AW];
    set_device(d0.id);
    if (nsh == 1) {
        copy_d2h(aff, (uint32_t *)d0.ws.out.p + PW, AW * 4, d0.stream);
        stream_sync(d0.stream);
    } else {
        for (size_t k = 0; k < nsh; ++k) {
            Device &dv = lane_of(ctx->devs[bases->shards[k].dev], lane);
            if (&dv != &d0) stream_wait_xdev(d0.stream, d0.id, dv.stream, dv.id);
        }
        set_device(d0.id);
        uint32_t *d_part = (uint32_t *)d0.ws.partials.p;
        k_partial_combine<F>(d0.stream, d_part, (uint32_t)nsh, nullptr, d_part + nsh * PW);
        copy_d2h(aff, d_part + nsh * PW, AW * 4, d0.stream);
        stream_sync(d0.stream);
    }
    memcpy(out_xy, aff, (AW - 1) * 4);
    if (out_inf) *out_inf = (uint8_t)aff[AW - 1];
}

template <class F>
void msm_host(Context *ctx, const Bases *bases, const uint64_t *scalars, size_t n, uint64_t *out_xy, uint8_t *out_inf) {
    msm_launch<F>(ctx, bases, scalars, n, 0);
    msm_finish<F>(ctx, bases, 0, out_xy, out_inf);
}

// ---------------------------------------------------------------------------------------
// quotient polynomial H = (A B - C) / Z from domain evaluations (see ntt_kernels.cuh)
//   evals: host, 3 arrays (a, b, c) of n Fr each (Montgomery), n = 2^log_n;  h: host, n Fr
// Returns false when A*B - C does not vanish on the domain (the reference's PolynomialDivisionFailed).
// ---------------------------------------------------------------------------------------
// constants, twiddles and scale tables of the size-2^log_n domain, cached per workspace.
// ntt_tw layout (Fr elements): tw[n/2] | twi[n/2] | scale[n] | fscale[n]
inline const uint32_t *ntt_prepare(Device &dv, uint32_t log_n) {
    Workspace &ws = dv.ws;
    uint32_t n = 1u << log_n;
    uint32_t *tw = ws.ntt_tw.as<uint32_t>((size_t)n * 24 + 8);
    uint32_t *consts = ws.ntt_consts.as<uint32_t>(k_ntt_const_words());
    if (ws.ntt_log_n != log_n) {
        k_ntt_setup(dv.stream, log_n, consts);
        if (n >= 2) k_ntt_twiddles(dv.stream, n, consts, tw, tw + (size_t)(n / 2) * 8);
        k_ntt_scale_tables(dv.stream, n, log_n, consts, tw + (size_t)n * 8, tw + (size_t)n * 16);
        ws.ntt_log_n = log_n;
    }
    return consts;
}
// abc: device, the evaluations of A, B, C back to back (3 n Fr, overwritten); flag: device word that counts the
// rows where A*B != C; out: device, n coefficients of H.  Asynchronous on dv.stream.
inline void quotient_device(Device &dv, uint32_t log_n, uint32_t *abc, uint32_t *flag, uint32_t *out) {
    stream_t s = dv.stream;
    uint32_t n = 1u << log_n;
    const uint32_t *consts = ntt_prepare(dv, log_n);
    const uint32_t *tw = (const uint32_t *)dv.ws.ntt_tw.p, *twi = tw + (size_t)(n / 2) * 8;
    const uint32_t *scale = tw + (size_t)n * 8, *fscale = tw + (size_t)n * 16;
    k_ntt_check_vanish(s, abc, n, flag);
    // coefficients (bit-reversed) with the coset shift folded into the last pass, then values on the coset
    k_ntt_transform(s, false, 3, abc, twi, log_n, scale);
    k_ntt_transform(s, true, 3, abc, tw, log_n, nullptr);
    k_ntt_quotient_pointwise(s, abc, consts, n);
    k_ntt_transform(s, false, 1, abc, twi, log_n, nullptr);
    k_ntt_final_permute(s, abc, fscale, n, log_n, out);
}
inline bool quotient_host(Device &dv, const uint64_t *a, const uint64_t *b, const uint64_t *c, uint32_t log_n, uint64_t *h) {
    Workspace &ws = dv.ws;
    stream_t s = dv.stream;
    uint32_t n = 1u << log_n;
    uint32_t *abc = ws.ntt_abc.as<uint32_t>((size_t)3 * n * 8 + 8);
    uint32_t *flag = abc + (size_t)3 * n * 8;
    uint32_t *out = ws.ntt_out.as<uint32_t>((size_t)n * 8);
    copy_h2d(abc, a, (size_t)n * 32, s);
    copy_h2d(abc + (size_t)n * 8, b, (size_t)n * 32, s);
    copy_h2d(abc + (size_t)2 * n * 8, c, (size_t)n * 32, s);
    dev_memset(flag, 0, 4, s);
    quotient_device(dv, log_n, abc, flag, out);
    uint32_t bad = 0;
    copy_d2h(h, out, (size_t)n * 32, s);
    copy_d2h(&bad, flag, 4, s);
    stream_sync(s);
    return bad == 0;
}

// ---------------------------------------------------------------------------------------
// fixed base
// ---------------------------------------------------------------------------------------
template <class F>
const uint32_t *fixed_base_table(Device &dv, const uint64_t *base_xy) {
    constexpr int slot = GroupOf<F>::id;
    Workspace &ws = dv.ws;
    stream_t s = dv.stream;
    const uint32_t *key = (const uint32_t *)base_xy;
    std::vector<uint32_t> &cached = ws.fb_table_key[slot];
    if (cached.size() == 2 * FieldWords<F>::N && memcmp(cached.data(), key, 2 * FieldWords<F>::N * 4) == 0 && ws.fb_table[slot].p)
        return (const uint32_t *)ws.fb_table[slot].p;
    uint32_t *d_base = ws.fb_base.as<uint32_t>(2 * FieldWords<F>::N);
    copy_h2d(d_base, base_xy, 2 * FieldWords<F>::N * 4, s);
    uint32_t *powers = ws.fb_powers.as<uint32_t>(FB_WINDOWS * 4 * FieldWords<F>::N);
    k_fb_powers<F>(s, d_base, powers);
    uint32_t *table = ws.fb_table[slot].as<uint32_t>((size_t)FB_WINDOWS * FB_ENTRIES * 2 * FieldWords<F>::N);
    k_fb_table<F>(s, powers, table);
    cached.assign(key, key + 2 * FieldWords<F>::N);
    return table;
}

template <class F>
void fixed_base_device(Device &dv, const uint64_t *base_xy, const uint32_t *d_scalars, size_t n, uint32_t *d_out) {
    const uint32_t *table = fixed_base_table<F>(dv, base_xy);
    k_fb_mul<F>(dv.stream, n, d_scalars, true, table, d_out);
}

template <class F>
void fixed_base_host(Context *ctx, const uint64_t *base_xy, const uint64_t *scalars, size_t n, uint64_t *out_xy,
                     uint8_t *out_inf) {
    size_t ndev = ctx->devs.size();
    constexpr size_t W = 2 * FieldWords<F>::N;
    for (size_t d = 0; d < ndev; ++d) {
        Device &dv = ctx->devs[d];
        set_device(dv.id);
        size_t lo = n * d / ndev, hi = n * (d + 1) / ndev, cnt = hi - lo;
        if (!cnt) continue;
        uint32_t *d_sc = dv.ws.scalars.as<uint32_t>(cnt * 8);
        copy_h2d(d_sc, scalars + lo * 4, cnt * 32, dv.stream);
        uint32_t *d_out = dv.ws.fb_out.as<uint32_t>(cnt * W);
        fixed_base_device<F>(dv, base_xy, d_sc, cnt, d_out);
        copy_d2h(out_xy + lo * (W / 2), d_out, cnt * W * 4, dv.stream);
        if (out_inf) {
            uint8_t *d_fl = dv.ws.fb_flags.as<uint8_t>(cnt);
            k_export_flags<F>(dv.stream, cnt, d_out, d_fl);
            copy_d2h(out_inf + lo, d_fl, cnt, dv.stream);
        }
    }
    for (size_t d = 0; d < ndev; ++d) { set_device(ctx->devs[d].id); stream_sync(ctx->devs[d].stream); }
}

}  // namespace g16
