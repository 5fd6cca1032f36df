// Host-callable launchers of the kernels, one declaration per kernel.  The definitions live
// in kernel_impl.cuh and are explicitly instantiated in the k_*.cu translation units so that
// the big kernels compile in parallel (and the cold ones with out-of-line field multiplies).
#pragma once
#include "rt.cuh"

namespace g16 {

struct MsmPlan {
    uint32_t c;        // window bits
    uint32_t nwin;     // number of windows = ceil(256 / c)
    uint32_t nb;       // buckets per window = 2^(c-1)
    uint32_t total;    // bwin * nb
    uint32_t bwin;     // windows that own a bucket set: nwin, or 1 when the bases carry precomputed
                       // multiples 2^(c w) P (then every window feeds the same buckets)
    uint32_t stride;   // precomputed tables: index of (w, i) is w * stride + i; 0 otherwise
    uint32_t offset;   // index of this call's first base inside the resident array (chunked launches)
};

template <class P> struct Fp;
struct FqParams;
struct Fq2;
using Fq = Fp<FqParams>;

// u32 words of one coordinate
template <class F> struct FieldWords;
template <> struct FieldWords<Fq> { static constexpr size_t N = 12; static constexpr int group = 1; };
template <> struct FieldWords<Fq2> { static constexpr size_t N = 24; static constexpr int group = 2; };

// scalars -> bucket histogram / bucket-ordered entries
// digits of scalars [i0, i0 + cnt) of an n-scalar call
void k_digit_decompose(stream_t s, size_t n, const uint32_t *scalars, bool mont, MsmPlan plan, uint32_t *counts,
                        uint32_t *codes, uint32_t *ranks, size_t i0 = 0, size_t cnt = ~(size_t)0);
void k_scatter_ranked(stream_t s, size_t n, const uint32_t *codes, const uint32_t *ranks, MsmPlan plan,
                      const uint32_t *offsets, uint32_t *entries);
// two-pass scatter through a staging area of (position, entry) pairs (device build only; see msm_kernels.cuh):
// staging holds 2 words per entry, part_cursor k_scatter_parts(..) zeroed words
uint32_t k_scatter_log_part(size_t max_entries);
void k_scatter_partitioned(stream_t s, size_t n, const uint32_t *codes, const uint32_t *ranks, MsmPlan plan,
                           const uint32_t *offsets, size_t max_entries, uint32_t *part_cursor, uint32_t *staging,
                           uint32_t *entries);
// work items (bucket slices ordered by length); see msm_kernels.cuh
struct WorkItem;
size_t k_item_bins();
size_t k_item_bytes();
uint32_t k_item_max();
void k_item_count(stream_t s, size_t buckets, const uint32_t *offsets, uint32_t item_max, uint32_t *bin_counts);
void k_item_scatter(stream_t s, size_t buckets, const uint32_t *offsets, uint32_t item_max, uint32_t *bin_cursor,
                    WorkItem *items, uint32_t *split_list);
size_t k_scan_tmp_words(size_t n);
void k_exclusive_scan(stream_t s, const uint32_t *in, uint32_t *out, size_t n, uint32_t *tmp);

template <class F>
// add_to: continue the sums already stored in `buckets` (an earlier chunk of the same MSM) instead of starting at infinity
void k_accumulate(stream_t s, size_t max_items, const uint32_t *pts, const uint32_t *entries, const WorkItem *work,
                  const uint32_t *n_items, uint32_t *buckets, uint32_t *chunk_out, bool add_to);
template <class F>
void k_chunk_merge(stream_t s, size_t max_split, const uint32_t *split_list, const uint32_t *chunk_out, uint32_t *buckets,
                   bool add_to, uint32_t sm_count);
template <class F>
void k_reduce_level(stream_t s, size_t threads, const uint32_t *X, const uint32_t *Y, uint32_t n_in, uint32_t n_out,
                    uint32_t L, uint32_t shift, uint32_t *Xo, uint32_t *Yo);
// block-cooperative level: tile = T entries (power of two <= k_tile_entries()), grid (n_out, windows, X | Y blocks)
uint32_t k_tile_entries();
template <class F>
void k_tile_reduce(stream_t s, uint32_t windows, const uint32_t *X, const uint32_t *Y1, const uint32_t *Y2, uint32_t n_in,
                   uint32_t n_out, uint32_t T, uint32_t shift, uint32_t *Xo, uint32_t *Y1o, uint32_t *Y2o);
template <class F>
void k_window_combine(stream_t s, const uint32_t *X, const uint32_t *Y, const uint32_t *Y2, uint32_t nwin, uint32_t c,
                      uint32_t *out_xyzz, uint32_t *out_aff);
template <class F>
void k_partial_combine(stream_t s, const uint32_t *partials, uint32_t k, uint32_t *out_xyzz, uint32_t *out_aff);
// out_xyzz[j] = scalars[j] * aff[j] for n points (affine records `stride` words apart: x, y, infinity word)
template <class F>
void k_scalar_mul_affine(stream_t s, size_t n, const uint32_t *scalars, const uint32_t *aff, uint32_t stride, uint32_t *out_xyzz);
// table[w * n + i] = affine(2^(c w) * pts[i]) for w < nwin (w = 0 is a copy)
template <class F>
void k_precompute_bases(stream_t s, size_t n, const uint32_t *pts, uint32_t c, uint32_t nwin, uint32_t *table);
template <class F>
void k_import_bases(stream_t s, size_t n, const uint32_t *xy, const uint8_t *inf, uint32_t *pts);
template <class F>
void k_export_flags(stream_t s, size_t n, const uint32_t *pts, uint8_t *inf);

// Fixed-base windows: signed digits of FB_BITS bits, d in [-(2^(FB_BITS-1) - 1), 2^(FB_BITS-1)], so a window's table holds
// the 2^(FB_BITS-1) positive multiples only (the sign flips y).  16 bits: 16 mixed additions per scalar, tables of
// 16 x 32768 affine points = 50 MB (G1) / 100 MB (G2) per base, L2 resident on B200 (126 MB).  The host-emulation build
// (tests only) walks the same code with 8-bit windows: its one-thread-at-a-time loop cannot build 2^19 table entries.
#ifndef G16_EMU
constexpr uint32_t FB_BITS = 16;
#else
constexpr uint32_t FB_BITS = 8;
#endif
constexpr uint32_t FB_WINDOWS = 256 / FB_BITS;         // 16
constexpr uint32_t FB_ENTRIES = 1u << (FB_BITS - 1);   // multiples 1 .. 2^(FB_BITS-1) of 2^(FB_BITS j) * base
constexpr uint32_t FB_TABLE_GROUP = 8;                 // consecutive table entries one thread builds (shared inversion)
template <class F>
void k_fb_powers(stream_t s, const uint32_t *base_xy, uint32_t *powers);
template <class F>
void k_fb_table(stream_t s, const uint32_t *powers, uint32_t *table);
template <class F>
void k_fb_mul(stream_t s, size_t n, const uint32_t *scalars, bool mont, const uint32_t *table, uint32_t *out);

// quotient polynomial (Fr NTT), see ntt_kernels.cuh
size_t k_ntt_const_words();
void k_ntt_setup(stream_t s, uint32_t log_n, uint32_t *consts);
void k_ntt_twiddles(stream_t s, uint32_t n, const uint32_t *consts, uint32_t *tw, uint32_t *twi);
void k_ntt_stage(stream_t s, bool dit, size_t batch, uint32_t *x, const uint32_t *tw, uint32_t n, uint32_t half);
// whole transform (fused stages on the device); scale: optional per-position factors applied with a DIF transform
void k_ntt_transform(stream_t s, bool dit, size_t batch, uint32_t *x, const uint32_t *tw, uint32_t log_n, const uint32_t *scale);
// scale[p] = g^br(p) / n, fscale[j] = g^-j / n (cached per domain size next to the twiddles)
void k_ntt_scale_tables(stream_t s, uint32_t n, uint32_t log_n, const uint32_t *consts, uint32_t *scale, uint32_t *fscale);
void k_ntt_final_permute(stream_t s, const uint32_t *x, const uint32_t *fscale, uint32_t n, uint32_t log_n, uint32_t *out);
void k_ntt_quotient_pointwise(stream_t s, uint32_t *abc, const uint32_t *consts, uint32_t n);
void k_ntt_check_vanish(stream_t s, const uint32_t *abc, uint32_t n, uint32_t *flag);

// sparse R1CS (Fr), see r1cs_kernels.cuh.  Stacked CSR: `lines` lines, out[(t / seg) * seg_out + t % seg]
void k_spmv(stream_t s, size_t lines, const uint32_t *line_ptr, const uint32_t *idx, const uint32_t *val, const uint32_t *vec,
            uint32_t seg, uint32_t seg_out, const uint32_t *long_lines, size_t n_long, uint32_t *out);
uint32_t k_spmv_long_threshold();
void k_truncate64(stream_t s, size_t n, const uint32_t *in, uint32_t *out);
size_t k_setup_scalar_words();
void k_setup_scalars(stream_t s, const uint32_t *params, const uint32_t *consts, uint32_t log_n, uint32_t truncate, uint32_t *blk);
void k_lagrange_at(stream_t s, size_t n, const uint32_t *consts, const uint32_t *blk, uint32_t *out);
void k_crs_exponents(stream_t s, size_t num_vars, const uint32_t *vals, const uint32_t *blk, uint32_t num_public, uint32_t *ab,
                     uint32_t *ic);
void k_crs_h_exponents(stream_t s, size_t n, const uint32_t *blk, uint32_t *h);
void k_validate_row(stream_t s, const uint32_t *abc, uint32_t n, uint32_t *flag);

// wire format (ark CanonicalSerialize of G1Affine / G2Affine), see wire_kernels.cuh
constexpr uint8_t WIRE_STATUS_OK = 0, WIRE_STATUS_INVALID_DATA = 1, WIRE_STATUS_UNEXPECTED_FLAGS = 2;
template <class F>
void k_point_encode(stream_t s, size_t n, const uint32_t *pts, bool compressed, uint32_t *out_bytes);
template <class F>
void k_point_decode(stream_t s, size_t n, const uint32_t *in_bytes, bool compressed, bool validate, uint32_t *pts, uint8_t *status);

// test hooks
void k_debug_fq_op(stream_t s, size_t n, int op, const uint32_t *a, const uint32_t *b, uint32_t *out);
void k_debug_fr_from_mont(stream_t s, size_t n, const uint32_t *a, uint32_t *out);
template <class F>
void k_debug_add(stream_t s, size_t n, const uint32_t *p, const uint32_t *q, uint32_t *out);
unsigned long long launch_count();

}  // namespace g16
