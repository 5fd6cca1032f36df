// Definitions of the launchers declared in kernel_api.cuh (see there).
#pragma once
#include <algorithm>
#include "kernel_api.cuh"
#include "fixed_base_kernels.cuh"
#include "msm_kernels.cuh"
#include "pair_g2.cuh"
#include <type_traits>

namespace g16 {

template <class F>
void k_accumulate(stream_t s, size_t max_items, const uint32_t *pts, const uint32_t *entries, const WorkItem *work,
                  const uint32_t *n_items, uint32_t *buckets, uint32_t *chunk_out, bool add_to) {
#if !defined(G16_EMU) && !G16_G2_ACC_THREAD
    if constexpr (std::is_same<F, Fq2>::value) {   // G2 on the device: one lane pair per work item (pair_g2.cuh)
        if (max_items == 0) return;
        unsigned blocks = (unsigned)((2 * max_items + G16_PAIR_BLOCK - 1) / G16_PAIR_BLOCK);
        if (add_to) accumulate_pair_g2_kernel<true><<<blocks, G16_PAIR_BLOCK, 0, s>>>(pts, entries, work, n_items, buckets, chunk_out);
        else accumulate_pair_g2_kernel<false><<<blocks, G16_PAIR_BLOCK, 0, s>>>(pts, entries, work, n_items, buckets, chunk_out);
        G16_CUDA_CHECK(cudaGetLastError());
        note_launch();
        return;
    }
#endif
    if (add_to) launch<BucketAccumulate<F, true>>(max_items, s, pts, entries, work, n_items, buckets, chunk_out);
    else launch<BucketAccumulate<F, false>>(max_items, s, pts, entries, work, n_items, buckets, chunk_out);
}
template <class F>
void k_chunk_merge(stream_t s, size_t max_split, const uint32_t *split_list, const uint32_t *chunk_out, uint32_t *buckets,
                   bool add_to, uint32_t sm_count) {
#ifndef G16_EMU
    launch<ChunkMergeSerial<F>>(max_split, s, split_list, chunk_out, MERGE_SERIAL_MAX, buckets, add_to ? 1u : 0u);
    size_t smem = (size_t)MERGE_THREADS * 4 * F::N * sizeof(uint32_t);
    // grid-stride over the split list: four blocks per SM of this device
    chunk_merge_kernel<F><<<std::max(1u, sm_count) * 4, MERGE_THREADS, smem, s>>>(split_list, chunk_out, buckets, add_to ? 1u : 0u);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
#else
    (void)sm_count;
    launch<ChunkMergeSerial<F>>(max_split, s, split_list, chunk_out, 0xffffffffu, buckets, add_to ? 1u : 0u);
#endif
}
template <class F>
void k_reduce_level(stream_t s, size_t threads, const uint32_t *X, const uint32_t *Y, uint32_t n_in, uint32_t n_out,
                    uint32_t L, uint32_t shift, uint32_t *Xo, uint32_t *Yo) {
    launch<ReduceLevel<F>>(threads, s, X, Y, n_in, n_out, L, shift, Xo, Yo);
}
template <class F>
void k_tile_reduce(stream_t s, uint32_t windows, const uint32_t *X, const uint32_t *Y1, const uint32_t *Y2, uint32_t n_in,
                   uint32_t n_out, uint32_t T, uint32_t shift, uint32_t *Xo, uint32_t *Y1o, uint32_t *Y2o) {
#ifndef G16_EMU
    // T = entries per tile; elements (quads of lanes) = T / TILE_K, at least one warp of 8
    uint32_t elems = T / TileShape<F>::K < 8 ? 8 : T / TileShape<F>::K;
    size_t smem = (size_t)elems * (4 * F::N + 1) * sizeof(uint32_t);
    tile_reduce_kernel<F><<<dim3(n_out, windows, (Y1 || Y2) ? 2 : 1), 4 * elems, smem, s>>>(X, Y1, Y2, n_in, n_out, T, shift, Xo, Y1o, Y2o);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
#else
    launch<TileReduceSerial<F>>((size_t)windows * n_out, s, X, Y1, Y2, n_in, n_out, T, shift, Xo, Y1o, Y2o);
#endif
}
template <class F>
void k_window_combine(stream_t s, const uint32_t *X, const uint32_t *Y, const uint32_t *Y2, uint32_t nwin, uint32_t c,
                      uint32_t *out_xyzz, uint32_t *out_aff) {
    launch<WindowCombine<F>>(1, s, X, Y, Y2, nwin, c, out_xyzz, out_aff);
}
template <class F>
void k_partial_combine(stream_t s, const uint32_t *partials, uint32_t k, uint32_t *out_xyzz, uint32_t *out_aff) {
#ifndef G16_EMU
    if (k >= 4) {   // a handful of partials (one per GPU): warp-cooperative fold
        size_t smem = (size_t)8 * (4 * F::N + 1) * sizeof(uint32_t);
        partial_combine_warp_kernel<F><<<1, 32, smem, s>>>(partials, k, out_xyzz, out_aff);
        G16_CUDA_CHECK(cudaGetLastError());
        note_launch();
        return;
    }
#endif
    launch<PartialCombine<F>>(1, s, partials, k, out_xyzz, out_aff);
}
template <class F>
void k_scalar_mul_affine(stream_t s, size_t n, const uint32_t *scalars, const uint32_t *aff, uint32_t stride, uint32_t *out_xyzz) {
#ifndef G16_EMU
    for (size_t j0 = 0; j0 < n; j0 += 8) {   // eight chains per warp
        uint32_t cnt = (uint32_t)std::min<size_t>(8, n - j0);
        scalar_mul_quad_kernel<F><<<1, 32, 0, s>>>(cnt, scalars + j0 * 8, aff + j0 * stride, stride, out_xyzz + j0 * 4 * F::N);
        G16_CUDA_CHECK(cudaGetLastError());
        note_launch();
    }
#else
    launch<ScalarMulAffine<F>>(n, s, scalars, aff, stride, out_xyzz);
#endif
}
template <class F>
void k_precompute_bases(stream_t s, size_t n, const uint32_t *pts, uint32_t c, uint32_t nwin, uint32_t *table) {
    launch<PrecomputeBases<F>>(n, s, pts, n, c, nwin, table);
}
template <class F>
void k_import_bases(stream_t s, size_t n, const uint32_t *xy, const uint8_t *inf, uint32_t *pts) {
    launch<ImportBases<F>>(n, s, xy, inf, pts);
}
template <class F>
void k_export_flags(stream_t s, size_t n, const uint32_t *pts, uint8_t *inf) {
    launch<ExportFlags<F>>(n, s, pts, inf);
}
template <class F>
void k_fb_powers(stream_t s, const uint32_t *base_xy, uint32_t *powers) {
    launch<FbPowers<F>>(1, s, base_xy, powers);
}
template <class F>
void k_fb_table(stream_t s, const uint32_t *powers, uint32_t *table) {
    launch<FbTable<F>>((size_t)FB_WINDOWS * (FB_ENTRIES / FB_TABLE_GROUP), s, powers, table);
}
template <class F>
void k_fb_mul(stream_t s, size_t n, const uint32_t *scalars, bool mont, const uint32_t *table, uint32_t *out) {
    launch<FbMul<F>>((n + FB_GROUP - 1) / FB_GROUP, s, scalars, mont, table, n, out);
}

}  // namespace g16
