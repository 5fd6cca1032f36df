// digit decomposition, counting sort helpers, prefix scan
#include "kernel_impl.cuh"
#include "scan.cuh"
#include <atomic>
namespace g16 {
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void k_digit_decompose(stream_t s, size_t n, const uint32_t *scalars, bool mont, MsmPlan plan, uint32_t *counts,
                        uint32_t *codes, uint32_t *ranks, size_t i0, size_t cnt) {
    if (cnt == ~(size_t)0) cnt = n - i0;
    launch<DigitDecompose>(cnt, s, scalars, mont, plan, n, counts, codes, ranks, i0);
}
void k_scatter_ranked(stream_t s, size_t n, const uint32_t *codes, const uint32_t *ranks, MsmPlan plan,
                      const uint32_t *offsets, uint32_t *entries) {
    launch<ScatterRanked>(n * plan.nwin, s, codes, ranks, plan, n, offsets, entries);
}
// about 256 partitions of at least 2^20 entries (4 MB of `entries`): a partition stays L2 resident while pass B fills
// it, and a block's tile leaves runs of ~16 pairs per partition, so its staging writes coalesce
uint32_t k_scatter_log_part(size_t max_entries) {
    uint32_t lp = 20;
    while ((max_entries >> lp) + 1 > 256) ++lp;
    return lp;
}
void k_scatter_partitioned(stream_t s, size_t n, const uint32_t *codes, const uint32_t *ranks, MsmPlan plan,
                           const uint32_t *offsets, size_t max_entries, uint32_t *part_cursor, uint32_t *staging,
                           uint32_t *entries) {
#ifndef G16_EMU
    uint32_t log_part = k_scatter_log_part(max_entries);
    uint32_t n_parts = (uint32_t)(max_entries >> log_part) + 1;
    size_t total = n * plan.nwin;
    size_t blocks = (total + SCATTER_TILE - 1) / SCATTER_TILE;
    G16_CUDA_CHECK(cudaMemsetAsync(part_cursor, 0, n_parts * sizeof(uint32_t), s));
    size_t smem = (size_t)SCATTER_TILE * sizeof(uint2) + 3 * (size_t)n_parts * sizeof(uint32_t);
    scatter_partition_kernel<<<(unsigned)blocks, SCATTER_THREADS, smem, s>>>(
        codes, ranks, plan, n, offsets, log_part, n_parts, part_cursor, reinterpret_cast<uint2 *>(staging));
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
    // the entry count sits behind the last bucket offset
    launch<ScatterFinal>(max_entries, s, (const uint32_t *)staging, offsets + plan.total, entries);
#else
    (void)max_entries; (void)part_cursor; (void)staging;
    launch<ScatterRanked>(n * plan.nwin, s, codes, ranks, plan, n, offsets, entries);
#endif
}
size_t k_item_bins() { return ITEM_BINS; }
size_t k_item_bytes() { return sizeof(WorkItem); }
uint32_t k_item_max() { return ITEM_MAX; }
uint32_t k_tile_entries() { return (uint32_t)(TILE_K * TILE_ELEMS); }
void k_item_count(stream_t s, size_t buckets, const uint32_t *offsets, uint32_t item_max, uint32_t *bin_counts) {
#ifndef G16_EMU
    if (!buckets) return;
    item_count_kernel<<<(unsigned)((buckets + 255) / 256), 256, 0, s>>>(buckets, offsets, item_max, bin_counts);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
#else
    launch<ItemCount>(buckets, s, offsets, item_max, bin_counts);
#endif
}
void k_item_scatter(stream_t s, size_t buckets, const uint32_t *offsets, uint32_t item_max, uint32_t *bin_cursor,
                    WorkItem *items, uint32_t *split_list) {
#ifndef G16_EMU
    if (!buckets) return;
    item_scatter_kernel<<<(unsigned)((buckets + 255) / 256), 256, 0, s>>>(buckets, offsets, item_max, bin_cursor, items, split_list);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
#else
    launch<ItemScatter>(buckets, s, offsets, item_max, bin_cursor, items, split_list);
#endif
}
size_t k_scan_tmp_words(size_t n) { return scan_tmp_words(n); }
void k_exclusive_scan(stream_t s, const uint32_t *in, uint32_t *out, size_t n, uint32_t *tmp) {
    exclusive_scan_u32(in, out, n, tmp, s);
}
}  // namespace g16
