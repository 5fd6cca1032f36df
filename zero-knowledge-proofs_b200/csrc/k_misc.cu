// digit decomposition, counting sort helpers, prefix scan
#include "kernel_impl.cuh"
#include "scan.cuh"
#include <atomic>
namespace g16 {
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void k_digit_count(stream_t s, size_t n, const uint32_t *scalars, bool mont, MsmPlan plan, uint32_t *counts) {
    launch<DigitCount>(n, s, scalars, mont, plan, counts);
}
void k_digit_scatter(stream_t s, size_t n, const uint32_t *scalars, bool mont, MsmPlan plan, uint32_t *cursor,
                     uint32_t *entries) {
    launch<DigitScatter>(n, s, scalars, mont, plan, cursor, entries);
}
size_t k_scan_tmp_words(size_t n) { return scan_tmp_words(n); }
void k_exclusive_scan(stream_t s, const uint32_t *in, uint32_t *out, size_t n, uint32_t *tmp) {
    exclusive_scan_u32(in, out, n, tmp, s);
}
}  // namespace g16
