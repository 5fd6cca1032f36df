// Exclusive prefix sum over u32 counters (bucket sizes -> bucket offsets).  Hand-written
// block scan (warp shuffles + one shared-memory hop) applied recursively; the arrays are at
// most a few million counters, i.e. a few percent of one HBM pass over the point stream.
#pragma once
#include "rt.cuh"

namespace g16 {

#ifndef G16_EMU
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const uint32_t *in, uint32_t *out, uint32_t *tile_sums,
                                                                  size_t n) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        sum += v[k];
    }
    uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t wi = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += o;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;  // exclusive
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + incl - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *out, const uint32_t *tile_prefix, size_t n) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_prefix[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// in-place capable (in == out).  `tmp` must hold scan_tmp_words(n) u32.
inline size_t scan_tmp_words(size_t n) {
    size_t total = 0;
    while (n > 1) { n = (n + SCAN_TILE - 1) / SCAN_TILE; total += n; if (n == 1) break; }
    return total + 1;
}
inline void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *tmp, stream_t s) {
    if (n == 0) return;
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, out, tiles > 1 ? tmp : nullptr, n);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
    if (tiles > 1) {
        exclusive_scan_u32(tmp, tmp, tiles, tmp + tiles, s);
        scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(out, tmp, n);
        G16_CUDA_CHECK(cudaGetLastError());
        note_launch();
    }
}
#else
inline size_t scan_tmp_words(size_t) { return 1; }
inline void exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, uint32_t *, stream_t) {
    uint32_t run = 0;
    for (size_t i = 0; i < n; ++i) { uint32_t v = in[i]; out[i] = run; run += v; }
}
#endif

}  // namespace g16
