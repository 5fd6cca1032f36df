// Thin runtime layer: device memory, copies and per-thread kernel launches.
//
// Product build (nvcc, sm_100a): real CUDA allocations, stream-ordered copies, and one
// `__global__` wrapper per kernel body.
// G16_EMU build (g++, tests/emu only): the same kernel *bodies* are executed by a host loop so
// the pipeline logic can be single-stepped in the GPU-less build container.  G16_EMU is never
// defined for libg16cuda.so; there is no runtime switch between the two.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "g16_defs.cuh"
#include "../../include/g16_cuda.h"  // error codes

#ifndef G16_EMU
#include <cuda_runtime.h>
#endif

namespace g16 {

struct Error {
    int code;
    std::string msg;
};

#ifndef G16_EMU
#define G16_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            throw ::g16::Error{G16_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)}; \
    } while (0)
typedef cudaStream_t stream_t;
#else
typedef void *stream_t;
#endif


inline void *dev_alloc(size_t bytes) {
    if (bytes == 0) bytes = 16;
#ifndef G16_EMU
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) throw Error{G16_ERR_OOM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e)};
    return p;
#else
    void *p = malloc(bytes);
    if (!p) throw Error{G16_ERR_OOM, "malloc"};
    return p;
#endif
}
inline void dev_free(void *p) {
    if (!p) return;
#ifndef G16_EMU
    cudaFree(p);
#else
    free(p);
#endif
}
inline void copy_h2d(void *d, const void *h, size_t bytes, stream_t s) {
    if (!bytes) return;
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
#else
    (void)s; memcpy(d, h, bytes);
#endif
}
inline void copy_d2h(void *h, const void *d, size_t bytes, stream_t s) {
    if (!bytes) return;
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
#else
    (void)s; memcpy(h, d, bytes);
#endif
}
inline void copy_d2d(void *d, const void *s_, size_t bytes, stream_t s) {
    if (!bytes) return;
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaMemcpyAsync(d, s_, bytes, cudaMemcpyDeviceToDevice, s));
#else
    (void)s; memmove(d, s_, bytes);
#endif
}
inline void dev_memset(void *d, int v, size_t bytes, stream_t s) {
    if (!bytes) return;
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaMemsetAsync(d, v, bytes, s));
#else
    (void)s; memset(d, v, bytes);
#endif
}
inline void stream_sync(stream_t s) {
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaStreamSynchronize(s));
#else
    (void)s;
#endif
}

// lane_b waits (on the device) for everything issued so far on lane_a
inline void stream_wait(stream_t waiter, stream_t producer) {
#ifndef G16_EMU
    cudaEvent_t ev;
    G16_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    G16_CUDA_CHECK(cudaEventRecord(ev, producer));
    G16_CUDA_CHECK(cudaStreamWaitEvent(waiter, ev, 0));
    G16_CUDA_CHECK(cudaEventDestroy(ev));   // released once the wait has been satisfied
#else
    (void)waiter; (void)producer;
#endif
}

// explicit event handles for waits that are queued later than the point they refer to
#ifndef G16_EMU
typedef cudaEvent_t event_t;
#else
typedef void *event_t;
#endif
inline event_t event_record(stream_t producer) {
#ifndef G16_EMU
    cudaEvent_t ev;
    G16_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, producer);
    if (e != cudaSuccess) { cudaEventDestroy(ev); G16_CUDA_CHECK(e); }
    return ev;
#else
    (void)producer; return nullptr;
#endif
}
inline void event_wait_and_release(stream_t waiter, event_t ev) {
#ifndef G16_EMU
    cudaError_t e = cudaStreamWaitEvent(waiter, ev, 0);
    cudaEventDestroy(ev);   // released once the wait has been satisfied
    G16_CUDA_CHECK(e);
#else
    (void)waiter; (void)ev;
#endif
}

// the same across devices: the event is created and recorded on the producer's device, the wait is queued on the
// waiter's; the current device is left at the waiter's
inline void stream_wait_xdev(stream_t waiter, int waiter_dev, stream_t producer, int producer_dev) {
#ifndef G16_EMU
    cudaEvent_t ev;
    G16_CUDA_CHECK(cudaSetDevice(producer_dev));
    G16_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    G16_CUDA_CHECK(cudaEventRecord(ev, producer));
    G16_CUDA_CHECK(cudaSetDevice(waiter_dev));
    G16_CUDA_CHECK(cudaStreamWaitEvent(waiter, ev, 0));
    G16_CUDA_CHECK(cudaEventDestroy(ev));
#else
    (void)waiter; (void)waiter_dev; (void)producer; (void)producer_dev;
#endif
}
// copy between two devices of one process (unified addressing picks the route), queued on the source's stream
inline void copy_peer(void *dst, const void *src, size_t bytes, stream_t s) {
    if (!bytes) return;
#ifndef G16_EMU
    G16_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
#else
    (void)s; memmove(dst, src, bytes);
#endif
}

// atomic add usable from kernel bodies
G16_HD uint32_t atomic_add_u32(uint32_t *p, uint32_t v) {
#if G16_DEVICE_CODE
    return atomicAdd(p, v);
#else
    uint32_t o = *p; *p = o + v; return o;
#endif
}

void note_launch();  // defined in k_misc.cu (one counter for all translation units)
// number of kernels launched by this library since load (bench.py reports the per-step delta)
unsigned long long launch_count();

// ---- launch of a per-thread body -------------------------------------------------------
// A body may declare `static constexpr int MIN_BLOCKS` (resident blocks per SM the register allocation must allow,
// the second argument of __launch_bounds__); bodies without it get 1.
template <class Body, class = void> struct MinBlocksOf { static constexpr int value = 1; };
template <class Body> struct MinBlocksOf<Body, decltype((void)Body::MIN_BLOCKS)> { static constexpr int value = Body::MIN_BLOCKS; };
#ifndef G16_EMU
template <class Body, class... A>
__global__ void __launch_bounds__(Body::BLOCK, MinBlocksOf<Body>::value) thread_kernel(size_t n, A... args) {
    size_t t = (size_t)blockIdx.x * Body::BLOCK + threadIdx.x;
    if (t < n) Body::run(t, args...);
}
template <class Body, class... A>
inline void launch(size_t n, stream_t s, A... args) {
    if (n == 0) return;
    size_t blocks = (n + Body::BLOCK - 1) / Body::BLOCK;
    thread_kernel<Body, A...><<<(unsigned)blocks, Body::BLOCK, 0, s>>>(n, args...);
    G16_CUDA_CHECK(cudaGetLastError());
    note_launch();
}
#else
template <class Body, class... A>
inline void launch(size_t n, stream_t, A... args) {
    for (size_t t = 0; t < n; ++t) Body::run(t, args...);
}
#endif

// grow-only device buffer (workspace cache: no cudaMalloc on the steady-state path)
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void *need(size_t bytes) {
        if (bytes > cap) {
            dev_free(p); p = nullptr; cap = 0;
            size_t want = bytes + bytes / 8;
            p = dev_alloc(want); cap = want;
        }
        return p;
    }
    void release() { dev_free(p); p = nullptr; cap = 0; }
    template <class T> T *as(size_t count) { return (T *)need(count * sizeof(T)); }
};

}  // namespace g16
