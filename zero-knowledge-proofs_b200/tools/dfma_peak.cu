// FP64 pipe microbenchmark for sm_100a: is the double-precision FMA pipe a second multiplier for big-integer work?
// (Emmart et al.: a 52 x 52 -> 104 bit product costs two DFMAs, hi = fma_rz(a, b, 0), lo = fma(a, b, -hi).)
//   dfma        : fma.rz.f64, 8 independent chains per thread
//   imad_wide   : mad.wide.u32, 8 independent chains per thread (same as imad_peak)
//   mixed       : 4 DFMA chains + 4 IMAD.WIDE chains interleaved in one thread -- do the two pipes overlap?
// Prints one JSON object.  Usage: dfma_peak [scale]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int INNER = 64;

__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b, int iters) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 0.12345) out[0] = s;
}
__global__ void __launch_bounds__(256) k_wide(double *out, uint32_t a, uint32_t b, int iters) {
    unsigned long long x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((uint32_t)x[i]), "r"(b + a));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345678ull) out[0] = (double)s;
}
__global__ void __launch_bounds__(256) k_mixed(double *out, double a, double b, uint32_t ia, uint32_t ib, int iters) {
    double x[4];
    unsigned long long y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = threadIdx.x + i; y[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[i]) : "r"((uint32_t)y[i]), "r"(ib + ia));
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += x[i] + (double)y[i];
    if (s == 0.12345) out[0] = s;
}

// The inner step of a double-precision-limb multiplication as Emmart et al. formulate it: per 52 x 52 limb product
//   hi = fma_rz(a, b, 2^104);  s = (2^104 + 2^52) - hi;  lo = fma_rz(a, b, s)
// and the two results are accumulated as raw 64-bit integers (the exponent bits are constant and come off at the end).
// 2 DFMA + 1 DADD + 2 integer 64-bit adds per product; a 381-bit Montgomery multiplication needs 128 of them.
__global__ void __launch_bounds__(256) k_dpf_product(double *out, double a0, double b0, int iters) {
    const double C1 = 20282409603651670423947251286016.0;                    // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;   // 2^104 + 2^52
    double a[4], b[4];
    long long acc_hi[4], acc_lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = a0 + threadIdx.x + i; b[i] = b0 + i; acc_hi[i] = 0; acc_lo[i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double hi, sdiff, lo;
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(a[i]), "d"(b[i]), "d"(C1));
                asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(sdiff) : "d"(C2), "d"(hi));
                asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(a[i]), "d"(b[i]), "d"(sdiff));
                asm volatile("add.s64 %0, %0, %1;" : "+l"(acc_hi[i]) : "l"(__double_as_longlong(hi)));
                asm volatile("add.s64 %0, %0, %1;" : "+l"(acc_lo[i]) : "l"(__double_as_longlong(lo)));
            }
    }
    long long s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s ^= acc_hi[i] ^ acc_lo[i];
    if (s == 0x123456789ll) out[0] = (double)s;
}

template <class L>
static double time_ms(L launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main(int argc, char **argv) {
    double scale = argc > 1 ? atof(argv[1]) : 1.0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double *d_out;
    CK(cudaMalloc(&d_out, 64));
    int blocks = sms * 8, threads = 256, iters = (int)(1000 * scale);
    double total = (double)blocks * threads * iters * INNER;
    double ms_d = time_ms([&] { k_dfma<<<blocks, threads>>>(d_out, 1.0000001, 0.5, iters); });
    double ms_w = time_ms([&] { k_wide<<<blocks, threads>>>(d_out, 3u, 5u, iters); });
    double ms_m = time_ms([&] { k_mixed<<<blocks, threads>>>(d_out, 1.0000001, 0.5, 3u, 5u, iters); });
    double ms_p = time_ms([&] { k_dpf_product<<<blocks, threads>>>(d_out, 1125899906842624.0, 2251799813685248.0, iters); });
    double dpf_products = total * 4 / (ms_p * 1e-3);
    printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_per_s\": %.6e, \"imad_wide_per_s\": %.6e, "
           "\"mixed_dfma_per_s\": %.6e, \"mixed_imad_wide_per_s\": %.6e, \"ms\": {\"dfma\": %.3f, \"imad_wide\": %.3f, \"mixed\": %.3f}, "
           "\"bits2_per_s\": {\"dfma_52x52_two_fma\": %.4e, \"imad_wide_32x32\": %.4e}, "
           "\"dpf_limb_products_per_s\": %.6e, \"dpf_ms\": %.3f, \"dpf_fq_mul_per_s_at_128_products\": %.4e}\n",
           prop.name, sms, total * 8 / (ms_d * 1e-3), total * 8 / (ms_w * 1e-3), total * 4 / (ms_m * 1e-3), total * 4 / (ms_m * 1e-3),
           ms_d, ms_w, ms_m, total * 8 / (ms_d * 1e-3) / 2 * 2704, total * 8 / (ms_w * 1e-3) * 1024, dpf_products, ms_p,
           dpf_products / 128.0);
    cudaFree(d_out);
    return 0;
}
