// Integer-pipe roofline microbenchmark for sm_100a: measures the issue rate of the multiply-add
// forms the field multiplication is built from, plus the field multiplication itself.
//   imad      : mad.lo.u32        (IMAD)            -- the roofline denominator of SURVEY.md 8(d)
//   imad_hi   : mad.hi.u32        (IMAD.HI)
//   imad_wide : mad.wide.u32      (IMAD.WIDE.U32, 32x32+64 -> 64)
//   imad_wide_cc : mad.lo.cc/madc.hi.cc pairs (IMAD.WIDE.U32.X carry chain)
//   fq_mul    : 381-bit Montgomery multiplications/s (this engine's Fq::mul)
// Prints one JSON object.  Usage: imad_peak [seconds_per_variant]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../csrc/fp.cuh"

using namespace g16;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int CHAINS = 8;
constexpr int INNER = 64;

__global__ void __launch_bounds__(256) k_imad(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) k_imad_hi(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) k_imad_wide(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    unsigned long long x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"((uint32_t)x[i]), "r"(b + a));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= x[i];
    if (s == 0x12345678ull) out[0] = (uint32_t)s;
}
// carry chains: 4 independent chains of 4 wide MADs each (lo.cc / madc.hi.cc pairs)
__global__ void __launch_bounds__(256) k_imad_wide_cc(uint32_t *out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[4][8];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[c][i] = threadIdx.x + i + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER / 4; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                x[c][0] = mad_lo_cc(a, b + c, x[c][0]);
                x[c][1] = madc_hi_cc(a, b + c, x[c][1]);
                x[c][2] = madc_lo_cc(a, b + c + 1, x[c][2]);
                x[c][3] = madc_hi_cc(a, b + c + 1, x[c][3]);
                x[c][4] = madc_lo_cc(a, b + c + 2, x[c][4]);
                x[c][5] = madc_hi_cc(a, b + c + 2, x[c][5]);
                x[c][6] = madc_lo_cc(a, b + c + 3, x[c][6]);
                x[c][7] = madc_hi(a, b + c + 3, x[c][7]);
            }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 8; ++i) s ^= x[c][i];
    if (s == 0x12345678u) out[0] = s;
}
// two independent multiplication chains per thread
__global__ void __launch_bounds__(256) k_fq_mul(uint32_t *out, int iters) {
    Fq x = Fq::one(), y = Fq::one(), m;
#pragma unroll
    for (int i = 0; i < 12; ++i) { m.l[i] = FqParams::R2(i) ^ threadIdx.x; x.l[i] ^= (uint32_t)(i * 7 + threadIdx.x); }
    for (int it = 0; it < iters; ++it) {
        x = Fq::mul(x, m);
        y = Fq::mul(y, x);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s ^= x.l[i] ^ y.l[i];
    if (s == 0x12345678u) out[0] = s;
}

template <class L>
static double time_ms(L launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch();  // warm-up
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
    uint32_t *d_out;
    CK(cudaMalloc(&d_out, 64));
    int blocks = sms * 8, threads = 256;
    int iters = (int)(2000 * scale);
    double total_threads = (double)blocks * threads;
    double ops = total_threads * (double)iters * INNER * CHAINS;
    double ms_imad = time_ms([&] { k_imad<<<blocks, threads>>>(d_out, 3u, 5u, iters); });
    double ms_hi = time_ms([&] { k_imad_hi<<<blocks, threads>>>(d_out, 3u, 5u, iters); });
    double ms_wide = time_ms([&] { k_imad_wide<<<blocks, threads>>>(d_out, 3u, 5u, iters); });
    double ops_cc = total_threads * (double)iters * (INNER / 4) * 4 * 4;  // wide MADs
    double ms_cc = time_ms([&] { k_imad_wide_cc<<<blocks, threads>>>(d_out, 3u, 5u, iters); });
    int mul_iters = (int)(400 * scale);
    double muls = total_threads * (double)mul_iters * 2;
    double ms_mul = time_ms([&] { k_fq_mul<<<blocks, threads>>>(d_out, mul_iters); });
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
           "\"imad_per_s\": %.6e, \"imad_hi_per_s\": %.6e, \"imad_wide_per_s\": %.6e, \"imad_wide_cc_per_s\": %.6e, "
           "\"fq_mul_per_s\": %.6e, \"ms\": {\"imad\": %.3f, \"imad_hi\": %.3f, \"imad_wide\": %.3f, \"imad_wide_cc\": %.3f, \"fq_mul\": %.3f}}\n",
           prop.name, sms, clk, ops / (ms_imad * 1e-3), ops / (ms_hi * 1e-3), ops / (ms_wide * 1e-3), ops_cc / (ms_cc * 1e-3),
           muls / (ms_mul * 1e-3), ms_imad, ms_hi, ms_wide, ms_cc, ms_mul);
    cudaFree(d_out);
    return 0;
}
