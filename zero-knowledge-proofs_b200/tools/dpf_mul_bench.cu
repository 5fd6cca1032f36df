// Device measurement for the next round (NOT part of the library): BLS12-381 Fq Montgomery multiplication on 8 x 48-bit
// limbs held in doubles (the algorithm of tools/dpf_mul_prototype.cpp) against the engine's IMAD.WIDE Fq::mul -- same
// chained workload as imad_peak's fq_mul (two dependent multiplication chains per thread), plus a device-side
// comparison of the two results.  Prints one JSON object.  Usage: dpf_mul_bench [scale]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../csrc/fp.cuh"

using namespace g16;
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__constant__ double c_q[8];
__constant__ u64 c_ninv;
constexpr u64 MASK48 = (1ull << 48) - 1;
constexpr double C1 = 1267650600228229401496703205376.0;                       // 2^100
constexpr double C2 = 1267650600228229401496703205376.0 + 4503599627370496.0;   // 2^100 + 2^52
constexpr u64 K_HI = 0x4630000000000000ull;   // bit pattern of 2^100
constexpr u64 K_LO = 0x4330000000000000ull;   // bit pattern of 2^52

struct D8 { double l[8]; };

__device__ __forceinline__ void product(double a, double b, u64 &acc_hi, u64 &acc_lo) {
    double hi = __fma_rz(a, b, C1);
    double sub = __dsub_rz(C2, hi);
    double lo = __fma_rz(a, b, sub);
    acc_hi += (u64)__double_as_longlong(hi);
    acc_lo += (u64)__double_as_longlong(lo);
}

__device__ __forceinline__ D8 dpf_mul(const D8 &a, const D8 &b) {
    double m[8];
    u64 carry = 0, prev_hi = 0, out[8];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        u64 acc_hi = 0, acc_lo = 0;
        int terms = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i <= k && k - i < 8) { product(a.l[i], b.l[k - i], acc_hi, acc_lo); ++terms; }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < k && k - i < 8) { product(m[i], c_q[k - i], acc_hi, acc_lo); ++terms; }
        }
        u64 col = (acc_lo - (u64)terms * K_LO) + prev_hi + carry;
        u64 hi_sum = acc_hi - (u64)terms * K_HI;
        if (k < 8) {
            u64 mk = ((col & MASK48) * c_ninv) & MASK48;
            m[k] = __ull2double_rz(mk);
            u64 h2 = 0, l2 = 0;
            product(m[k], c_q[0], h2, l2);
            col += l2 - K_LO;
            hi_sum += h2 - K_HI;
        } else {
            out[k - 8] = col & MASK48;
        }
        carry = col >> 48;
        prev_hi = hi_sum;
    }
    // conditional subtraction of q (result < 2q)
    u64 t[8];
    long long borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long d = (long long)out[i] - (long long)__double2ull_rz(c_q[i]) - borrow;
        borrow = d < 0;
        t[i] = (u64)(d + (borrow ? (1ll << 48) : 0ll)) & MASK48;
    }
    bool ge = !borrow || (prev_hi + carry) != 0;
    D8 r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = __ull2double_rz(ge ? t[i] : out[i]);
    return r;
}

__device__ __forceinline__ D8 to_d8(const Fq &x) {
    D8 r;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int bit = 48 * i, w = bit >> 5, sh = bit & 31;
        u64 lo = x.l[w] | ((u64)(w + 1 < 12 ? x.l[w + 1] : 0u) << 32);
        u64 v = lo >> sh;
        if (sh > 16 && w + 2 < 12) v |= (u64)x.l[w + 2] << (64 - sh);
        r.l[i] = __ull2double_rz(v & MASK48);
    }
    return r;
}
__device__ __forceinline__ Fq from_d8(const D8 &a) {
    Fq r = Fq::zero();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int bit = 48 * i, w = bit >> 5, sh = bit & 31;
        u64 v = __double2ull_rz(a.l[i]);
        r.l[w] |= (uint32_t)(v << sh);
        if (w + 1 < 12) r.l[w + 1] |= (uint32_t)(v >> (32 - sh));
        if (sh > 16 && w + 2 < 12) r.l[w + 2] |= (uint32_t)(v >> (64 - sh));
    }
    return r;
}

__global__ void __launch_bounds__(256) k_dpf(uint32_t *out, int iters, int check) {
    Fq x = Fq::one(), y = Fq::one(), mm;
#pragma unroll
    for (int i = 0; i < 12; ++i) { mm.l[i] = FqParams::R2(i) ^ threadIdx.x; x.l[i] ^= (uint32_t)(i * 7 + threadIdx.x); }
    mm.l[11] &= 0x0fffffffu; x.l[11] &= 0x0fffffffu;
    D8 dx = to_d8(x), dy = to_d8(y), dm = to_d8(mm);
    for (int it = 0; it < iters; ++it) {
        dx = dpf_mul(dx, dm);
        dy = dpf_mul(dy, dx);
    }
    Fq rx = from_d8(dx), ry = from_d8(dy);
    if (check) {
        for (int it = 0; it < iters; ++it) { x = Fq::mul(x, mm); y = Fq::mul(y, x); }
        if (!(rx == x) || !(ry == y)) atomicAdd(&out[1], 1u);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s ^= rx.l[i] ^ ry.l[i];
    if (s == 0x12345678u) out[0] = s;
}
__global__ void __launch_bounds__(256) k_imad_mul(uint32_t *out, int iters) {
    Fq x = Fq::one(), y = Fq::one(), m;
#pragma unroll
    for (int i = 0; i < 12; ++i) { m.l[i] = FqParams::R2(i) ^ threadIdx.x; x.l[i] ^= (uint32_t)(i * 7 + threadIdx.x); }
    for (int it = 0; it < iters; ++it) { x = Fq::mul(x, m); y = Fq::mul(y, x); }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s ^= x.l[i] ^ y.l[i];
    if (s == 0x12345678u) out[0] = s;
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
    // modulus in 48-bit limbs, -q^-1 mod 2^48
    u64 q48[8];
    {
        unsigned __int128 acc = 0;
        int have = 0, w = 0;
        for (int i = 0; i < 8; ++i) {
            while (have < 48 && w < 12) { acc |= (unsigned __int128)FqParams::MOD(w) << have; have += 32; ++w; }
            q48[i] = (u64)acc & MASK48;
            acc >>= 48; have -= 48;
        }
    }
    double qd[8];
    for (int i = 0; i < 8; ++i) qd[i] = (double)q48[i];
    u64 inv = 1;
    for (int k = 0; k < 6; ++k) inv *= 2 - q48[0] * inv;
    u64 ninv = (0 - inv) & MASK48;
    CK(cudaMemcpyToSymbol(c_q, qd, sizeof(qd)));
    CK(cudaMemcpyToSymbol(c_ninv, &ninv, sizeof(ninv)));
    uint32_t *d_out;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMemset(d_out, 0, 64));
    k_dpf<<<8, 256>>>(d_out, 50, 1);          // correctness: 2 x 50 chained multiplications per thread, both ways
    CK(cudaDeviceSynchronize());
    uint32_t h[2];
    CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = (int)(200 * scale);
    double muls = (double)blocks * threads * iters * 2;
    double ms_d = time_ms([&] { k_dpf<<<blocks, threads>>>(d_out, iters, 0); });
    double ms_i = time_ms([&] { k_imad_mul<<<blocks, threads>>>(d_out, iters); });
    printf("{\"device\": \"%s\", \"dpf_mismatching_threads\": %u, \"dpf_fq_mul_per_s\": %.6e, \"imad_fq_mul_per_s\": %.6e, "
           "\"ratio\": %.3f, \"ms\": {\"dpf\": %.3f, \"imad\": %.3f}}\n",
           prop.name, h[1], muls / (ms_d * 1e-3), muls / (ms_i * 1e-3), ms_i / ms_d, ms_d, ms_i);
    return 0;
}
