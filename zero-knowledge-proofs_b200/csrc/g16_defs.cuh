// Common definitions for the groth16-cuda engine (sm_100a only).
//
// Every arithmetic routine is `__host__ __device__`: the device path is inline PTX
// (carry-chain IMAD), the host path is a bit-exact C emulation of the same PTX
// primitives.  The host path exists ONLY so that tests/emu can single-step the kernel
// bodies in this GPU-less build container; the shipped library (libg16cuda.so) never
// executes it -- all product entry points launch CUDA kernels and fail if no device exists.
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(__CUDACC__)
#define G16_HD __host__ __device__ __forceinline__
#define G16_D __device__ __forceinline__
#else
// host emulation build (tests/emu): plain `inline` keeps g++ compile time sane
#define G16_HD inline
#define G16_D inline
#endif

#if defined(__CUDA_ARCH__)
#define G16_DEVICE_CODE 1
#else
#define G16_DEVICE_CODE 0
#endif

// Field multiplication is force-inlined in the hot kernels and an out-of-line call in the cold
// ones (G16_COLD translation units: reduction, combine, fixed-base tables) -- this keeps ptxas
// compile time and code size in check without touching the hot path.
#if defined(__CUDACC__) && defined(G16_COLD)
#define G16_MUL_HD __host__ __device__ __noinline__
#elif defined(__CUDACC__)
#define G16_MUL_HD G16_HD
#else
#define G16_MUL_HD inline __attribute__((noinline))
#endif

// G2 kernels off the hot path (G16_COLD_FQ2 translation units: bucket reduction): the Fq multiplication stays inline
// but the Fq2 multiplication / squaring (3 / 2 Fq multiplications, ~1000 instructions) become calls -- one copy of the
// code instead of 40 per point addition, and the call overhead is spread over a thousand instructions instead of the
// three hundred of an out-of-line Fq multiplication.
#if defined(__CUDACC__) && defined(G16_COLD_FQ2)
#define G16_FQ2_MUL_HD __host__ __device__ __noinline__
#else
#define G16_FQ2_MUL_HD G16_HD
#endif

namespace g16 {

// ---------------------------------------------------------------------------------------
// PTX carry-chain primitives.  Device: one PTX instruction each (CC = the PTX condition
// code register; ptxas maps it onto predicate registers, so independent chains can still
// be interleaved in SASS).  Host: emulation with an explicit carry variable.
// ---------------------------------------------------------------------------------------
#if !G16_DEVICE_CODE
struct EmuCC { static uint32_t &cc() { static thread_local uint32_t c = 0; return c; } };
#endif

G16_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a + b; EmuCC::cc() = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
G16_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a + b + EmuCC::cc(); EmuCC::cc() = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
G16_HD uint32_t addc(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return (uint32_t)((uint64_t)a + b + EmuCC::cc());
#endif
}
G16_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    // PTX: CC.CF is the borrow-out (1 when a < b)
    uint64_t t = (uint64_t)a - b; EmuCC::cc() = (uint32_t)(t >> 63); return (uint32_t)t;
#endif
}
G16_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a - b - EmuCC::cc(); EmuCC::cc() = (uint32_t)(t >> 63); return (uint32_t)t;
#endif
}
G16_HD uint32_t subc(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return (uint32_t)((uint64_t)a - b - EmuCC::cc());
#endif
}
G16_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
G16_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if G16_DEVICE_CODE
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
G16_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + c; EmuCC::cc() = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
G16_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + c + EmuCC::cc(); EmuCC::cc() = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
G16_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (((uint64_t)a * b) >> 32) + c; EmuCC::cc() = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
G16_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (((uint64_t)a * b) >> 32) + c + EmuCC::cc(); EmuCC::cc() = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
G16_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#if G16_DEVICE_CODE
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return (uint32_t)((((uint64_t)a * b) >> 32) + c + EmuCC::cc());
#endif
}

}  // namespace g16
