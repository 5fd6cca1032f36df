// Quotient polynomial H = (A*B - C) / Z on the GPU -- SURVEY.md 8(f) rank 2, the stage that feeds the H MSM.
//
// Replaces `QAP::compute_quotient_polynomial` (/root/reference/crates/groth16-qap/src/lib.rs:225-271, called
// from Prover::prove at crates/groth16-core/src/lib.rs:200): the reference sums dense per-variable polynomials
// (Theta(N*n)), multiplies by FFT and divides by Z(x) = x^n - 1.  Given the evaluations of A, B, C on the
// radix-2 domain (a_i = <A-row i, w>, which a sparse R1CS yields in O(nnz)), the same H is obtained with
// seven size-n NTTs over Fr:
//     a, b, c  --iNTT-->  coefficients  --x g^j, NTT-->  values on the coset g*<omega>
//     h~_i = (a~_i b~_i - c~_i) / (g^n - 1)            (Z is the constant g^n - 1 on the coset)
//     h~  --iNTT, x g^-j-->  coefficients of H  (degree <= n - 2, so n coset points determine it)
// Domain = ark-poly's Radix2EvaluationDomain: omega = TWO_ADIC_ROOT^(2^(32 - log n)), TWO_ADIC_ROOT =
// 7^((r-1)/2^32); coset generator g = 7 (any g outside the domain gives the same H).
//
// Transforms: decimation-in-frequency (natural in, bit-reversed out) and decimation-in-time (bit-reversed in, natural
// out), so no separate permutation pass is needed.  On the device up to seven consecutive stages run on a tile held in
// shared memory (ntt_fused_kernel); the one-stage kernels state the same butterflies for the host emulation build.
#pragma once
#include "fp.cuh"
#include "kernel_api.cuh"

namespace g16 {

G16_HD Fr fr_load(const uint32_t *p, size_t i) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = p[8 * i + k];
    return r;
}
G16_HD void fr_store(uint32_t *p, size_t i, const Fr &v) {
#pragma unroll
    for (int k = 0; k < 8; ++k) p[8 * i + k] = v.l[k];
}
G16_HD uint32_t bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t b = 0; b < bits; ++b) { r = (r << 1) | (x & 1u); x >>= 1; }
    return r;
}

// Layout of the constant block (Fr elements, Montgomery form):
//   [0, 32)   w2[j]  = omega^(2^j)          [32, 64)  wi2[j] = omega^-(2^j)
//   [64, 96)  g2[j]  = g^(2^j)              [96, 128) gi2[j] = g^-(2^j)
//   128: n^-1      129: (g^n - 1)^-1
constexpr uint32_t NTT_CONST_WORDS = 130 * 8;
// Fr(7) and TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32) in Montgomery form (ark-bls12-381 Fr)
G16_HD Fr fr_small(uint32_t v) {
    Fr x = Fr::zero();
    x.l[0] = v;
    return Fr::to_mont(x);
}

struct NttSetup {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, uint32_t log_n, uint32_t *consts) {
        // TWO_ADIC_ROOT = 7^((r-1) >> 32): exponent = (r - 1) / 2^32, square-and-multiply over its 223 bits
        Fr g = fr_small(7);
        uint32_t e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = FrParams::MOD(k);
        e[0] -= 1u;   // r - 1 (r is odd, no borrow)
        Fr root = Fr::one();
        for (int bit = 255; bit >= 32; --bit) {
            root = Fr::sqr(root);
            if ((e[bit >> 5] >> (bit & 31)) & 1u) root = Fr::mul(root, g);
        }
        // omega = root^(2^(32 - log_n))
        Fr w = root;
        for (uint32_t k = log_n; k < 32; ++k) w = Fr::sqr(w);
        Fr wi = Fr::inv(w), gi = Fr::inv(g);
        Fr a = w, b = wi, c = g, d = gi;
        for (uint32_t j = 0; j < 32; ++j) {
            fr_store(consts, j, a); fr_store(consts, 32 + j, b); fr_store(consts, 64 + j, c); fr_store(consts, 96 + j, d);
            a = Fr::sqr(a); b = Fr::sqr(b); c = Fr::sqr(c); d = Fr::sqr(d);
        }
        // n^-1 and (g^n - 1)^-1 ;  g^n = g2[log_n]
        Fr n_fr = Fr::zero();
        n_fr.l[log_n >> 5] = 1u << (log_n & 31);
        fr_store(consts, 128, Fr::inv(Fr::to_mont(n_fr)));
        Fr gn = fr_load(consts, 64 + log_n);
        fr_store(consts, 129, Fr::inv(Fr::sub(gn, Fr::one())));
    }
};

// base^k from the table of base^(2^j)
G16_HD Fr pow_from_table(const uint32_t *consts, uint32_t table, uint32_t k) {
    Fr acc = Fr::one();
    for (uint32_t j = 0; k; ++j, k >>= 1)
        if (k & 1u) acc = Fr::mul(acc, fr_load(consts, table + j));
    return acc;
}

// tw[k] = omega^k, twi[k] = omega^-k for k < n/2
struct NttTwiddles {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t k, const uint32_t *consts, uint32_t *tw, uint32_t *twi) {
        fr_store(tw, k, pow_from_table(consts, 0, (uint32_t)k));
        fr_store(twi, k, pow_from_table(consts, 32, (uint32_t)k));
    }
};

// one DIF stage on `batch` arrays of n elements stored back to back (half = distance of the pair)
struct NttStageDif {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t t, uint32_t *x, const uint32_t *tw, uint32_t n, uint32_t half) {
        size_t arr = t / (n / 2);
        uint32_t q = (uint32_t)(t % (n / 2));
        uint32_t j = q % half, blk = q / half;
        size_t i = arr * n + (size_t)blk * 2 * half + j;
        Fr u = fr_load(x, i), v = fr_load(x, i + half);
        fr_store(x, i, Fr::add(u, v));
        fr_store(x, i + half, Fr::mul(Fr::sub(u, v), fr_load(tw, (size_t)j * (n / (2 * half)))));
    }
};
struct NttStageDit {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t t, uint32_t *x, const uint32_t *tw, uint32_t n, uint32_t half) {
        size_t arr = t / (n / 2);
        uint32_t q = (uint32_t)(t % (n / 2));
        uint32_t j = q % half, blk = q / half;
        size_t i = arr * n + (size_t)blk * 2 * half + j;
        Fr u = fr_load(x, i), v = Fr::mul(fr_load(x, i + half), fr_load(tw, (size_t)j * (n / (2 * half))));
        fr_store(x, i, Fr::add(u, v));
        fr_store(x, i + half, Fr::sub(u, v));
    }
};

// ---- fused stages (device build) ---------------------------------------------------------------------------------
// The kernels above stream the whole array through HBM once per stage (64 B per butterfly per stage).  Here a block
// keeps a tile of 2^S rows x C columns in shared memory and runs S consecutive stages on it, so a transform of 2^20
// points makes 3 passes over HBM instead of 20.  Same butterflies, same twiddle table, same element order as the
// one-stage kernels -- the results are bit-identical.  Stage k (k = 0 .. log_n - 1) has half = n >> (k + 1); a pass
// covers stages k0 .. k0 + S - 1, whose butterfly network couples the 2^S elements
//     i = B * 2^S * stride + r * stride + off,   r < 2^S,   stride = n >> (k0 + S),   off < stride, B < 2^k0.
// Middle passes (stride >= C) take C consecutive offsets as columns (C * 32 B contiguous per row); the innermost pass
// (stride = 1) takes C consecutive blocks B as columns (the tile is 2^S * C contiguous elements).
#if !defined(G16_EMU) && defined(__CUDACC__)
constexpr int NTT_FUSED_THREADS = 256;
constexpr uint32_t NTT_FUSED_MAX_STAGES = 7;
constexpr uint32_t NTT_FUSED_MAX_COLS = 8;
__device__ __forceinline__ Fr fr_load4(const uint32_t *p) {
    Fr r;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void fr_store4(uint32_t *p, const Fr &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// x: `batch` arrays of n elements back to back; tw: omega^k (DIT, forward) or omega^-k (DIF, inverse), k < n / 2.
// scale (optional, DIF innermost pass or DIT first pass): per-element factors applied when the tile is stored (DIF) /
// loaded (DIT) -- the coset shift g^(br(p)) / n of the bit-reversed coefficients, see NttCosetTable.
static __global__ void __launch_bounds__(NTT_FUSED_THREADS) ntt_fused_kernel(uint32_t *x, const uint32_t *tw, uint32_t log_n, uint32_t k0,
                                                                             uint32_t S, uint32_t log_c, uint32_t tiles_per_array,
                                                                             int dit, const uint32_t *scale) {
    extern __shared__ uint4 ntt_tile4[];
    uint32_t *tile = reinterpret_cast<uint32_t *>(ntt_tile4);
    const uint32_t n = 1u << log_n, R = 1u << S, C = 1u << log_c;
    const uint32_t log_stride = log_n - k0 - S, stride = 1u << log_stride;
    const uint32_t arr = blockIdx.x / tiles_per_array, t = blockIdx.x % tiles_per_array;
    uint32_t *xa = x + (size_t)arr * n * 8;
    const bool inner = log_stride == 0;
    // element (r, c) of the tile lives at base(c) + r * stride
    uint32_t base0, off0 = 0;
    if (inner) base0 = (t << log_c) << S;                                   // block B = t C + c, index B R + r
    else {
        uint32_t per = stride >> log_c;                                     // tiles per block B
        uint32_t B = t / per;
        off0 = (t % per) << log_c;
        base0 = ((B << S) << log_stride) + off0;
    }
    const uint32_t elems = R << log_c;
    // load: consecutive threads walk the contiguous direction (rows when inner, columns otherwise)
    for (uint32_t e = threadIdx.x; e < elems; e += NTT_FUSED_THREADS) {
        uint32_t r, c;
        size_t gi;
        if (inner) { r = e & (R - 1); c = e >> S; gi = (size_t)base0 + ((size_t)c << S) + r; }
        else { c = e & (C - 1); r = e >> log_c; gi = (size_t)base0 + ((size_t)r << log_stride) + c; }
        Fr v = fr_load4(xa + gi * 8);
        if (dit && scale) v = Fr::mul(v, fr_load4(scale + gi * 8));
        fr_store4(tile + ((size_t)(r << log_c) + c) * 8, v);
    }
    __syncthreads();
    const uint32_t bfs = elems >> 1;
    for (uint32_t s = 0; s < S; ++s) {
        // DIF walks the local half R/2 .. 1, DIT 1 .. R/2
        const uint32_t log_lh = dit ? s : S - 1 - s, lh = 1u << log_lh;
        const uint32_t log_h = log_lh + log_stride;                         // global half of this stage
        const uint32_t tw_shift = log_n - 1 - log_h;                        // twiddle index = j << tw_shift
        for (uint32_t bf = threadIdx.x; bf < bfs; bf += NTT_FUSED_THREADS) {
            uint32_t c = bf & (C - 1), pr = bf >> log_c;
            uint32_t jr = pr & (lh - 1), r0 = ((pr >> log_lh) << (log_lh + 1)) + jr;
            uint32_t j = inner ? jr : (jr << log_stride) + off0 + c;
            uint32_t *p0 = tile + ((size_t)(r0 << log_c) + c) * 8, *p1 = p0 + ((size_t)lh << log_c) * 8;
            Fr u = fr_load4(p0), v = fr_load4(p1), w = fr_load4(tw + ((size_t)j << tw_shift) * 8);
            if (dit) {
                v = Fr::mul(v, w);
                fr_store4(p0, Fr::add(u, v));
                fr_store4(p1, Fr::sub(u, v));
            } else {
                fr_store4(p0, Fr::add(u, v));
                fr_store4(p1, Fr::mul(Fr::sub(u, v), w));
            }
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < elems; e += NTT_FUSED_THREADS) {
        uint32_t r, c;
        size_t gi;
        if (inner) { r = e & (R - 1); c = e >> S; gi = (size_t)base0 + ((size_t)c << S) + r; }
        else { c = e & (C - 1); r = e >> log_c; gi = (size_t)base0 + ((size_t)r << log_stride) + c; }
        Fr v = fr_load4(tile + ((size_t)(r << log_c) + c) * 8);
        if (!dit && scale) v = Fr::mul(v, fr_load4(scale + gi * 8));
        fr_store4(xa + gi * 8, v);
    }
}
#endif

// scale[p] = g^(br(p)) / n for p < n: the factor the coset shift applies to the bit-reversed coefficient p.  Built once
// per domain size (cached with the twiddles), so the shift costs one multiplication per element inside the fused pass
// instead of a ~log n multiplication power ladder per element per prove.
struct NttCosetTable {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t p, const uint32_t *consts, uint32_t log_n, uint32_t *scale) {
        fr_store(scale, p, Fr::mul(pow_from_table(consts, 64, bitrev((uint32_t)p, log_n)), fr_load(consts, 128)));
    }
};
// x[t] *= scale[t mod n] (the emulation build and n = 1 apply the table with a kernel of its own)
struct NttCosetScaleTable {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t, uint32_t *x, const uint32_t *scale, uint32_t n) {
        fr_store(x, t, Fr::mul(fr_load(x, t), fr_load(scale, t % n)));
    }
};
// fscale[j] = g^-j / n for j < n (the factor of output coefficient j)
struct NttFinalTable {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t j, const uint32_t *consts, uint32_t *fscale) {
        fr_store(fscale, j, Fr::mul(pow_from_table(consts, 96, (uint32_t)j), fr_load(consts, 128)));
    }
};
// out[j] = x[br(j)] * fscale[j]
struct NttFinalPermute {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t j, const uint32_t *x, const uint32_t *fscale, uint32_t log_n, uint32_t *out) {
        fr_store(out, j, Fr::mul(fr_load(x, bitrev((uint32_t)j, log_n)), fr_load(fscale, j)));
    }
};

// h~ = (a~ b~ - c~) / (g^n - 1) into a~   (a, b, c stored back to back)
struct NttQuotientPointwise {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, uint32_t *abc, const uint32_t *consts, uint32_t n) {
        Fr a = fr_load(abc, i), b = fr_load(abc, (size_t)n + i), c = fr_load(abc, 2 * (size_t)n + i);
        fr_store(abc, i, Fr::mul(Fr::sub(Fr::mul(a, b), c), fr_load(consts, 129)));
    }
};

// the reference fails with PolynomialDivisionFailed unless A*B - C vanishes on the domain
struct NttCheckVanish {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, const uint32_t *abc, uint32_t n, uint32_t *flag) {
        Fr a = fr_load(abc, i), b = fr_load(abc, (size_t)n + i), c = fr_load(abc, 2 * (size_t)n + i);
        if (Fr::mul(a, b) != c) atomic_add_u32(flag, 1u);   // Montgomery forms: mont(a) * mont(b) = mont(ab)
    }
};

}  // namespace g16
