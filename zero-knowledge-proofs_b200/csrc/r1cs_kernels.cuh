// Sparse R1CS kernels over Fr -- SURVEY.md 8(f) rank 3: the O(nnz) replacement of the reference's dense
// R1CS -> QAP path, on both sides of the MSM engine.
//
// The reference materialises num_constraints x num_variables dense matrices and num_variables dense
// polynomials (`QAP::from_r1cs`, /root/reference/crates/groth16-qap/src/lib.rs:95-187), so neither setup nor
// prove can reach 2^20 constraints.  Because A_j(x) is the interpolant of column j of the constraint matrix
// on ark-poly's radix-2 domain {omega^i}:
//   setup : A_j(s) = sum_i A[i][j] L_i(s),  L_i(s) = (s^n - 1)/n * omega^i / (s - omega^i)
//           -> LagrangeAt + one sparse matrix-vector product by COLUMNS            (replaces
//           `qap.a_polys.par_iter().map(|p| p.evaluate(&s))`, crates/groth16-setup/src/lib.rs:174-182)
//   prove : A(omega^i) = <A-row i, assignment>  -> one sparse matrix-vector product by ROWS; these are the
//           domain evaluations g16_quotient_h starts from (replaces the dense sums of
//           `compute_quotient_polynomial`, crates/groth16-qap/src/lib.rs:236-257)
// plus the scalar side of `CRS::generate_from_qap` (crates/groth16-setup/src/lib.rs:155-241): every
// "F -> Fr" conversion of the reference keeps the low 64-bit limb only (Truncate64).
//
// Matrices are stored as stacked CSR: the three matrices A, B, C back to back (3 * rows "lines"), one
// line_ptr array of global offsets, one index array, one coefficient array (Fr, Montgomery form).
#pragma once
#include "ntt_kernels.cuh"

namespace g16 {

constexpr uint32_t SPMV_LONG = 256;   // lines with more entries are summed by a whole block

// out[(t / seg) * seg_out + t % seg] = sum_k val[k] * vec[idx[k]] over line t (lines longer than SPMV_LONG
// are left to SpmvLong)
struct SpmvThread {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t t, const uint32_t *line_ptr, const uint32_t *idx, const uint32_t *val, const uint32_t *vec,
                           uint32_t seg, uint32_t seg_out, uint32_t *out) {
        uint32_t b = line_ptr[t], e = line_ptr[t + 1];
        if (e - b > SPMV_LONG) return;
        Fr acc = Fr::zero();
        for (uint32_t k = b; k < e; ++k) acc = Fr::add(acc, Fr::mul(fr_load(val, k), fr_load(vec, idx[k])));
        fr_store(out, (size_t)(t / seg) * seg_out + t % seg, acc);
    }
};

// one block per long line (GPU); serial in the host emulation build
#if !defined(G16_EMU) && defined(__CUDACC__)
constexpr int SPMV_LONG_THREADS = 256;
__global__ void __launch_bounds__(SPMV_LONG_THREADS) spmv_long_kernel(const uint32_t *long_lines, const uint32_t *line_ptr,
                                                                      const uint32_t *idx, const uint32_t *val, const uint32_t *vec,
                                                                      uint32_t seg, uint32_t seg_out, uint32_t *out) {
    __shared__ uint32_t sm[SPMV_LONG_THREADS * 8];
    const uint32_t t = long_lines[blockIdx.x];
    const uint32_t b = line_ptr[t], e = line_ptr[t + 1];
    const int j = threadIdx.x;
    Fr acc = Fr::zero();
    for (uint32_t k = b + j; k < e; k += SPMV_LONG_THREADS) acc = Fr::add(acc, Fr::mul(fr_load(val, k), fr_load(vec, idx[k])));
    for (int d = SPMV_LONG_THREADS >> 1; d >= 1; d >>= 1) {
#pragma unroll
        for (int w = 0; w < 8; ++w) sm[w * SPMV_LONG_THREADS + j] = acc.l[w];
        __syncthreads();
        if (j < d) {
            Fr o;
#pragma unroll
            for (int w = 0; w < 8; ++w) o.l[w] = sm[w * SPMV_LONG_THREADS + j + d];
            acc = Fr::add(acc, o);
        }
        __syncthreads();
    }
    if (j == 0) fr_store(out, (size_t)(t / seg) * seg_out + t % seg, acc);
}
#endif
struct SpmvLongSerial {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t i, const uint32_t *long_lines, const uint32_t *line_ptr, const uint32_t *idx, const uint32_t *val,
                           const uint32_t *vec, uint32_t seg, uint32_t seg_out, uint32_t *out) {
        uint32_t t = long_lines[i];
        Fr acc = Fr::zero();
        for (uint32_t k = line_ptr[t]; k < line_ptr[t + 1]; ++k) acc = Fr::add(acc, Fr::mul(fr_load(val, k), fr_load(vec, idx[k])));
        fr_store(out, (size_t)(t / seg) * seg_out + t % seg, acc);
    }
};

// `Fr::from(x.into_bigint().as_ref()[0])`: keep the low 64-bit limb (Montgomery in, Montgomery out)
G16_HD Fr fr_truncate64(const Fr &x) {
    Fr c = Fr::from_mont(x);
#pragma unroll
    for (int k = 2; k < 8; ++k) c.l[k] = 0;
    return Fr::to_mont(c);
}
struct Truncate64 {
    static constexpr int BLOCK = 256;
    G16_HD static void run(size_t i, const uint32_t *in, uint32_t *out) { fr_store(out, i, fr_truncate64(fr_load(in, i))); }
};

// Scalar block of a setup (Fr, Montgomery), filled by SetupScalars from the five SetupParams:
//   0 alpha_t  1 beta_t  2 gamma_t  3 delta_t  4 s_t   (truncated, setup/src/lib.rs:155-159)
//   5 1/gamma_t  6 1/delta_t  7 (s_t^n - 1)/n  8 flags word (bit 0: gamma_t == 0, bit 1: delta_t == 0)
constexpr uint32_t SETUP_SCALARS = 9;
struct SetupScalars {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *params /* alpha, beta, gamma, delta, s */, const uint32_t *consts,
                           uint32_t log_n, uint32_t truncate, uint32_t *blk) {
        Fr t[5];
        for (int k = 0; k < 5; ++k) {
            t[k] = fr_load(params, k);
            if (truncate) t[k] = fr_truncate64(t[k]);
            fr_store(blk, k, t[k]);
        }
        fr_store(blk, 5, Fr::inv(t[2]));
        fr_store(blk, 6, Fr::inv(t[3]));
        Fr sn = t[4];
        for (uint32_t k = 0; k < log_n; ++k) sn = Fr::sqr(sn);
        fr_store(blk, 7, Fr::mul(Fr::sub(sn, Fr::one()), fr_load(consts, 128)));
        Fr fl = Fr::zero();
        fl.l[0] = (t[2].is_zero() ? 1u : 0u) | (t[3].is_zero() ? 2u : 0u);
        fr_store(blk, 8, fl);
    }
};

// L_i(s) for the domain of size n (consts = the NTT constant block of that size: omega^(2^j), 1/n)
struct LagrangeAt {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, const uint32_t *consts, const uint32_t *blk, uint32_t *out) {
        Fr w = pow_from_table(consts, 0, (uint32_t)i);
        Fr d = Fr::sub(fr_load(blk, 4), w);
        // s on the domain: Z(s) = 0 and L_i(s) is the indicator of s == omega^i
        Fr l = d.is_zero() ? Fr::one() : Fr::mul(Fr::mul(fr_load(blk, 7), w), Fr::inv(d));
        fr_store(out, i, l);
    }
};

// CRS exponents (setup/src/lib.rs:185-241), all truncated to 64 bits like the reference does before the
// scalar multiplication.  vals = a_vals | b_vals | c_vals (num_vars each).  Outputs (Montgomery):
//   ab[j]           = t64(a_vals[j]), ab[num_vars + j] = t64(b_vals[j])                      j < num_vars
//   ic[j]           = t64((beta a_j + alpha b_j + c_j) / (j <= num_public ? gamma : delta))   j < num_vars
//                     (the first num_public + 1 entries are the verification key's, the rest the proving key's)
struct CrsExponents {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t j, const uint32_t *vals, const uint32_t *blk, uint32_t num_vars, uint32_t num_public,
                           uint32_t *ab, uint32_t *ic) {
        Fr a = fr_load(vals, j), b = fr_load(vals, (size_t)num_vars + j), c = fr_load(vals, 2 * (size_t)num_vars + j);
        fr_store(ab, j, fr_truncate64(a));
        fr_store(ab, (size_t)num_vars + j, fr_truncate64(b));
        Fr term = Fr::add(Fr::add(Fr::mul(fr_load(blk, 1), a), Fr::mul(fr_load(blk, 0), b)), c);
        Fr scaled = Fr::mul(term, fr_load(blk, j <= num_public ? 5 : 6));
        fr_store(ic, j, fr_truncate64(scaled));
    }
};
// h[i] = t64(s^i / delta)  (no Z(s): the reference's H query, setup/src/lib.rs:231-241)
struct CrsHExponents {
    static constexpr int BLOCK = 128;
    G16_HD static void run(size_t i, const uint32_t *blk, uint32_t *h) {
        Fr s = fr_load(blk, 4), acc = Fr::one();
        for (uint32_t k = (uint32_t)i; k; k >>= 1) {
            if (k & 1u) acc = Fr::mul(acc, s);
            s = Fr::sqr(s);
        }
        fr_store(h, i, fr_truncate64(Fr::mul(acc, fr_load(blk, 6))));
    }
};

// Witness::validate (core/src/lib.rs:112-131) evaluates at omega = omega^1 only: row (1 mod n) must satisfy
// a * b == c.  flag[1] is set when it does not (flag[0] counts all violated rows, NttCheckVanish).
struct ValidateRow {
    static constexpr int BLOCK = 32;
    G16_HD static void run(size_t, const uint32_t *abc, uint32_t n, uint32_t *flag) {
        size_t i = n > 1 ? 1 : 0;
        Fr a = fr_load(abc, i), b = fr_load(abc, (size_t)n + i), c = fr_load(abc, 2 * (size_t)n + i);
        if (Fr::mul(a, b) != c) flag[1] = 1u;
    }
};

}  // namespace g16
