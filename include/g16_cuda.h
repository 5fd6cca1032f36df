/* groth16-cuda: C ABI of the B200-native MSM engine (libg16cuda.so).
 *
 * This is the drop-in boundary for the Groth16 prover/setup of vats98754/zero-knowledge-proofs.
 * The reference has no FFI today; these entry points are what a `groth16-cuda` Rust crate binds
 * (see INTEGRATION.md and zero-knowledge-proofs_b200/rust/groth16-cuda) to replace:
 *
 *   g16_g1_msm / g16_g1_msm_oneshot   ark `G1Projective::msm(&points,&scalars)` + `.into_affine()`
 *                                     crates/groth16-core/src/lib.rs:282,285 (Prover::multi_scalar_mult_g1, :275-286)
 *   g16_g2_msm / g16_g2_msm_oneshot   `G2Projective::msm` + `.into_affine()`          lib.rs:296,299 (:289-300)
 *   g16_g1_fixed_base_mul             `(g1_gen * fr).into_affine()` per element        crates/groth16-setup/src/lib.rs:166-171,185-191,194-199,210-218,221-229,232-241
 *   g16_g2_fixed_base_mul             `(g2_gen * fr).into_affine()` per element        crates/groth16-setup/src/lib.rs:168-171,201-207
 *   g16_pk_upload / g16_prove         the 4 x G1 + 1 x G2 MSM schedule of Prover::prove crates/groth16-core/src/lib.rs:164-271
 *
 * Data layout = ark-ff 0.4 / ark-ec 0.4 in-memory values, no conversion on the host:
 *   Fr scalar : 4 x u64 little-endian limbs, Montgomery form (a * 2^256 mod r)
 *   Fq        : 6 x u64 little-endian limbs, Montgomery form (a * 2^384 mod q)
 *   G1 affine : 12 x u64 = x[6], y[6];  infinity flag in a separate byte array (ark's identity is
 *               {x:0, y:0, infinity:true}; outputs follow that convention)
 *   G2 affine : 24 x u64 = x.c0[6], x.c1[6], y.c0[6], y.c1[6]; flag as above
 *   Proof     : a (G1), b (G2), c (G1) in that order, as in `struct Proof` lib.rs:27-36
 *
 * All functions return 0 on success and a non-zero G16_ERR_* code otherwise; the message is
 * available from g16_last_error().  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with G16_ERR_NO_DEVICE.  A ctx may be used by one thread at a time.
 * Caller owns all input/output buffers; handles are owned by the library until *_free/_destroy.
 * Lifetime: g16_bases, g16_pk and g16_r1cs handles belong to the context they were created with and must not be
 * USED after g16_ctx_destroy; freeing them afterwards is allowed (they release their own device memory and do not
 * reach through the context), freeing them first is the normal order.
 */
#ifndef G16_CUDA_H
#define G16_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G16_OK 0
#define G16_ERR_INVALID 1   /* bad argument (NULL pointer, wrong group, ...) */
#define G16_ERR_CUDA 2      /* CUDA runtime error */
#define G16_ERR_NO_DEVICE 3 /* no usable CUDA device */
#define G16_ERR_OOM 4       /* device allocation failed */
#define G16_ERR_LENGTH 5    /* bases / scalars length mismatch (ark: Err(min_len), lib.rs:283,297) */

typedef struct g16_ctx g16_ctx;
typedef struct g16_bases g16_bases; /* device-resident base points of one group, sharded over the ctx devices */
typedef struct g16_pk g16_pk;       /* device-resident ProvingKey arrays (crates/groth16-setup/src/lib.rs:27-52) */

/* ---- context ------------------------------------------------------------------------- */
/* devices == NULL or ndev == 0: use the current CUDA device.  With ndev > 1 every bases
 * array is partitioned by index range over the devices and each MSM ends with a partial-sum
 * combine on devices[0]. */
int g16_ctx_create(const int *devices, int ndev, g16_ctx **out);
void g16_ctx_destroy(g16_ctx *ctx);
const char *g16_last_error(const g16_ctx *ctx); /* ctx may be NULL: last error of ctx creation */
/* run single-device work on the caller's cudaStream_t (e.g. torch's current stream) */
int g16_ctx_set_stream(g16_ctx *ctx, void *cuda_stream);
int g16_ctx_synchronize(g16_ctx *ctx);
/* tuning: window bits for the next MSMs (0 = choose from n) */
int g16_ctx_set_window_bits(g16_ctx *ctx, unsigned c);
/* tuning: host-scalar MSMs (g16_*_msm, g16_*_msm_oneshot, g16_*_msm_async) of at least min_scalars scalars per
 * device copy their scalars in three index ranges on a second stream; every range is sorted and accumulated into
 * the same bucket array as soon as it has arrived, so only the first small copy is exposed.  0 restores the default
 * (2^19).  Results are unchanged. */
int g16_ctx_set_h2d_pipeline_min(g16_ctx *ctx, size_t min_scalars);
/* tuning: the longest run of additions one thread walks serially in the bucket accumulation (buckets with more entries
 * are cut into slices whose sums are folded afterwards).  0 restores the default: chosen per call from its size (16 for
 * small calls, up to 256 for large ones).  Range [4, 256].  Results are unchanged. */
int g16_ctx_set_item_max(g16_ctx *ctx, unsigned item_max);
int g16_device_count(void);
const char *g16_version(void);

/* ---- variable-base MSM --------------------------------------------------------------- */
int g16_g1_bases_upload(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, g16_bases **out);
int g16_g2_bases_upload(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, g16_bases **out);
/* wrap points that already live on the ctx's (single) device in the packed layout above with
 * (0,0) meaning infinity -- e.g. the output of g16_g1_fixed_base_mul_device.  Not owned. */
int g16_g1_bases_from_device(g16_ctx *ctx, const void *dev_xy, size_t n, g16_bases **out);
int g16_g2_bases_from_device(g16_ctx *ctx, const void *dev_xy, size_t n, g16_bases **out);
/* One-time preprocessing of resident bases (CRS arrays are fixed per ProvingKey): stores the multiples
 * 2^(c w) P_i for every window w, so that all windows of an MSM feed one shared bucket set -- no Horner
 * fold, one bucket reduction instead of ceil(256/c).  window_bits = 0 chooses c from the array length and
 * budget_bytes (0 = 48 GiB) of device memory per shard; *used_bits receives c.  Results are unchanged. */
int g16_bases_precompute(g16_ctx *ctx, g16_bases *bases, unsigned window_bits, size_t budget_bytes, unsigned *used_bits);
void g16_bases_free(g16_bases *bases);
size_t g16_bases_len(const g16_bases *bases);

/* sum_i scalars[i] * bases[i] for i < n (n <= len(bases)); host scalars, host result */
int g16_g1_msm(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, uint64_t out_xy[12],
               uint8_t *out_inf);
int g16_g2_msm(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, uint64_t out_xy[24],
               uint8_t *out_inf);
/* one call = what Prover::multi_scalar_mult_g1/_g2 does today: fresh bases + scalars from the host */
int g16_g1_msm_oneshot(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, const uint64_t *scalars, size_t n,
                       uint64_t out_xy[12], uint8_t *out_inf);
int g16_g2_msm_oneshot(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, const uint64_t *scalars, size_t n,
                       uint64_t out_xy[24], uint8_t *out_inf);

/* Device-resident variants (single-device ctx; asynchronous on the ctx stream).
 * dev_scalars: n x 4 u64 Montgomery on the device.
 * dev_out_affine: 12 (G1) / 24 (G2) u64 + one u32 infinity word, may be NULL.
 * dev_out_partial: projective partial sum, G16_G1_PARTIAL_WORDS / G16_G2_PARTIAL_WORDS u32, may be NULL.
 * Partials of index-range shards computed by different processes/GPUs are exchanged by the caller
 * (NCCL all-gather of raw bytes) and folded with g16_*_combine_partials_device. */
#define G16_G1_PARTIAL_WORDS 48
#define G16_G2_PARTIAL_WORDS 96
#define G16_G1_AFFINE_WORDS 25
#define G16_G2_AFFINE_WORDS 49
int g16_g1_msm_device(g16_ctx *ctx, const g16_bases *bases, const void *dev_scalars, size_t n, void *dev_out_affine,
                      void *dev_out_partial);
int g16_g2_msm_device(g16_ctx *ctx, const g16_bases *bases, const void *dev_scalars, size_t n, void *dev_out_affine,
                      void *dev_out_partial);
/* HOST scalars in, results on the device, asynchronous on the ctx stream (single-device ctx, unsharded
 * bases).  The H2D copy is pipelined against the computation in chunks for large n.  `scalars` must stay
 * valid until the stream has been synchronised (g16_ctx_synchronize or the caller's own stream sync). */
int g16_g1_msm_async(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, void *dev_out_affine,
                     void *dev_out_partial);
int g16_g2_msm_async(g16_ctx *ctx, const g16_bases *bases, const uint64_t *scalars, size_t n, void *dev_out_affine,
                     void *dev_out_partial);
int g16_g1_combine_partials_device(g16_ctx *ctx, const void *dev_partials, size_t k, void *dev_out_affine);
int g16_g2_combine_partials_device(g16_ctx *ctx, const void *dev_partials, size_t k, void *dev_out_affine);

/* ---- fixed-base scalar multiplication (CRS generation) --------------------------------- */
/* out[i] = (base * scalars[i]).into_affine();  out_xy: n x 12 (24) u64, out_inf: n bytes */
int g16_g1_fixed_base_mul(g16_ctx *ctx, const uint64_t base_xy[12], const uint64_t *scalars, size_t n,
                          uint64_t *out_xy, uint8_t *out_inf);
int g16_g2_fixed_base_mul(g16_ctx *ctx, const uint64_t base_xy[24], const uint64_t *scalars, size_t n,
                          uint64_t *out_xy, uint8_t *out_inf);
/* device in / device out (packed points, (0,0) = infinity); single-device ctx */
int g16_g1_fixed_base_mul_device(g16_ctx *ctx, const uint64_t base_xy[12], const void *dev_scalars, size_t n,
                                 void *dev_out_xy);
int g16_g2_fixed_base_mul_device(g16_ctx *ctx, const uint64_t base_xy[24], const void *dev_scalars, size_t n,
                                 void *dev_out_xy);

/* ---- Groth16 prove schedule -------------------------------------------------------------- */
/* Mirrors `ProvingKey` (crates/groth16-setup/src/lib.rs:27-52): single points as 12/24 u64 with an
 * infinity byte, vectors as n x 12/24 u64 with optional infinity byte arrays (NULL = none). */
typedef struct g16_pk_host {
    const uint64_t *alpha_g1, *beta_g1, *delta_g1; /* 12 u64 each */
    const uint64_t *beta_g2, *delta_g2;            /* 24 u64 each */
    const uint64_t *a_g1; const uint8_t *a_g1_inf; size_t a_len;
    const uint64_t *b_g1; const uint8_t *b_g1_inf; size_t b1_len;
    const uint64_t *b_g2; const uint8_t *b_g2_inf; size_t b2_len;
    const uint64_t *ic_g1; const uint8_t *ic_g1_inf; size_t ic_len;
    const uint64_t *h_g1; const uint8_t *h_g1_inf; size_t h_len;
    size_t num_public;
} g16_pk_host;
int g16_pk_upload(g16_ctx *ctx, const g16_pk_host *pk, g16_pk **out);
/* one-time g16_bases_precompute of the five resident arrays (skipped for arrays of < 256 points) */
int g16_pk_precompute(g16_ctx *ctx, g16_pk *pk);
/* The same with a promise about the scalars g16_prove will see: the assignment and the H coefficients are below
 * 2^scalar_bits (0 = full width).  The reference truncates both to 64 bits (crates/groth16-core/src/lib.rs:156-161,
 * 203-208), so a drop-in for it may pass 64: the tables are then built for a window size that suits 64-bit scalars
 * (fewer buckets to reduce).  Only a tuning hint: the tables cover all 256 bits, proofs for wider scalars (r, s and
 * any assignment) stay correct.  Replaces earlier tables of the key. */
int g16_pk_precompute_bits(g16_ctx *ctx, g16_pk *pk, unsigned scalar_bits);
void g16_pk_free(g16_pk *pk);
/* The group part of Prover::prove (crates/groth16-core/src/lib.rs:164-271).
 *   assignment_fr : num_vars x 4 u64, the already truncated `assignment_fr` of lib.rs:156-161
 *   h_coeffs      : h_len_used x 4 u64, `h_coeffs` of lib.rs:203-208 (may be NULL / 0)
 *   r, s          : 4 u64 each (lib.rs:152-153)
 * Outputs: proof.a (12 u64 + flag), proof.b (24 u64 + flag), proof.c (12 u64 + flag). */
int g16_prove(g16_ctx *ctx, const g16_pk *pk, const uint64_t *assignment_fr, size_t num_vars,
              const uint64_t *h_coeffs, size_t num_h, const uint64_t r[4], const uint64_t s[4],
              uint64_t a_xy[12], uint8_t *a_inf, uint64_t b_xy[24], uint8_t *b_inf, uint64_t c_xy[12], uint8_t *c_inf);

/* ---- quotient polynomial (the stage that feeds the H MSM) ---------------------------- */
/* H = (A*B - C) / Z with Z = x^n - 1, from the evaluations of A, B, C on ark-poly's radix-2 domain of size n
 * (a_evals[i] = <A-row i, assignment>; rows beyond the constraint count are zero).  Same polynomial as
 * `QAP::compute_quotient_polynomial` (crates/groth16-qap/src/lib.rs:225-271, called at
 * crates/groth16-core/src/lib.rs:200) for a satisfying assignment.  n must be a power of two; all arrays
 * are n x 4 u64 Montgomery Fr; h_coeffs receives n coefficients (the top one is zero).  The four buffers may be
 * pageable or pinned host memory (pinned: the copies run at PCIe rate and asynchronously to the host).  Returns
 * G16_ERR_INVALID ("Polynomial division failed") when A*B - C does not vanish on the domain. */
int g16_quotient_h(g16_ctx *ctx, const uint64_t *a_evals, const uint64_t *b_evals, const uint64_t *c_evals, size_t n,
                   uint64_t *h_coeffs);
/* Device in / device out, asynchronous on the ctx stream (single-device ctx): dev_abc = the evaluations of A, B, C
 * back to back (3 n x 4 u64, OVERWRITTEN), dev_h receives the n coefficients, *dev_bad_rows (one u32 on the device)
 * the number of domain points where A*B != C (0 = the division is exact). */
int g16_quotient_h_device(g16_ctx *ctx, void *dev_abc, size_t n, void *dev_h, void *dev_bad_rows);

/* ---- sparse R1CS: setup and prove for real circuits (rows 1-3 chained on the device) -------- */
/* The reference turns an R1CS into dense per-variable polynomials (`QAP::from_r1cs`,
 * crates/groth16-qap/src/lib.rs:95-187), Theta(constraints x variables) in memory.  These entry points keep the
 * three constraint matrices sparse and give the same values in O(non-zeros):
 *   g16_r1cs_domain_evals   <A-row i, w>, <B-row i, w>, <C-row i, w> for the n = next_power_of_two(constraints)
 *                           domain points = the evaluations of sum_k w_k A_k(x) etc. that
 *                           `compute_quotient_polynomial` (qap/src/lib.rs:225-271) sums densely
 *   g16_r1cs_eval_at        A_j(s), B_j(s), C_j(s) for every variable j = `qap.a_polys[j].evaluate(&s)`
 *                           (crates/groth16-setup/src/lib.rs:174-182), through the Lagrange basis at s
 *   g16_setup_crs           `CRS::generate_from_qap` (crates/groth16-setup/src/lib.rs:141-268): parameter checks,
 *                           64-bit truncations, IC / H exponents, and every `(gen * fr).into_affine()`
 *   g16_prove_r1cs          `Prover::prove` (crates/groth16-core/src/lib.rs:139-272) from the un-truncated witness:
 *                           Witness::validate, quotient polynomial, truncations, the five MSMs
 * A matrix is CSR over the constraints: row_ptr[num_constraints + 1], col[nnz] (variable indices; entries with
 * col >= num_variables are ignored like qap/src/lib.rs:121-138 does), val[nnz x 4 u64] (Fr, Montgomery).  When a
 * (row, variable) pair appears more than once in a matrix the last value wins, as in the reference's
 * `a_evals[row][var] = coeff` loop (its linear combinations are maps, so it never sees duplicates).
 * Single-device contexts only. */
typedef struct g16_r1cs g16_r1cs;
typedef struct g16_csr {
    const uint32_t *row_ptr;
    const uint32_t *col;
    const uint64_t *val;
} g16_csr;
int g16_r1cs_upload(g16_ctx *ctx, size_t num_constraints, size_t num_variables, const g16_csr *a, const g16_csr *b,
                    const g16_csr *c, g16_r1cs **out);
void g16_r1cs_free(g16_r1cs *r1cs);
size_t g16_r1cs_domain_size(const g16_r1cs *r1cs);
/* assignment: num_vars x 4 u64 (must equal num_variables, else G16_ERR_LENGTH); *_evals: domain_size x 4 u64 */
int g16_r1cs_domain_evals(g16_ctx *ctx, const g16_r1cs *r1cs, const uint64_t *assignment, size_t num_vars,
                          uint64_t *a_evals, uint64_t *b_evals, uint64_t *c_evals);
/* s is used as given (no truncation); *_vals: num_variables x 4 u64 */
int g16_r1cs_eval_at(g16_ctx *ctx, const g16_r1cs *r1cs, const uint64_t s[4], uint64_t *a_vals, uint64_t *b_vals,
                     uint64_t *c_vals);
/* Outputs of g16_setup_crs on the host, ark layout; every pointer may be NULL (not wanted).  Lengths:
 * a_g1, b_g1, b_g2: num_variables; ic_g1: num_variables - num_public - 1 (ProvingKey.ic_g1); vk_ic_g1:
 * num_public + 1 (VerificationKey.ic_g1); h_g1: domain size (= qap.degree()). */
typedef struct g16_crs_host {
    uint64_t *alpha_g1, *beta_g1, *delta_g1;   /* 12 u64 each */
    uint64_t *beta_g2, *gamma_g2, *delta_g2;   /* 24 u64 each */
    uint64_t *a_g1; uint8_t *a_g1_inf;
    uint64_t *b_g1; uint8_t *b_g1_inf;
    uint64_t *b_g2; uint8_t *b_g2_inf;
    uint64_t *ic_g1; uint8_t *ic_g1_inf;
    uint64_t *vk_ic_g1; uint8_t *vk_ic_g1_inf;
    uint64_t *h_g1; uint8_t *h_g1_inf;
} g16_crs_host;
/* alpha .. s: the five SetupParams (4 u64 Montgomery each, full width; the reference truncates them itself where
 * it does).  out may be NULL; pk_out (may be NULL) receives the device-resident proving key, ready for g16_prove /
 * g16_prove_r1cs without a host round trip.  Errors mirror SetupError::InvalidParams. */
int g16_setup_crs(g16_ctx *ctx, const g16_r1cs *r1cs, const uint64_t alpha[4], const uint64_t beta[4],
                  const uint64_t gamma[4], const uint64_t delta[4], const uint64_t s[4], size_t num_public,
                  g16_crs_host *out, g16_pk **pk_out);
/* assignment: the witness as the reference holds it (F elements, NOT truncated).  Errors: G16_ERR_LENGTH
 * (assignment length), G16_ERR_INVALID "Invalid witness: ..." (Witness::validate) or "Polynomial division
 * failed" (QAPError::PolynomialDivisionFailed). */
int g16_prove_r1cs(g16_ctx *ctx, const g16_pk *pk, const g16_r1cs *r1cs, const uint64_t *assignment, size_t num_vars,
                   const uint64_t r[4], const uint64_t s[4], uint64_t a_xy[12], uint8_t *a_inf, uint64_t b_xy[24],
                   uint8_t *b_inf, uint64_t c_xy[12], uint8_t *c_inf);

/* ---- wire format ------------------------------------------------------------------- */
/* ark `CanonicalSerialize` / `CanonicalDeserialize` as ark-bls12-381 0.4.0 implements them for G1Affine and
 * G2Affine (the Zcash encoding): Fq = 48 bytes big-endian canonical; G1 = x (compressed, 48 B) or x || y (96 B);
 * G2 = x.c1 || x.c0 (96 B) or x.c1 || x.c0 || y.c1 || y.c0 (192 B); top bits of byte 0: 0x80 compressed, 0x40
 * infinity, 0x20 "y is the larger root".  `Proof` (crates/groth16-core/src/lib.rs:27-36, derive
 * CanonicalSerialize) = a || b || c = 192 bytes compressed / 384 uncompressed -- what `serialize_compressed` /
 * `deserialize_compressed` of the reference's proof produce and accept.  The vector entry points are the batch
 * form a key file needs (`Vec<G1Affine>` body; ark's u64 little-endian length prefix is the caller's).
 * Deserialisation errors mirror ark's SerializationError: G16_ERR_INVALID with "UnexpectedFlags" (compression
 * bit does not match) or "InvalidData" (coordinate >= q, no point with that x, or -- validate != 0, ark's
 * Validate::Yes -- not in the prime-order subgroup); status (n bytes, may be NULL) receives 0 / 1 (InvalidData) /
 * 2 (UnexpectedFlags) per element, rejected elements come back as the identity.  Uncompressed input is also
 * checked against the curve equation when validate != 0.  Single-device contexts. */
int g16_g1_serialize(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out);
int g16_g2_serialize(g16_ctx *ctx, const uint64_t *xy, const uint8_t *inf, size_t n, int compressed, uint8_t *out);
int g16_g1_deserialize(g16_ctx *ctx, const uint8_t *bytes, size_t n, int compressed, int validate, uint64_t *out_xy,
                       uint8_t *out_inf, uint8_t *status);
int g16_g2_deserialize(g16_ctx *ctx, const uint8_t *bytes, size_t n, int compressed, int validate, uint64_t *out_xy,
                       uint8_t *out_inf, uint8_t *status);
int g16_proof_serialize(g16_ctx *ctx, const uint64_t a_xy[12], uint8_t a_inf, const uint64_t b_xy[24], uint8_t b_inf,
                        const uint64_t c_xy[12], uint8_t c_inf, int compressed, uint8_t *out);
int g16_proof_deserialize(g16_ctx *ctx, const uint8_t *bytes, int compressed, int validate, uint64_t a_xy[12], uint8_t *a_inf,
                          uint64_t b_xy[24], uint8_t *b_inf, uint64_t c_xy[12], uint8_t *c_inf);

/* ---- test hooks (used by tests/ and bench.py only) ------------------------------------------ */
/* kernels launched by the library since it was loaded */
unsigned long long g16_launch_count(void);
/* CUDA-event timing of the stages of the last MSM on a single-device ctx, in ms:
 * [digit count, offset scan, scatter, bucket accumulate, bucket reduce, window combine];
 * plan = {window bits, windows, buckets per window} */
int g16_ctx_enable_stage_timing(g16_ctx *ctx, int on);
int g16_ctx_last_stage_ms(g16_ctx *ctx, float ms[6], unsigned plan[3]);
/* with stage timing enabled: the timeline of the last g16_prove on a single-device ctx.  t[7 * lane + k] = ms from the
 * start of the prove to mark k (0 start, 1 digits done, 2 items, 3 sort, 4 accumulate, 5 reduce, 6 fold) of lane
 * 0 pi_A, 1 pi_B (G2), 2 H, 3 pi_B' (G1), 4 private part of pi_C */
int g16_ctx_prove_timeline(g16_ctx *ctx, float t[35]);
/* element-wise Fq ops on the device: op 0 mul, 1 add, 2 sub, 3 inverse(a), 4 square(a), 5 negate(a);
 * a, b, out: n x 6 u64 Montgomery (b may be NULL for unary ops) */
int g16_debug_fq_op(g16_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* Fr Montgomery -> canonical (`into_bigint()`), n x 4 u64 */
int g16_debug_fr_from_mont(g16_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
/* out[i] = affine(P[i] + Q[i]) through the XYZZ mixed addition, all exceptional cases included */
int g16_debug_g1_add(g16_ctx *ctx, const uint64_t *p, const uint8_t *p_inf, const uint64_t *q, const uint8_t *q_inf,
                     uint64_t *out_xy, uint8_t *out_inf, size_t n);
int g16_debug_g2_add(g16_ctx *ctx, const uint64_t *p, const uint8_t *p_inf, const uint64_t *q, const uint8_t *q_inf,
                     uint64_t *out_xy, uint8_t *out_inf, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* G16_CUDA_H */
