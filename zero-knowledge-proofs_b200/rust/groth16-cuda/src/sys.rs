//! Raw FFI declarations: one `extern "C"` item per function of `include/g16_cuda.h`, same order, same
//! argument meaning.  `tests/test_abi_exports.py` checks mechanically that this list and the header agree.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_uint, c_ulonglong, c_void};

#[repr(C)] pub struct g16_ctx { _p: [u8; 0] }
#[repr(C)] pub struct g16_bases { _p: [u8; 0] }
#[repr(C)] pub struct g16_pk { _p: [u8; 0] }
#[repr(C)] pub struct g16_r1cs { _p: [u8; 0] }

pub const G16_OK: c_int = 0;
pub const G16_ERR_INVALID: c_int = 1;
pub const G16_ERR_CUDA: c_int = 2;
pub const G16_ERR_NO_DEVICE: c_int = 3;
pub const G16_ERR_OOM: c_int = 4;
pub const G16_ERR_LENGTH: c_int = 5;
pub const G16_G1_PARTIAL_WORDS: usize = 48;
pub const G16_G2_PARTIAL_WORDS: usize = 96;
pub const G16_G1_AFFINE_WORDS: usize = 25;
pub const G16_G2_AFFINE_WORDS: usize = 49;

/// `struct g16_pk_host` (mirrors `ProvingKey`, crates/groth16-setup/src/lib.rs:27-52)
#[repr(C)]
pub struct g16_pk_host {
    pub alpha_g1: *const u64, pub beta_g1: *const u64, pub delta_g1: *const u64,
    pub beta_g2: *const u64, pub delta_g2: *const u64,
    pub a_g1: *const u64, pub a_g1_inf: *const u8, pub a_len: usize,
    pub b_g1: *const u64, pub b_g1_inf: *const u8, pub b1_len: usize,
    pub b_g2: *const u64, pub b_g2_inf: *const u8, pub b2_len: usize,
    pub ic_g1: *const u64, pub ic_g1_inf: *const u8, pub ic_len: usize,
    pub h_g1: *const u64, pub h_g1_inf: *const u8, pub h_len: usize,
    pub num_public: usize,
}
#[repr(C)]
pub struct g16_csr { pub row_ptr: *const u32, pub col: *const u32, pub val: *const u64 }
#[repr(C)]
pub struct g16_crs_host {
    pub alpha_g1: *mut u64, pub beta_g1: *mut u64, pub delta_g1: *mut u64,
    pub beta_g2: *mut u64, pub gamma_g2: *mut u64, pub delta_g2: *mut u64,
    pub a_g1: *mut u64, pub a_g1_inf: *mut u8,
    pub b_g1: *mut u64, pub b_g1_inf: *mut u8,
    pub b_g2: *mut u64, pub b_g2_inf: *mut u8,
    pub ic_g1: *mut u64, pub ic_g1_inf: *mut u8,
    pub vk_ic_g1: *mut u64, pub vk_ic_g1_inf: *mut u8,
    pub h_g1: *mut u64, pub h_g1_inf: *mut u8,
}

extern "C" {
    // ---- context
    pub fn g16_ctx_create(devices: *const c_int, ndev: c_int, out: *mut *mut g16_ctx) -> c_int;
    pub fn g16_ctx_destroy(ctx: *mut g16_ctx);
    pub fn g16_last_error(ctx: *const g16_ctx) -> *const c_char;
    pub fn g16_ctx_set_stream(ctx: *mut g16_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn g16_ctx_synchronize(ctx: *mut g16_ctx) -> c_int;
    pub fn g16_ctx_set_window_bits(ctx: *mut g16_ctx, c: c_uint) -> c_int;
    pub fn g16_ctx_set_h2d_pipeline_min(ctx: *mut g16_ctx, min_scalars: usize) -> c_int;
    pub fn g16_ctx_set_item_max(ctx: *mut g16_ctx, item_max: c_uint) -> c_int;
    pub fn g16_device_count() -> c_int;
    pub fn g16_version() -> *const c_char;
    // ---- variable-base MSM
    pub fn g16_g1_bases_upload(ctx: *mut g16_ctx, xy: *const u64, inf: *const u8, n: usize, out: *mut *mut g16_bases) -> c_int;
    pub fn g16_g2_bases_upload(ctx: *mut g16_ctx, xy: *const u64, inf: *const u8, n: usize, out: *mut *mut g16_bases) -> c_int;
    pub fn g16_g1_bases_from_device(ctx: *mut g16_ctx, dev_xy: *const c_void, n: usize, out: *mut *mut g16_bases) -> c_int;
    pub fn g16_g2_bases_from_device(ctx: *mut g16_ctx, dev_xy: *const c_void, n: usize, out: *mut *mut g16_bases) -> c_int;
    pub fn g16_bases_precompute(ctx: *mut g16_ctx, bases: *mut g16_bases, window_bits: c_uint, budget_bytes: usize, used_bits: *mut c_uint) -> c_int;
    pub fn g16_bases_free(bases: *mut g16_bases);
    pub fn g16_bases_len(bases: *const g16_bases) -> usize;
    pub fn g16_g1_msm(ctx: *mut g16_ctx, bases: *const g16_bases, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    pub fn g16_g2_msm(ctx: *mut g16_ctx, bases: *const g16_bases, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    pub fn g16_g1_msm_oneshot(ctx: *mut g16_ctx, xy: *const u64, inf: *const u8, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    pub fn g16_g2_msm_oneshot(ctx: *mut g16_ctx, xy: *const u64, inf: *const u8, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    pub fn g16_g1_msm_device(ctx: *mut g16_ctx, bases: *const g16_bases, dev_scalars: *const c_void, n: usize, dev_out_affine: *mut c_void, dev_out_partial: *mut c_void) -> c_int;
    pub fn g16_g2_msm_device(ctx: *mut g16_ctx, bases: *const g16_bases, dev_scalars: *const c_void, n: usize, dev_out_affine: *mut c_void, dev_out_partial: *mut c_void) -> c_int;
    pub fn g16_g1_msm_async(ctx: *mut g16_ctx, bases: *const g16_bases, scalars: *const u64, n: usize, dev_out_affine: *mut c_void, dev_out_partial: *mut c_void) -> c_int;
    pub fn g16_g2_msm_async(ctx: *mut g16_ctx, bases: *const g16_bases, scalars: *const u64, n: usize, dev_out_affine: *mut c_void, dev_out_partial: *mut c_void) -> c_int;
    pub fn g16_g1_combine_partials_device(ctx: *mut g16_ctx, dev_partials: *const c_void, k: usize, dev_out_affine: *mut c_void) -> c_int;
    pub fn g16_g2_combine_partials_device(ctx: *mut g16_ctx, dev_partials: *const c_void, k: usize, dev_out_affine: *mut c_void) -> c_int;
    // ---- fixed base
    pub fn g16_g1_fixed_base_mul(ctx: *mut g16_ctx, base_xy: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    pub fn g16_g2_fixed_base_mul(ctx: *mut g16_ctx, base_xy: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    pub fn g16_g1_fixed_base_mul_device(ctx: *mut g16_ctx, base_xy: *const u64, dev_scalars: *const c_void, n: usize, dev_out_xy: *mut c_void) -> c_int;
    pub fn g16_g2_fixed_base_mul_device(ctx: *mut g16_ctx, base_xy: *const u64, dev_scalars: *const c_void, n: usize, dev_out_xy: *mut c_void) -> c_int;
    // ---- prove schedule
    pub fn g16_pk_upload(ctx: *mut g16_ctx, pk: *const g16_pk_host, out: *mut *mut g16_pk) -> c_int;
    pub fn g16_pk_precompute(ctx: *mut g16_ctx, pk: *mut g16_pk) -> c_int;
    pub fn g16_pk_precompute_bits(ctx: *mut g16_ctx, pk: *mut g16_pk, scalar_bits: c_uint) -> c_int;
    pub fn g16_pk_free(pk: *mut g16_pk);
    pub fn g16_prove(ctx: *mut g16_ctx, pk: *const g16_pk, assignment_fr: *const u64, num_vars: usize, h_coeffs: *const u64, num_h: usize,
                     r: *const u64, s: *const u64, a_xy: *mut u64, a_inf: *mut u8, b_xy: *mut u64, b_inf: *mut u8, c_xy: *mut u64, c_inf: *mut u8) -> c_int;
    // ---- quotient polynomial
    pub fn g16_quotient_h(ctx: *mut g16_ctx, a_evals: *const u64, b_evals: *const u64, c_evals: *const u64, n: usize, h_coeffs: *mut u64) -> c_int;
    pub fn g16_quotient_h_device(ctx: *mut g16_ctx, dev_abc: *mut c_void, n: usize, dev_h: *mut c_void, dev_bad_rows: *mut c_void) -> c_int;
    // ---- sparse R1CS
    pub fn g16_r1cs_upload(ctx: *mut g16_ctx, num_constraints: usize, num_variables: usize, a: *const g16_csr, b: *const g16_csr, c: *const g16_csr, out: *mut *mut g16_r1cs) -> c_int;
    pub fn g16_r1cs_free(r1cs: *mut g16_r1cs);
    pub fn g16_r1cs_domain_size(r1cs: *const g16_r1cs) -> usize;
    pub fn g16_r1cs_domain_evals(ctx: *mut g16_ctx, r1cs: *const g16_r1cs, assignment: *const u64, num_vars: usize, a_evals: *mut u64, b_evals: *mut u64, c_evals: *mut u64) -> c_int;
    pub fn g16_r1cs_eval_at(ctx: *mut g16_ctx, r1cs: *const g16_r1cs, s: *const u64, a_vals: *mut u64, b_vals: *mut u64, c_vals: *mut u64) -> c_int;
    pub fn g16_setup_crs(ctx: *mut g16_ctx, r1cs: *const g16_r1cs, alpha: *const u64, beta: *const u64, gamma: *const u64, delta: *const u64, s: *const u64,
                         num_public: usize, out: *mut g16_crs_host, pk_out: *mut *mut g16_pk) -> c_int;
    pub fn g16_prove_r1cs(ctx: *mut g16_ctx, pk: *const g16_pk, r1cs: *const g16_r1cs, assignment: *const u64, num_vars: usize, r: *const u64, s: *const u64,
                          a_xy: *mut u64, a_inf: *mut u8, b_xy: *mut u64, b_inf: *mut u8, c_xy: *mut u64, c_inf: *mut u8) -> c_int;
    // ---- wire format
    pub fn g16_g1_serialize(ctx: *mut g16_ctx, xy: *const u64, inf: *const u8, n: usize, compressed: c_int, out: *mut u8) -> c_int;
    pub fn g16_g2_serialize(ctx: *mut g16_ctx, xy: *const u64, inf: *const u8, n: usize, compressed: c_int, out: *mut u8) -> c_int;
    pub fn g16_g1_deserialize(ctx: *mut g16_ctx, bytes: *const u8, n: usize, compressed: c_int, validate: c_int, out_xy: *mut u64, out_inf: *mut u8, status: *mut u8) -> c_int;
    pub fn g16_g2_deserialize(ctx: *mut g16_ctx, bytes: *const u8, n: usize, compressed: c_int, validate: c_int, out_xy: *mut u64, out_inf: *mut u8, status: *mut u8) -> c_int;
    pub fn g16_proof_serialize(ctx: *mut g16_ctx, a_xy: *const u64, a_inf: u8, b_xy: *const u64, b_inf: u8, c_xy: *const u64, c_inf: u8, compressed: c_int, out: *mut u8) -> c_int;
    pub fn g16_proof_deserialize(ctx: *mut g16_ctx, bytes: *const u8, compressed: c_int, validate: c_int, a_xy: *mut u64, a_inf: *mut u8, b_xy: *mut u64, b_inf: *mut u8,
                                 c_xy: *mut u64, c_inf: *mut u8) -> c_int;
    // ---- test hooks
    pub fn g16_launch_count() -> c_ulonglong;
    pub fn g16_ctx_enable_stage_timing(ctx: *mut g16_ctx, on: c_int) -> c_int;
    pub fn g16_ctx_last_stage_ms(ctx: *mut g16_ctx, ms: *mut f32, plan: *mut c_uint) -> c_int;
    pub fn g16_ctx_prove_timeline(ctx: *mut g16_ctx, t: *mut f32) -> c_int;
    pub fn g16_debug_fq_op(ctx: *mut g16_ctx, op: c_int, a: *const u64, b: *const u64, out: *mut u64, n: usize) -> c_int;
    pub fn g16_debug_fr_from_mont(ctx: *mut g16_ctx, a: *const u64, out: *mut u64, n: usize) -> c_int;
    pub fn g16_debug_g1_add(ctx: *mut g16_ctx, p: *const u64, p_inf: *const u8, q: *const u64, q_inf: *const u8, out_xy: *mut u64, out_inf: *mut u8, n: usize) -> c_int;
    pub fn g16_debug_g2_add(ctx: *mut g16_ctx, p: *const u64, p_inf: *const u8, q: *const u64, q_inf: *const u8, out_xy: *mut u64, out_inf: *mut u8, n: usize) -> c_int;
}
