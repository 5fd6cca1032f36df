//! `groth16-cuda`: safe Rust wrapper of the B200 MSM engine (C ABI in `include/g16_cuda.h`, raw items in `sys`).
//!
//! Drop-in seam (reference = vats98754/zero-knowledge-proofs):
//! * `msm_g1` / `msm_g2` replace `G1Projective::msm(&points, &scalars)` / `G2Projective::msm` at
//!   `crates/groth16-core/src/lib.rs:282,296` (inside `Prover::multi_scalar_mult_g1/_g2`, `:275-300`) -- the
//!   two-line edit of INTEGRATION.md 1, bases uploaded per call;
//! * `DeviceBases` keeps one CRS array resident (optionally with the precomputed table of multiples), and
//!   `DeviceProvingKey::upload` + `DeviceProvingKey::prove` keep the whole `ProvingKey` resident and run the
//!   4 x G1 + 1 x G2 MSM schedule of `Prover::prove` (`lib.rs:164-271`) in one call -- the measured fast path,
//!   INTEGRATION.md 2;
//! * `fixed_base_mul_g1/_g2` replace the `par_iter().map(|v| (gen * fr).into_affine())` blocks of
//!   `CRS::generate_from_qap` at `crates/groth16-setup/src/lib.rs:185-241`;
//! * `DeviceR1cs` (+ `setup_crs`, `prove_r1cs`, `quotient_h`) are the sparse path for circuits the dense
//!   `QAP::from_r1cs` cannot hold;
//! * `serialize_*` / `deserialize_*` / `proof_to_bytes` / `proof_from_bytes` are ark's `CanonicalSerialize` bytes.
//!
//! Marshalling copies ark's raw limbs (`p.x.0.0`, already Montgomery) -- no field arithmetic on the host.
//! Errors become `String`s that the callers map to `GrothError::MSMError(..)` exactly like `lib.rs:283`.
//! There is no CPU fallback: every call needs libg16cuda.so and a CUDA device.

pub mod sys;

use ark_bls12_381::{Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ff::{BigInt, Fp};
use std::marker::PhantomData;
use std::os::raw::{c_int, c_uint, c_void};
use sys::*;

/// One engine context (one or more GPUs of one box).  `!Sync`: one caller at a time, like the
/// single-threaded reference prover; wrap in a `Mutex` to share.
pub struct Context { raw: *mut g16_ctx, _not_sync: PhantomData<*mut ()> }
unsafe impl Send for Context {}

fn err(ctx: *const g16_ctx, code: c_int) -> String {
    let msg = unsafe { std::ffi::CStr::from_ptr(g16_last_error(ctx)) }.to_string_lossy().into_owned();
    format!("{} (code {})", msg, code)
}
fn check(ctx: *const g16_ctx, rc: c_int) -> Result<(), String> { if rc == G16_OK { Ok(()) } else { Err(err(ctx, rc)) } }

impl Context {
    /// `devices = &[]` uses the current CUDA device; several devices shard every bases array by index range.
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { g16_ctx_create(if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() }, devices.len() as c_int, &mut raw) };
        if rc != G16_OK { return Err(err(std::ptr::null(), rc)); }
        Ok(Context { raw, _not_sync: PhantomData })
    }
    pub fn device_count() -> i32 { unsafe { g16_device_count() } }
    /// e.g. "groth16-cuda 0.2 (sm_100a) src:<digest of the sources the library was built from>"
    pub fn version() -> String { unsafe { std::ffi::CStr::from_ptr(g16_version()) }.to_string_lossy().into_owned() }
    /// Run single-device work on the caller's `cudaStream_t`.
    /// # Safety
    /// `cuda_stream` must be a live stream of the context's device until the context is dropped or re-pointed.
    pub unsafe fn set_stream(&self, cuda_stream: *mut c_void) -> Result<(), String> { check(self.raw, g16_ctx_set_stream(self.raw, cuda_stream)) }
    pub fn synchronize(&self) -> Result<(), String> { check(self.raw, unsafe { g16_ctx_synchronize(self.raw) }) }
    /// Window bits of the next MSMs over plain (not precomputed) bases; 0 = choose from the length.
    pub fn set_window_bits(&self, c: u32) -> Result<(), String> { check(self.raw, unsafe { g16_ctx_set_window_bits(self.raw, c as c_uint) }) }
    /// Host-scalar MSMs of at least this many scalars per device pipeline their H2D copy; 0 = default (2^19).
    pub fn set_h2d_pipeline_min(&self, min_scalars: usize) -> Result<(), String> { check(self.raw, unsafe { g16_ctx_set_h2d_pipeline_min(self.raw, min_scalars) }) }
    /// Tuning: longest serial run of additions per thread in the bucket accumulation (0 = chosen per call).
    pub fn set_item_max(&self, item_max: u32) -> Result<(), String> { check(self.raw, unsafe { g16_ctx_set_item_max(self.raw, item_max as c_uint) }) }
    /// Kernels launched by the library since it was loaded (bench / tests).
    pub fn launch_count() -> u64 { unsafe { g16_launch_count() as u64 } }
    /// CUDA-event timing of the stages of the next MSMs on a single-device context (bench / tests).
    pub fn enable_stage_timing(&self, on: bool) -> Result<(), String> { check(self.raw, unsafe { g16_ctx_enable_stage_timing(self.raw, on as c_int) }) }
    /// ([digits, scan, scatter, accumulate, reduce, combine] in ms, [window bits, windows, buckets per window]).
    pub fn last_stage_ms(&self) -> Result<([f32; 6], [u32; 3]), String> {
        let (mut ms, mut plan) = ([0f32; 6], [0 as c_uint; 3]);
        check(self.raw, unsafe { g16_ctx_last_stage_ms(self.raw, ms.as_mut_ptr(), plan.as_mut_ptr()) })?;
        Ok((ms, [plan[0] as u32, plan[1] as u32, plan[2] as u32]))
    }
    /// With stage timing enabled: ms from the start of the last `prove` to the seven stage marks of its five lanes.
    pub fn prove_timeline(&self) -> Result<[[f32; 7]; 5], String> {
        let mut t = [0f32; 35];
        check(self.raw, unsafe { g16_ctx_prove_timeline(self.raw, t.as_mut_ptr()) })?;
        let mut out = [[0f32; 7]; 5];
        for lane in 0..5 { out[lane].copy_from_slice(&t[7 * lane..7 * lane + 7]); }
        Ok(out)
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { g16_ctx_destroy(self.raw) } } }

// ---- marshalling: ark in-memory limbs <-> packed u64 buffers ------------------------------------
fn push_fq(out: &mut Vec<u64>, f: &Fq) { out.extend_from_slice(&(f.0).0); }
fn push_fq2(out: &mut Vec<u64>, f: &Fq2) { push_fq(out, &f.c0); push_fq(out, &f.c1); }
fn fq_from(l: &[u64]) -> Fq { let mut a = [0u64; 6]; a.copy_from_slice(l); Fp(BigInt(a), PhantomData) }
fn fr_from(l: &[u64]) -> Fr { let mut a = [0u64; 4]; a.copy_from_slice(l); Fp(BigInt(a), PhantomData) }
fn pack_scalars(s: &[Fr]) -> Vec<u64> { let mut v = Vec::with_capacity(4 * s.len()); for x in s { v.extend_from_slice(&(x.0).0); } v }
fn unpack_scalars(v: &[u64]) -> Vec<Fr> { v.chunks_exact(4).map(fr_from).collect() }
fn pack_g1(p: &[G1Affine]) -> (Vec<u64>, Vec<u8>) {
    let (mut xy, mut inf) = (Vec::with_capacity(12 * p.len()), Vec::with_capacity(p.len()));
    for q in p { push_fq(&mut xy, &q.x); push_fq(&mut xy, &q.y); inf.push(q.infinity as u8); }
    (xy, inf)
}
fn pack_g2(p: &[G2Affine]) -> (Vec<u64>, Vec<u8>) {
    let (mut xy, mut inf) = (Vec::with_capacity(24 * p.len()), Vec::with_capacity(p.len()));
    for q in p { push_fq2(&mut xy, &q.x); push_fq2(&mut xy, &q.y); inf.push(q.infinity as u8); }
    (xy, inf)
}
/// a single point for the `g16_pk_host` fields that carry no flag: the identity travels as all-zero coordinates
fn pack_single_g1(p: &G1Affine) -> Vec<u64> { if p.infinity { vec![0u64; 12] } else { pack_g1(std::slice::from_ref(p)).0 } }
fn pack_single_g2(p: &G2Affine) -> Vec<u64> { if p.infinity { vec![0u64; 24] } else { pack_g2(std::slice::from_ref(p)).0 } }
fn g1_from(xy: &[u64], inf: u8) -> G1Affine {
    if inf != 0 { return G1Affine::identity(); }
    G1Affine::new_unchecked(fq_from(&xy[0..6]), fq_from(&xy[6..12]))
}
fn g2_from(xy: &[u64], inf: u8) -> G2Affine {
    if inf != 0 { return G2Affine::identity(); }
    G2Affine::new_unchecked(Fq2::new(fq_from(&xy[0..6]), fq_from(&xy[6..12])), Fq2::new(fq_from(&xy[12..18]), fq_from(&xy[18..24])))
}
fn g1_vec(xy: &[u64], inf: &[u8]) -> Vec<G1Affine> { (0..inf.len()).map(|i| g1_from(&xy[12 * i..12 * i + 12], inf[i])).collect() }
fn g2_vec(xy: &[u64], inf: &[u8]) -> Vec<G2Affine> { (0..inf.len()).map(|i| g2_from(&xy[24 * i..24 * i + 24], inf[i])).collect() }

// ---- the seam of Prover::multi_scalar_mult_g1/_g2 ------------------------------------------------
/// Σ scalars[i]·bases[i] as an affine point.  Length mismatch -> `Err` (ark returns `Err(min_len)`).
pub fn msm_g1(ctx: &Context, bases: &[G1Affine], scalars: &[Fr]) -> Result<G1Affine, String> {
    if bases.len() != scalars.len() { return Err(format!("{}", bases.len().min(scalars.len()))); }
    let (xy, inf) = pack_g1(bases);
    let sc = pack_scalars(scalars);
    let (mut out, mut oinf) = ([0u64; 12], 0u8);
    check(ctx.raw, unsafe { g16_g1_msm_oneshot(ctx.raw, xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), bases.len(), out.as_mut_ptr(), &mut oinf) })?;
    Ok(g1_from(&out, oinf))
}
pub fn msm_g2(ctx: &Context, bases: &[G2Affine], scalars: &[Fr]) -> Result<G2Affine, String> {
    if bases.len() != scalars.len() { return Err(format!("{}", bases.len().min(scalars.len()))); }
    let (xy, inf) = pack_g2(bases);
    let sc = pack_scalars(scalars);
    let (mut out, mut oinf) = ([0u64; 24], 0u8);
    check(ctx.raw, unsafe { g16_g2_msm_oneshot(ctx.raw, xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), bases.len(), out.as_mut_ptr(), &mut oinf) })?;
    Ok(g2_from(&out, oinf))
}

// ---- resident CRS arrays ----------------------------------------------------------------------------
/// Base points that stay on the device(s) between proofs.  Tied to its context by the lifetime.
pub struct DeviceBases<'a> { ctx: &'a Context, raw: *mut g16_bases, g2: bool }
impl<'a> DeviceBases<'a> {
    pub fn upload_g1(ctx: &'a Context, p: &[G1Affine]) -> Result<Self, String> {
        let (xy, inf) = pack_g1(p);
        let mut raw = std::ptr::null_mut();
        check(ctx.raw, unsafe { g16_g1_bases_upload(ctx.raw, xy.as_ptr(), inf.as_ptr(), p.len(), &mut raw) })?;
        Ok(DeviceBases { ctx, raw, g2: false })
    }
    pub fn upload_g2(ctx: &'a Context, p: &[G2Affine]) -> Result<Self, String> {
        let (xy, inf) = pack_g2(p);
        let mut raw = std::ptr::null_mut();
        check(ctx.raw, unsafe { g16_g2_bases_upload(ctx.raw, xy.as_ptr(), inf.as_ptr(), p.len(), &mut raw) })?;
        Ok(DeviceBases { ctx, raw, g2: true })
    }
    /// Wrap packed points that already live on the (single) device, e.g. the output of `fixed_base_mul_*_device`.
    /// # Safety
    /// `dev_xy` must point to `n` packed points ((0,0) = identity) on the context's device and outlive `self`.
    pub unsafe fn from_device(ctx: &'a Context, g2: bool, dev_xy: *const c_void, n: usize) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        let rc = if g2 { g16_g2_bases_from_device(ctx.raw, dev_xy, n, &mut raw) } else { g16_g1_bases_from_device(ctx.raw, dev_xy, n, &mut raw) };
        check(ctx.raw, rc)?;
        Ok(DeviceBases { ctx, raw, g2 })
    }
    pub fn len(&self) -> usize { unsafe { g16_bases_len(self.raw) } }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    /// One-time table of multiples 2^(c w) P_i: all windows of later MSMs share one bucket set.  `window_bits = 0`
    /// chooses c from the length and `budget_bytes` (0 = 48 GiB per shard).  Returns c (0 = nothing fitted).
    pub fn precompute(&mut self, window_bits: u32, budget_bytes: usize) -> Result<u32, String> {
        let mut used: c_uint = 0;
        check(self.ctx.raw, unsafe { g16_bases_precompute(self.ctx.raw, self.raw, window_bits as c_uint, budget_bytes, &mut used) })?;
        Ok(used as u32)
    }
    /// Σ scalars[i]·bases[i] over the first `scalars.len()` resident bases (zero scalars cost nothing).
    pub fn msm_g1(&self, scalars: &[Fr]) -> Result<G1Affine, String> {
        if self.g2 { return Err("G2 bases passed to msm_g1".into()); }
        let sc = pack_scalars(scalars);
        let (mut out, mut oinf) = ([0u64; 12], 0u8);
        check(self.ctx.raw, unsafe { g16_g1_msm(self.ctx.raw, self.raw, sc.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut oinf) })?;
        Ok(g1_from(&out, oinf))
    }
    pub fn msm_g2(&self, scalars: &[Fr]) -> Result<G2Affine, String> {
        if !self.g2 { return Err("G1 bases passed to msm_g2".into()); }
        let sc = pack_scalars(scalars);
        let (mut out, mut oinf) = ([0u64; 24], 0u8);
        check(self.ctx.raw, unsafe { g16_g2_msm(self.ctx.raw, self.raw, sc.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut oinf) })?;
        Ok(g2_from(&out, oinf))
    }
    /// Device scalars in, device results out (affine: 12 / 24 u64 + one u32 flag word; partial: 48 / 96 u32), asynchronous
    /// on the context stream.  Index-range shards of several processes exchange the partials themselves
    /// (NCCL all-gather of raw bytes) and fold them with `combine_partials_device`.
    /// # Safety
    /// All pointers are device pointers of the context's device with the sizes above; null = not wanted.
    pub unsafe fn msm_device(&self, dev_scalars: *const c_void, n: usize, dev_out_affine: *mut c_void, dev_out_partial: *mut c_void) -> Result<(), String> {
        let rc = if self.g2 { g16_g2_msm_device(self.ctx.raw, self.raw, dev_scalars, n, dev_out_affine, dev_out_partial) }
                 else { g16_g1_msm_device(self.ctx.raw, self.raw, dev_scalars, n, dev_out_affine, dev_out_partial) };
        check(self.ctx.raw, rc)
    }
    /// HOST scalars (pinned memory for a truly asynchronous copy) in, device results out, asynchronous on the context
    /// stream; the H2D copy is pipelined against the computation.
    /// # Safety
    /// `scalars` must stay valid until the stream has been synchronised; outputs as in `msm_device`.
    pub unsafe fn msm_async(&self, scalars: *const u64, n: usize, dev_out_affine: *mut c_void, dev_out_partial: *mut c_void) -> Result<(), String> {
        let rc = if self.g2 { g16_g2_msm_async(self.ctx.raw, self.raw, scalars, n, dev_out_affine, dev_out_partial) }
                 else { g16_g1_msm_async(self.ctx.raw, self.raw, scalars, n, dev_out_affine, dev_out_partial) };
        check(self.ctx.raw, rc)
    }
}
impl<'a> Drop for DeviceBases<'a> { fn drop(&mut self) { unsafe { g16_bases_free(self.raw) } } }

/// Fold `k` projective partial sums (device) into one affine point (device).
/// # Safety
/// Device pointers of the context's device: `k` x 48 (G1) / 96 (G2) u32 in, 25 / 49 u32 out.
pub unsafe fn combine_partials_device(ctx: &Context, g2: bool, dev_partials: *const c_void, k: usize, dev_out_affine: *mut c_void) -> Result<(), String> {
    let rc = if g2 { g16_g2_combine_partials_device(ctx.raw, dev_partials, k, dev_out_affine) } else { g16_g1_combine_partials_device(ctx.raw, dev_partials, k, dev_out_affine) };
    check(ctx.raw, rc)
}

// ---- setup: fixed-base batch scalar multiplication ------------------------------------------------------
/// `scalars.iter().map(|s| (base * s).into_affine())` in one batched call.
pub fn fixed_base_mul_g1(ctx: &Context, base: &G1Affine, scalars: &[Fr]) -> Result<Vec<G1Affine>, String> {
    let (bxy, _) = pack_g1(std::slice::from_ref(base));
    let sc = pack_scalars(scalars);
    let (mut out, mut inf) = (vec![0u64; 12 * scalars.len()], vec![0u8; scalars.len()]);
    check(ctx.raw, unsafe { g16_g1_fixed_base_mul(ctx.raw, bxy.as_ptr(), sc.as_ptr(), scalars.len(), out.as_mut_ptr(), inf.as_mut_ptr()) })?;
    Ok(g1_vec(&out, &inf))
}
pub fn fixed_base_mul_g2(ctx: &Context, base: &G2Affine, scalars: &[Fr]) -> Result<Vec<G2Affine>, String> {
    let (bxy, _) = pack_g2(std::slice::from_ref(base));
    let sc = pack_scalars(scalars);
    let (mut out, mut inf) = (vec![0u64; 24 * scalars.len()], vec![0u8; scalars.len()]);
    check(ctx.raw, unsafe { g16_g2_fixed_base_mul(ctx.raw, bxy.as_ptr(), sc.as_ptr(), scalars.len(), out.as_mut_ptr(), inf.as_mut_ptr()) })?;
    Ok(g2_vec(&out, &inf))
}
/// Device scalars in, packed device points out ((0,0) = identity); single-device context.
/// # Safety
/// `dev_scalars`: n x 4 u64 on the device; `dev_out_xy`: n x 12 u64 on the device.
pub unsafe fn fixed_base_mul_g1_device(ctx: &Context, base: &G1Affine, dev_scalars: *const c_void, n: usize, dev_out_xy: *mut c_void) -> Result<(), String> {
    let (xy, _) = pack_g1(std::slice::from_ref(base));
    check(ctx.raw, g16_g1_fixed_base_mul_device(ctx.raw, xy.as_ptr(), dev_scalars, n, dev_out_xy))
}
/// # Safety
/// `dev_scalars`: n x 4 u64 on the device; `dev_out_xy`: n x 24 u64 on the device.
pub unsafe fn fixed_base_mul_g2_device(ctx: &Context, base: &G2Affine, dev_scalars: *const c_void, n: usize, dev_out_xy: *mut c_void) -> Result<(), String> {
    let (xy, _) = pack_g2(std::slice::from_ref(base));
    check(ctx.raw, g16_g2_fixed_base_mul_device(ctx.raw, xy.as_ptr(), dev_scalars, n, dev_out_xy))
}

// ---- the whole ProvingKey resident: Prover::prove's group part in one call ------------------------------------
/// Borrowed view of the fields of `groth16_setup::ProvingKey<F>` the prover's MSMs read
/// (`crates/groth16-setup/src/lib.rs:27-52`); build it with `ProvingKeyView::from(&pk)` style field access in the caller.
pub struct ProvingKeyView<'k> {
    pub alpha_g1: &'k G1Affine, pub beta_g1: &'k G1Affine, pub delta_g1: &'k G1Affine,
    pub beta_g2: &'k G2Affine, pub delta_g2: &'k G2Affine,
    pub a_g1: &'k [G1Affine], pub b_g1: &'k [G1Affine], pub b_g2: &'k [G2Affine],
    pub ic_g1: &'k [G1Affine], pub h_g1: &'k [G1Affine],
    pub num_public: usize,
}
/// (proof.a, proof.b, proof.c) in the order of `struct Proof` (`crates/groth16-core/src/lib.rs:27-36`).
pub type ProofPoints = (G1Affine, G2Affine, G1Affine);

pub struct DeviceProvingKey<'a> { ctx: &'a Context, raw: *mut g16_pk }
impl<'a> DeviceProvingKey<'a> {
    /// Upload once per key (`g16_pk_upload`); the ad-hoc single points (alpha, beta, delta) ride in front of the arrays.
    pub fn upload(ctx: &'a Context, pk: &ProvingKeyView) -> Result<Self, String> {
        let (al, be1, de1) = (pack_single_g1(pk.alpha_g1), pack_single_g1(pk.beta_g1), pack_single_g1(pk.delta_g1));
        let (be2, de2) = (pack_single_g2(pk.beta_g2), pack_single_g2(pk.delta_g2));
        let (a, a_inf) = pack_g1(pk.a_g1);
        let (b1, b1_inf) = pack_g1(pk.b_g1);
        let (b2, b2_inf) = pack_g2(pk.b_g2);
        let (ic, ic_inf) = pack_g1(pk.ic_g1);
        let (h, h_inf) = pack_g1(pk.h_g1);
        let host = g16_pk_host {
            alpha_g1: al.as_ptr(), beta_g1: be1.as_ptr(), delta_g1: de1.as_ptr(), beta_g2: be2.as_ptr(), delta_g2: de2.as_ptr(),
            a_g1: a.as_ptr(), a_g1_inf: a_inf.as_ptr(), a_len: pk.a_g1.len(),
            b_g1: b1.as_ptr(), b_g1_inf: b1_inf.as_ptr(), b1_len: pk.b_g1.len(),
            b_g2: b2.as_ptr(), b_g2_inf: b2_inf.as_ptr(), b2_len: pk.b_g2.len(),
            ic_g1: ic.as_ptr(), ic_g1_inf: ic_inf.as_ptr(), ic_len: pk.ic_g1.len(),
            h_g1: h.as_ptr(), h_g1_inf: h_inf.as_ptr(), h_len: pk.h_g1.len(),
            num_public: pk.num_public,
        };
        let mut raw = std::ptr::null_mut();
        check(ctx.raw, unsafe { g16_pk_upload(ctx.raw, &host, &mut raw) })?;
        Ok(DeviceProvingKey { ctx, raw })
    }
    /// One-time tables of multiples for the five resident arrays (`g16_pk_precompute`).
    pub fn precompute(&mut self) -> Result<(), String> { check(self.ctx.raw, unsafe { g16_pk_precompute(self.ctx.raw, self.raw) }) }
    /// The same for scalars promised to be below `2^scalar_bits` (`g16_pk_precompute_bits`): `Prover::prove` truncates the
    /// assignment and the H coefficients to 64 bits (`crates/groth16-core/src/lib.rs:156-161,203-208`), so the drop-in
    /// passes 64.  A tuning hint only: wider scalars (r, s) stay correct.
    pub fn precompute_for_bits(&mut self, scalar_bits: u32) -> Result<(), String> {
        check(self.ctx.raw, unsafe { g16_pk_precompute_bits(self.ctx.raw, self.raw, scalar_bits as c_uint) })
    }
    /// The group part of `Prover::prove` (`lib.rs:164-271`): `assignment_fr` = the truncated assignment of `:156-161`,
    /// `h_coeffs` = the truncated quotient coefficients of `:203-208`, `r`, `s` of `:152-153`.
    pub fn prove(&self, assignment_fr: &[Fr], h_coeffs: &[Fr], r: &Fr, s: &Fr) -> Result<ProofPoints, String> {
        let (w, h) = (pack_scalars(assignment_fr), pack_scalars(h_coeffs));
        let (mut a, mut b, mut c, mut fl) = ([0u64; 12], [0u64; 24], [0u64; 12], [0u8; 3]);
        let flp = fl.as_mut_ptr();
        check(self.ctx.raw, unsafe {
            g16_prove(self.ctx.raw, self.raw, w.as_ptr(), assignment_fr.len(), if h_coeffs.is_empty() { std::ptr::null() } else { h.as_ptr() },
                      h_coeffs.len(), (r.0).0.as_ptr(), (s.0).0.as_ptr(), a.as_mut_ptr(), flp, b.as_mut_ptr(), flp.add(1), c.as_mut_ptr(), flp.add(2))
        })?;
        Ok((g1_from(&a, fl[0]), g2_from(&b, fl[1]), g1_from(&c, fl[2])))
    }
    /// `Prover::prove` (`lib.rs:139-272`) from the UN-truncated witness and the sparse constraint system: validation,
    /// quotient polynomial, truncations and the five MSMs chained on the device.  `Err("Invalid witness: ..")` /
    /// `Err("Polynomial division failed ..")` mirror the reference's error strings.
    pub fn prove_r1cs(&self, r1cs: &DeviceR1cs, assignment: &[Fr], r: &Fr, s: &Fr) -> Result<ProofPoints, String> {
        let w = pack_scalars(assignment);
        let (mut a, mut b, mut c, mut fl) = ([0u64; 12], [0u64; 24], [0u64; 12], [0u8; 3]);
        let flp = fl.as_mut_ptr();
        check(self.ctx.raw, unsafe {
            g16_prove_r1cs(self.ctx.raw, self.raw, r1cs.raw, w.as_ptr(), assignment.len(), (r.0).0.as_ptr(), (s.0).0.as_ptr(),
                           a.as_mut_ptr(), flp, b.as_mut_ptr(), flp.add(1), c.as_mut_ptr(), flp.add(2))
        })?;
        Ok((g1_from(&a, fl[0]), g2_from(&b, fl[1]), g1_from(&c, fl[2])))
    }
}
impl<'a> Drop for DeviceProvingKey<'a> { fn drop(&mut self) { unsafe { g16_pk_free(self.raw) } } }

// ---- quotient polynomial (the stage that feeds the H MSM) ---------------------------------------------------
/// `QAP::compute_quotient_polynomial` (`crates/groth16-qap/src/lib.rs:225-271`) from the domain evaluations of
/// A, B, C (`n` = power of two): coefficients of H, `Err("Polynomial division failed")` like the reference.
pub fn quotient_h(ctx: &Context, a: &[Fr], b: &[Fr], c: &[Fr]) -> Result<Vec<Fr>, String> {
    let n = a.len();
    if b.len() != n || c.len() != n { return Err("length mismatch".into()); }
    let (pa, pb, pc) = (pack_scalars(a), pack_scalars(b), pack_scalars(c));
    let mut out = vec![0u64; 4 * n];
    check(ctx.raw, unsafe { g16_quotient_h(ctx.raw, pa.as_ptr(), pb.as_ptr(), pc.as_ptr(), n, out.as_mut_ptr()) })?;
    Ok(unpack_scalars(&out))
}

/// Device in / device out form of `quotient_h`, asynchronous on the context stream.
/// # Safety
/// `dev_abc`: 3 n x 4 u64 on the device (overwritten), `dev_h`: n x 4 u64, `dev_bad_rows`: one u32, all on the context's device.
pub unsafe fn quotient_h_device(ctx: &Context, dev_abc: *mut c_void, n: usize, dev_h: *mut c_void, dev_bad_rows: *mut c_void) -> Result<(), String> {
    check(ctx.raw, g16_quotient_h_device(ctx.raw, dev_abc, n, dev_h, dev_bad_rows))
}

// ---- sparse R1CS: setup and prove for real circuits ---------------------------------------------------------
/// One constraint matrix in CSR form over the constraints: `row_ptr[num_constraints + 1]`, `col[nnz]` (variable
/// indices), `val[nnz]`.  Build it from `R1CS::constraints[i].a.terms` etc. (`crates/groth16-r1cs`).
pub struct Csr<'m> { pub row_ptr: &'m [u32], pub col: &'m [u32], pub val: &'m [Fr] }
/// Host copy of what `CRS::generate_from_qap` returns (`crates/groth16-setup/src/lib.rs:141-268`).
pub struct CrsHost {
    pub alpha_g1: G1Affine, pub beta_g1: G1Affine, pub delta_g1: G1Affine,
    pub beta_g2: G2Affine, pub gamma_g2: G2Affine, pub delta_g2: G2Affine,
    pub a_g1: Vec<G1Affine>, pub b_g1: Vec<G1Affine>, pub b_g2: Vec<G2Affine>,
    pub ic_g1: Vec<G1Affine>, pub vk_ic_g1: Vec<G1Affine>, pub h_g1: Vec<G1Affine>,
}
pub struct DeviceR1cs<'a> { ctx: &'a Context, raw: *mut g16_r1cs, num_variables: usize }
impl<'a> DeviceR1cs<'a> {
    pub fn upload(ctx: &'a Context, num_constraints: usize, num_variables: usize, a: &Csr, b: &Csr, c: &Csr) -> Result<Self, String> {
        for m in [a, b, c] {
            if m.row_ptr.len() != num_constraints + 1 || m.col.len() != m.val.len() { return Err("CSR arrays are inconsistent".into()); }
        }
        let vals = [pack_scalars(a.val), pack_scalars(b.val), pack_scalars(c.val)];
        let raw_csr = |m: &Csr, v: &Vec<u64>| g16_csr { row_ptr: m.row_ptr.as_ptr(), col: m.col.as_ptr(), val: v.as_ptr() };
        let (ca, cb, cc) = (raw_csr(a, &vals[0]), raw_csr(b, &vals[1]), raw_csr(c, &vals[2]));
        let mut raw = std::ptr::null_mut();
        check(ctx.raw, unsafe { g16_r1cs_upload(ctx.raw, num_constraints, num_variables, &ca, &cb, &cc, &mut raw) })?;
        Ok(DeviceR1cs { ctx, raw, num_variables })
    }
    /// `num_constraints.next_power_of_two()` (`QAP::from_r1cs`, qap/src/lib.rs:100).
    pub fn domain_size(&self) -> usize { unsafe { g16_r1cs_domain_size(self.raw) } }
    /// (<A-row i, w>, <B-row i, w>, <C-row i, w>) on the domain.
    pub fn domain_evals(&self, assignment: &[Fr]) -> Result<(Vec<Fr>, Vec<Fr>, Vec<Fr>), String> {
        let n = self.domain_size();
        let w = pack_scalars(assignment);
        let (mut a, mut b, mut c) = (vec![0u64; 4 * n], vec![0u64; 4 * n], vec![0u64; 4 * n]);
        check(self.ctx.raw, unsafe { g16_r1cs_domain_evals(self.ctx.raw, self.raw, w.as_ptr(), assignment.len(), a.as_mut_ptr(), b.as_mut_ptr(), c.as_mut_ptr()) })?;
        Ok((unpack_scalars(&a), unpack_scalars(&b), unpack_scalars(&c)))
    }
    /// (A_j(s), B_j(s), C_j(s)) for every variable j (`qap.a_polys[j].evaluate(&s)`, setup/src/lib.rs:174-182).
    pub fn eval_at(&self, s: &Fr) -> Result<(Vec<Fr>, Vec<Fr>, Vec<Fr>), String> {
        let nv = self.num_variables;
        let (mut a, mut b, mut c) = (vec![0u64; 4 * nv], vec![0u64; 4 * nv], vec![0u64; 4 * nv]);
        check(self.ctx.raw, unsafe { g16_r1cs_eval_at(self.ctx.raw, self.raw, (s.0).0.as_ptr(), a.as_mut_ptr(), b.as_mut_ptr(), c.as_mut_ptr()) })?;
        Ok((unpack_scalars(&a), unpack_scalars(&b), unpack_scalars(&c)))
    }
    /// `CRS::generate_from_qap`: returns the host CRS (when `want_host`) and the device-resident proving key, ready for
    /// `DeviceProvingKey::prove` / `prove_r1cs` without a host round trip.  `params` = [alpha, beta, gamma, delta, s].
    pub fn setup_crs(&self, params: &[Fr; 5], num_public: usize, want_host: bool) -> Result<(Option<CrsHost>, DeviceProvingKey<'a>), String> {
        let (nv, n) = (self.num_variables, self.domain_size());
        let n_ic = nv.saturating_sub(num_public + 1);
        let n_vk = (num_public + 1).min(nv);
        let mut singles1 = vec![0u64; 3 * 12];
        let mut singles2 = vec![0u64; 3 * 24];
        let mut bufs: Vec<(Vec<u64>, Vec<u8>)> = [(nv, 12), (nv, 12), (nv, 24), (n_ic, 12), (n_vk, 12), (n, 12)]
            .iter().map(|&(len, w)| (vec![0u64; len * w], vec![0u8; len])).collect();
        let mut host = g16_crs_host {
            alpha_g1: std::ptr::null_mut(), beta_g1: std::ptr::null_mut(), delta_g1: std::ptr::null_mut(),
            beta_g2: std::ptr::null_mut(), gamma_g2: std::ptr::null_mut(), delta_g2: std::ptr::null_mut(),
            a_g1: std::ptr::null_mut(), a_g1_inf: std::ptr::null_mut(), b_g1: std::ptr::null_mut(), b_g1_inf: std::ptr::null_mut(),
            b_g2: std::ptr::null_mut(), b_g2_inf: std::ptr::null_mut(), ic_g1: std::ptr::null_mut(), ic_g1_inf: std::ptr::null_mut(),
            vk_ic_g1: std::ptr::null_mut(), vk_ic_g1_inf: std::ptr::null_mut(), h_g1: std::ptr::null_mut(), h_g1_inf: std::ptr::null_mut(),
        };
        if want_host {
            unsafe {
                host.alpha_g1 = singles1.as_mut_ptr(); host.beta_g1 = singles1.as_mut_ptr().add(12); host.delta_g1 = singles1.as_mut_ptr().add(24);
                host.beta_g2 = singles2.as_mut_ptr(); host.gamma_g2 = singles2.as_mut_ptr().add(24); host.delta_g2 = singles2.as_mut_ptr().add(48);
            }
            host.a_g1 = bufs[0].0.as_mut_ptr(); host.a_g1_inf = bufs[0].1.as_mut_ptr();
            host.b_g1 = bufs[1].0.as_mut_ptr(); host.b_g1_inf = bufs[1].1.as_mut_ptr();
            host.b_g2 = bufs[2].0.as_mut_ptr(); host.b_g2_inf = bufs[2].1.as_mut_ptr();
            host.ic_g1 = bufs[3].0.as_mut_ptr(); host.ic_g1_inf = bufs[3].1.as_mut_ptr();
            host.vk_ic_g1 = bufs[4].0.as_mut_ptr(); host.vk_ic_g1_inf = bufs[4].1.as_mut_ptr();
            host.h_g1 = bufs[5].0.as_mut_ptr(); host.h_g1_inf = bufs[5].1.as_mut_ptr();
        }
        let mut pk = std::ptr::null_mut();
        let p: Vec<&[u64; 4]> = params.iter().map(|x| &(x.0).0).collect();
        check(self.ctx.raw, unsafe {
            g16_setup_crs(self.ctx.raw, self.raw, p[0].as_ptr(), p[1].as_ptr(), p[2].as_ptr(), p[3].as_ptr(), p[4].as_ptr(), num_public,
                          if want_host { &mut host } else { std::ptr::null_mut() }, &mut pk)
        })?;
        let crs = if want_host {
            Some(CrsHost {
                alpha_g1: g1_from(&singles1[0..12], 0), beta_g1: g1_from(&singles1[12..24], 0), delta_g1: g1_from(&singles1[24..36], 0),
                beta_g2: g2_from(&singles2[0..24], 0), gamma_g2: g2_from(&singles2[24..48], 0), delta_g2: g2_from(&singles2[48..72], 0),
                a_g1: g1_vec(&bufs[0].0, &bufs[0].1), b_g1: g1_vec(&bufs[1].0, &bufs[1].1), b_g2: g2_vec(&bufs[2].0, &bufs[2].1),
                ic_g1: g1_vec(&bufs[3].0, &bufs[3].1), vk_ic_g1: g1_vec(&bufs[4].0, &bufs[4].1), h_g1: g1_vec(&bufs[5].0, &bufs[5].1),
            })
        } else { None };
        Ok((crs, DeviceProvingKey { ctx: self.ctx, raw: pk }))
    }
}
impl<'a> Drop for DeviceR1cs<'a> { fn drop(&mut self) { unsafe { g16_r1cs_free(self.raw) } } }

// ---- wire format: the bytes `CanonicalSerialize` writes for Vec<G1Affine> / Vec<G2Affine> bodies ------------
/// Same bytes as `for p in points { p.serialize_with_mode(&mut w, compress) }` (ark-bls12-381's Zcash encoding).
pub fn serialize_g1(ctx: &Context, points: &[G1Affine], compressed: bool) -> Result<Vec<u8>, String> {
    let (xy, inf) = pack_g1(points);
    let mut out = vec![0u8; points.len() * if compressed { 48 } else { 96 }];
    check(ctx.raw, unsafe { g16_g1_serialize(ctx.raw, xy.as_ptr(), inf.as_ptr(), points.len(), compressed as c_int, out.as_mut_ptr()) })?;
    Ok(out)
}
pub fn serialize_g2(ctx: &Context, points: &[G2Affine], compressed: bool) -> Result<Vec<u8>, String> {
    let (xy, inf) = pack_g2(points);
    let mut out = vec![0u8; points.len() * if compressed { 96 } else { 192 }];
    check(ctx.raw, unsafe { g16_g2_serialize(ctx.raw, xy.as_ptr(), inf.as_ptr(), points.len(), compressed as c_int, out.as_mut_ptr()) })?;
    Ok(out)
}
/// `G1Affine::deserialize_with_mode(.., compress, validate)` per element; `Err("InvalidData ..")` /
/// `Err("UnexpectedFlags ..")` name the first rejected element like ark's SerializationError.
pub fn deserialize_g1(ctx: &Context, bytes: &[u8], compressed: bool, validate: bool) -> Result<Vec<G1Affine>, String> {
    let per = if compressed { 48 } else { 96 };
    if bytes.len() % per != 0 { return Err("InvalidData: truncated input".into()); }
    let n = bytes.len() / per;
    let (mut out, mut inf) = (vec![0u64; 12 * n], vec![0u8; n]);
    check(ctx.raw, unsafe { g16_g1_deserialize(ctx.raw, bytes.as_ptr(), n, compressed as c_int, validate as c_int, out.as_mut_ptr(), inf.as_mut_ptr(), std::ptr::null_mut()) })?;
    Ok(g1_vec(&out, &inf))
}
pub fn deserialize_g2(ctx: &Context, bytes: &[u8], compressed: bool, validate: bool) -> Result<Vec<G2Affine>, String> {
    let per = if compressed { 96 } else { 192 };
    if bytes.len() % per != 0 { return Err("InvalidData: truncated input".into()); }
    let n = bytes.len() / per;
    let (mut out, mut inf) = (vec![0u64; 24 * n], vec![0u8; n]);
    check(ctx.raw, unsafe { g16_g2_deserialize(ctx.raw, bytes.as_ptr(), n, compressed as c_int, validate as c_int, out.as_mut_ptr(), inf.as_mut_ptr(), std::ptr::null_mut()) })?;
    Ok(g2_vec(&out, &inf))
}
/// `Proof::serialize_compressed` / `serialize_uncompressed` (192 / 384 bytes).
pub fn proof_to_bytes(ctx: &Context, proof: &ProofPoints, compressed: bool) -> Result<Vec<u8>, String> {
    let (a, ai) = pack_g1(std::slice::from_ref(&proof.0));
    let (b, bi) = pack_g2(std::slice::from_ref(&proof.1));
    let (c, ci) = pack_g1(std::slice::from_ref(&proof.2));
    let mut out = vec![0u8; if compressed { 192 } else { 384 }];
    check(ctx.raw, unsafe { g16_proof_serialize(ctx.raw, a.as_ptr(), ai[0], b.as_ptr(), bi[0], c.as_ptr(), ci[0], compressed as c_int, out.as_mut_ptr()) })?;
    Ok(out)
}
/// `Proof::deserialize_compressed` / `_uncompressed` (`validate` = ark's `Validate::Yes`).
pub fn proof_from_bytes(ctx: &Context, bytes: &[u8], compressed: bool, validate: bool) -> Result<ProofPoints, String> {
    if bytes.len() != if compressed { 192 } else { 384 } { return Err("InvalidData: a proof is 192 bytes compressed, 384 uncompressed".into()); }
    let (mut a, mut b, mut c, mut fl) = ([0u64; 12], [0u64; 24], [0u64; 12], [0u8; 3]);
    let flp = fl.as_mut_ptr();
    check(ctx.raw, unsafe {
        g16_proof_deserialize(ctx.raw, bytes.as_ptr(), compressed as c_int, validate as c_int, a.as_mut_ptr(), flp, b.as_mut_ptr(), flp.add(1), c.as_mut_ptr(), flp.add(2))
    })?;
    Ok((g1_from(&a, fl[0]), g2_from(&b, fl[1]), g1_from(&c, fl[2])))
}

// ---- fixed randomness through the UNCHANGED Prover::prove(pk, witness, rng) API ---------------------------------
/// `Fr::rand(rng)` (ark-ff 0.4.2, `Distribution<Fp> for Standard`) draws four `u64` limbs with `rng.next_u64()`, masks
/// the top limb to 255 bits and takes them AS the Montgomery representation (retrying while >= r).  `Prover::prove`
/// draws r, then s (`crates/groth16-core/src/lib.rs:152-153`).  This generator replays the raw limbs of the wanted
/// (r, s), so `Prover::prove(&pk, &witness, &mut FixedLimbsRng::for_r_s(&r, &s))` uses exactly those values; after the
/// eight limbs it keeps cycling (never called by `prove`).  It implements `rand_core::RngCore` (rand 0.8 / ark-std 0.4's
/// `Rng` blanket impl), which is what the `R: Rng + ?Sized` bound of `prove` (`lib.rs:139-147`) needs.
pub struct FixedLimbsRng { limbs: Vec<u64>, pos: usize }
impl FixedLimbsRng {
    pub fn for_r_s(r: &Fr, s: &Fr) -> Self { let mut l = (r.0).0.to_vec(); l.extend_from_slice(&(s.0).0); FixedLimbsRng { limbs: l, pos: 0 } }
    pub fn from_limbs(limbs: Vec<u64>) -> Self { assert!(!limbs.is_empty()); FixedLimbsRng { limbs, pos: 0 } }
}
impl rand_core::RngCore for FixedLimbsRng {
    fn next_u64(&mut self) -> u64 { let v = self.limbs[self.pos % self.limbs.len()]; self.pos += 1; v }
    fn next_u32(&mut self) -> u32 { self.next_u64() as u32 }
    fn fill_bytes(&mut self, dest: &mut [u8]) {
        for chunk in dest.chunks_mut(8) { let b = self.next_u64().to_le_bytes(); chunk.copy_from_slice(&b[..chunk.len()]); }
    }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), rand_core::Error> { self.fill_bytes(dest); Ok(()) }
}

// ---- test hooks (parity tests of the device arithmetic) -------------------------------------------------------
/// Element-wise Fq ops on the device: op 0 mul, 1 add, 2 sub, 3 inverse(a), 4 square(a), 5 negate(a).
pub fn debug_fq_op(ctx: &Context, op: i32, a: &[Fq], b: Option<&[Fq]>) -> Result<Vec<Fq>, String> {
    let pa: Vec<u64> = a.iter().flat_map(|x| (x.0).0).collect();
    let pb: Option<Vec<u64>> = b.map(|v| v.iter().flat_map(|x| (x.0).0).collect());
    let mut out = vec![0u64; 6 * a.len()];
    check(ctx.raw, unsafe { g16_debug_fq_op(ctx.raw, op as c_int, pa.as_ptr(), pb.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), out.as_mut_ptr(), a.len()) })?;
    Ok(out.chunks_exact(6).map(fq_from).collect())
}
/// `Fr::into_bigint()` on the device (canonical limbs).
pub fn debug_fr_from_mont(ctx: &Context, a: &[Fr]) -> Result<Vec<[u64; 4]>, String> {
    let pa = pack_scalars(a);
    let mut out = vec![0u64; 4 * a.len()];
    check(ctx.raw, unsafe { g16_debug_fr_from_mont(ctx.raw, pa.as_ptr(), out.as_mut_ptr(), a.len()) })?;
    Ok(out.chunks_exact(4).map(|l| [l[0], l[1], l[2], l[3]]).collect())
}
/// out[i] = p[i] + q[i] through the device's XYZZ mixed addition (all exceptional cases).
pub fn debug_g1_add(ctx: &Context, p: &[G1Affine], q: &[G1Affine]) -> Result<Vec<G1Affine>, String> {
    if p.len() != q.len() { return Err("length mismatch".into()); }
    let ((pxy, pinf), (qxy, qinf)) = (pack_g1(p), pack_g1(q));
    let (mut out, mut inf) = (vec![0u64; 12 * p.len()], vec![0u8; p.len()]);
    check(ctx.raw, unsafe { g16_debug_g1_add(ctx.raw, pxy.as_ptr(), pinf.as_ptr(), qxy.as_ptr(), qinf.as_ptr(), out.as_mut_ptr(), inf.as_mut_ptr(), p.len()) })?;
    Ok(g1_vec(&out, &inf))
}
pub fn debug_g2_add(ctx: &Context, p: &[G2Affine], q: &[G2Affine]) -> Result<Vec<G2Affine>, String> {
    if p.len() != q.len() { return Err("length mismatch".into()); }
    let ((pxy, pinf), (qxy, qinf)) = (pack_g2(p), pack_g2(q));
    let (mut out, mut inf) = (vec![0u64; 24 * p.len()], vec![0u8; p.len()]);
    check(ctx.raw, unsafe { g16_debug_g2_add(ctx.raw, pxy.as_ptr(), pinf.as_ptr(), qxy.as_ptr(), qinf.as_ptr(), out.as_mut_ptr(), inf.as_mut_ptr(), p.len()) })?;
    Ok(g2_vec(&out, &inf))
}
