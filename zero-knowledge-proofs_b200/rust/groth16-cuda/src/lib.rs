//! `groth16-cuda`: safe Rust wrapper of the B200 MSM engine (C ABI in `include/g16_cuda.h`).
//!
//! Drop-in seam (reference = vats98754/zero-knowledge-proofs):
//! * `msm_g1` / `msm_g2` replace `G1Projective::msm(&points, &scalars)` / `G2Projective::msm` at
//!   `crates/groth16-core/src/lib.rs:282,296` (inside `Prover::multi_scalar_mult_g1/_g2`, `:275-300`);
//! * `fixed_base_mul_g1/_g2` replace the `par_iter().map(|v| (gen * fr).into_affine())` blocks of
//!   `CRS::generate_from_qap` at `crates/groth16-setup/src/lib.rs:185-241`;
//! * `DeviceBases` keeps CRS arrays resident between proofs.
//!
//! Marshalling copies ark's raw limbs (`p.x.0.0`, already Montgomery) -- no field arithmetic on the host.
//! Errors become `String`s that the callers map to `GrothError::MSMError(..)` exactly like `lib.rs:283`.

use ark_bls12_381::{Fq, Fq2, Fr, G1Affine, G2Affine};
use ark_ff::{BigInt, Fp};
use std::marker::PhantomData;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct RawCtx { _p: [u8; 0] }
#[repr(C)]
pub struct RawBases { _p: [u8; 0] }

extern "C" {
    fn g16_ctx_create(devices: *const c_int, ndev: c_int, out: *mut *mut RawCtx) -> c_int;
    fn g16_ctx_destroy(ctx: *mut RawCtx);
    fn g16_last_error(ctx: *const RawCtx) -> *const c_char;
    fn g16_g1_bases_upload(ctx: *mut RawCtx, xy: *const u64, inf: *const u8, n: usize, out: *mut *mut RawBases) -> c_int;
    fn g16_g2_bases_upload(ctx: *mut RawCtx, xy: *const u64, inf: *const u8, n: usize, out: *mut *mut RawBases) -> c_int;
    fn g16_bases_free(b: *mut RawBases);
    fn g16_g1_msm(ctx: *mut RawCtx, b: *const RawBases, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    fn g16_g2_msm(ctx: *mut RawCtx, b: *const RawBases, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    fn g16_g1_msm_oneshot(ctx: *mut RawCtx, xy: *const u64, inf: *const u8, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    fn g16_g2_msm_oneshot(ctx: *mut RawCtx, xy: *const u64, inf: *const u8, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    fn g16_g1_fixed_base_mul(ctx: *mut RawCtx, base_xy: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    fn g16_g2_fixed_base_mul(ctx: *mut RawCtx, base_xy: *const u64, scalars: *const u64, n: usize, out_xy: *mut u64, out_inf: *mut u8) -> c_int;
    fn g16_quotient_h(ctx: *mut RawCtx, a: *const u64, b: *const u64, c: *const u64, n: usize, h: *mut u64) -> c_int;
    fn g16_g1_serialize(ctx: *mut RawCtx, xy: *const u64, inf: *const u8, n: usize, compressed: c_int, out: *mut u8) -> c_int;
    fn g16_g2_serialize(ctx: *mut RawCtx, xy: *const u64, inf: *const u8, n: usize, compressed: c_int, out: *mut u8) -> c_int;
    fn g16_g1_deserialize(ctx: *mut RawCtx, bytes: *const u8, n: usize, compressed: c_int, validate: c_int, out_xy: *mut u64, out_inf: *mut u8, status: *mut u8) -> c_int;
    fn g16_g2_deserialize(ctx: *mut RawCtx, bytes: *const u8, n: usize, compressed: c_int, validate: c_int, out_xy: *mut u64, out_inf: *mut u8, status: *mut u8) -> c_int;
}

/// One engine context (one or more GPUs of one box).  `!Sync`: one caller at a time, like the
/// single-threaded reference prover; wrap in a `Mutex` to share.
pub struct Context { raw: *mut RawCtx, _not_sync: PhantomData<*mut ()> }
unsafe impl Send for Context {}

fn err(ctx: *const RawCtx, code: c_int) -> String {
    let msg = unsafe { std::ffi::CStr::from_ptr(g16_last_error(ctx)) }.to_string_lossy().into_owned();
    format!("{} (code {})", msg, code)
}

impl Context {
    /// `devices = &[]` uses the current CUDA device; several devices shard every bases array by index range.
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { g16_ctx_create(if devices.is_empty() { std::ptr::null() } else { devices.as_ptr() }, devices.len() as c_int, &mut raw) };
        if rc != 0 { return Err(err(std::ptr::null(), rc)); }
        Ok(Context { raw, _not_sync: PhantomData })
    }
}
impl Drop for Context { fn drop(&mut self) { unsafe { g16_ctx_destroy(self.raw) } } }

// ---- marshalling: ark in-memory limbs <-> packed u64 buffers ------------------------------------
fn push_fq(out: &mut Vec<u64>, f: &Fq) { out.extend_from_slice(&(f.0).0); }
fn push_fq2(out: &mut Vec<u64>, f: &Fq2) { push_fq(out, &f.c0); push_fq(out, &f.c1); }
fn fq_from(l: &[u64]) -> Fq { let mut a = [0u64; 6]; a.copy_from_slice(l); Fp(BigInt(a), PhantomData) }
fn pack_scalars(s: &[Fr]) -> Vec<u64> { let mut v = Vec::with_capacity(4 * s.len()); for x in s { v.extend_from_slice(&(x.0).0); } v }
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
fn g1_from(xy: &[u64], inf: u8) -> G1Affine {
    if inf != 0 { return G1Affine::identity(); }
    G1Affine::new_unchecked(fq_from(&xy[0..6]), fq_from(&xy[6..12]))
}
fn g2_from(xy: &[u64], inf: u8) -> G2Affine {
    if inf != 0 { return G2Affine::identity(); }
    G2Affine::new_unchecked(Fq2::new(fq_from(&xy[0..6]), fq_from(&xy[6..12])), Fq2::new(fq_from(&xy[12..18]), fq_from(&xy[18..24])))
}

// ---- the seam of Prover::multi_scalar_mult_g1/_g2 ------------------------------------------------
/// Σ scalars[i]·bases[i] as an affine point.  Length mismatch -> `Err` (ark returns `Err(min_len)`).
pub fn msm_g1(ctx: &Context, bases: &[G1Affine], scalars: &[Fr]) -> Result<G1Affine, String> {
    if bases.len() != scalars.len() { return Err(format!("{}", bases.len().min(scalars.len()))); }
    let (xy, inf) = pack_g1(bases);
    let sc = pack_scalars(scalars);
    let (mut out, mut oinf) = ([0u64; 12], 0u8);
    let rc = unsafe { g16_g1_msm_oneshot(ctx.raw, xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), bases.len(), out.as_mut_ptr(), &mut oinf) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok(g1_from(&out, oinf))
}
pub fn msm_g2(ctx: &Context, bases: &[G2Affine], scalars: &[Fr]) -> Result<G2Affine, String> {
    if bases.len() != scalars.len() { return Err(format!("{}", bases.len().min(scalars.len()))); }
    let (xy, inf) = pack_g2(bases);
    let sc = pack_scalars(scalars);
    let (mut out, mut oinf) = ([0u64; 24], 0u8);
    let rc = unsafe { g16_g2_msm_oneshot(ctx.raw, xy.as_ptr(), inf.as_ptr(), sc.as_ptr(), bases.len(), out.as_mut_ptr(), &mut oinf) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok(g2_from(&out, oinf))
}

// ---- resident CRS arrays ----------------------------------------------------------------------------
pub struct DeviceBases<'a> { ctx: &'a Context, raw: *mut RawBases, g2: bool, len: usize }
impl<'a> DeviceBases<'a> {
    pub fn upload_g1(ctx: &'a Context, p: &[G1Affine]) -> Result<Self, String> {
        let (xy, inf) = pack_g1(p);
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { g16_g1_bases_upload(ctx.raw, xy.as_ptr(), inf.as_ptr(), p.len(), &mut raw) };
        if rc != 0 { return Err(err(ctx.raw, rc)); }
        Ok(DeviceBases { ctx, raw, g2: false, len: p.len() })
    }
    pub fn upload_g2(ctx: &'a Context, p: &[G2Affine]) -> Result<Self, String> {
        let (xy, inf) = pack_g2(p);
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { g16_g2_bases_upload(ctx.raw, xy.as_ptr(), inf.as_ptr(), p.len(), &mut raw) };
        if rc != 0 { return Err(err(ctx.raw, rc)); }
        Ok(DeviceBases { ctx, raw, g2: true, len: p.len() })
    }
    pub fn len(&self) -> usize { self.len }
    /// Σ scalars[i]·bases[i] over the first `scalars.len()` resident bases (zero scalars cost nothing).
    pub fn msm_g1(&self, scalars: &[Fr]) -> Result<G1Affine, String> {
        assert!(!self.g2);
        let sc = pack_scalars(scalars);
        let (mut out, mut oinf) = ([0u64; 12], 0u8);
        let rc = unsafe { g16_g1_msm(self.ctx.raw, self.raw, sc.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut oinf) };
        if rc != 0 { return Err(err(self.ctx.raw, rc)); }
        Ok(g1_from(&out, oinf))
    }
    pub fn msm_g2(&self, scalars: &[Fr]) -> Result<G2Affine, String> {
        assert!(self.g2);
        let sc = pack_scalars(scalars);
        let (mut out, mut oinf) = ([0u64; 24], 0u8);
        let rc = unsafe { g16_g2_msm(self.ctx.raw, self.raw, sc.as_ptr(), scalars.len(), out.as_mut_ptr(), &mut oinf) };
        if rc != 0 { return Err(err(self.ctx.raw, rc)); }
        Ok(g2_from(&out, oinf))
    }
}
impl<'a> Drop for DeviceBases<'a> { fn drop(&mut self) { unsafe { g16_bases_free(self.raw) } } }

// ---- setup: fixed-base batch scalar multiplication ------------------------------------------------------
/// `scalars.iter().map(|s| (base * s).into_affine())` in one batched call.
pub fn fixed_base_mul_g1(ctx: &Context, base: &G1Affine, scalars: &[Fr]) -> Result<Vec<G1Affine>, String> {
    let (bxy, _) = pack_g1(std::slice::from_ref(base));
    let sc = pack_scalars(scalars);
    let (mut out, mut inf) = (vec![0u64; 12 * scalars.len()], vec![0u8; scalars.len()]);
    let rc = unsafe { g16_g1_fixed_base_mul(ctx.raw, bxy.as_ptr(), sc.as_ptr(), scalars.len(), out.as_mut_ptr(), inf.as_mut_ptr()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok((0..scalars.len()).map(|i| g1_from(&out[12 * i..12 * i + 12], inf[i])).collect())
}
pub fn fixed_base_mul_g2(ctx: &Context, base: &G2Affine, scalars: &[Fr]) -> Result<Vec<G2Affine>, String> {
    let (bxy, _) = pack_g2(std::slice::from_ref(base));
    let sc = pack_scalars(scalars);
    let (mut out, mut inf) = (vec![0u64; 24 * scalars.len()], vec![0u8; scalars.len()]);
    let rc = unsafe { g16_g2_fixed_base_mul(ctx.raw, bxy.as_ptr(), sc.as_ptr(), scalars.len(), out.as_mut_ptr(), inf.as_mut_ptr()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok((0..scalars.len()).map(|i| g2_from(&out[24 * i..24 * i + 24], inf[i])).collect())
}

// ---- quotient polynomial (the stage that feeds the H MSM) ---------------------------------------------------
/// `QAP::compute_quotient_polynomial` (`crates/groth16-qap/src/lib.rs:225-271`) from the domain evaluations of
/// A, B, C (`n` = power of two): coefficients of H, `Err("Polynomial division failed")` like the reference.
pub fn quotient_h(ctx: &Context, a: &[Fr], b: &[Fr], c: &[Fr]) -> Result<Vec<Fr>, String> {
    let n = a.len();
    if b.len() != n || c.len() != n { return Err("length mismatch".into()); }
    let (pa, pb, pc) = (pack_scalars(a), pack_scalars(b), pack_scalars(c));
    let mut out = vec![0u64; 4 * n];
    let rc = unsafe { g16_quotient_h(ctx.raw, pa.as_ptr(), pb.as_ptr(), pc.as_ptr(), n, out.as_mut_ptr()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok((0..n).map(|i| Fp(BigInt([out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]]), PhantomData)).collect())
}

// ---- wire format: the bytes `CanonicalSerialize` writes for Vec<G1Affine> / Vec<G2Affine> bodies ------------
/// Same bytes as `for p in points { p.serialize_with_mode(&mut w, compress) }` (ark-bls12-381's Zcash encoding).
pub fn serialize_g1(ctx: &Context, points: &[G1Affine], compressed: bool) -> Result<Vec<u8>, String> {
    let (xy, inf) = pack_g1(points);
    let mut out = vec![0u8; points.len() * if compressed { 48 } else { 96 }];
    let rc = unsafe { g16_g1_serialize(ctx.raw, xy.as_ptr(), inf.as_ptr(), points.len(), compressed as c_int, out.as_mut_ptr()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok(out)
}
pub fn serialize_g2(ctx: &Context, points: &[G2Affine], compressed: bool) -> Result<Vec<u8>, String> {
    let (xy, inf) = pack_g2(points);
    let mut out = vec![0u8; points.len() * if compressed { 96 } else { 192 }];
    let rc = unsafe { g16_g2_serialize(ctx.raw, xy.as_ptr(), inf.as_ptr(), points.len(), compressed as c_int, out.as_mut_ptr()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok(out)
}
/// `G1Affine::deserialize_with_mode(.., compress, validate)` per element; `Err("InvalidData ..")` /
/// `Err("UnexpectedFlags ..")` name the first rejected element like ark's SerializationError.
pub fn deserialize_g1(ctx: &Context, bytes: &[u8], compressed: bool, validate: bool) -> Result<Vec<G1Affine>, String> {
    let per = if compressed { 48 } else { 96 };
    if bytes.len() % per != 0 { return Err("InvalidData: truncated input".into()); }
    let n = bytes.len() / per;
    let (mut out, mut inf) = (vec![0u64; 12 * n], vec![0u8; n]);
    let rc = unsafe { g16_g1_deserialize(ctx.raw, bytes.as_ptr(), n, compressed as c_int, validate as c_int, out.as_mut_ptr(), inf.as_mut_ptr(), std::ptr::null_mut()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok((0..n).map(|i| g1_from(&out[12 * i..12 * i + 12], inf[i])).collect())
}
pub fn deserialize_g2(ctx: &Context, bytes: &[u8], compressed: bool, validate: bool) -> Result<Vec<G2Affine>, String> {
    let per = if compressed { 96 } else { 192 };
    if bytes.len() % per != 0 { return Err("InvalidData: truncated input".into()); }
    let n = bytes.len() / per;
    let (mut out, mut inf) = (vec![0u64; 24 * n], vec![0u8; n]);
    let rc = unsafe { g16_g2_deserialize(ctx.raw, bytes.as_ptr(), n, compressed as c_int, validate as c_int, out.as_mut_ptr(), inf.as_mut_ptr(), std::ptr::null_mut()) };
    if rc != 0 { return Err(err(ctx.raw, rc)); }
    Ok((0..n).map(|i| g2_from(&out[24 * i..24 * i + 24], inf[i])).collect())
}

/// Fixed (r, s) through the unchanged `Prover::prove(pk, witness, rng)` API: `Fr::rand` takes four
/// `next_u64` limbs as the Montgomery representation (top limb masked to 255 bits, rejected if >= r),
/// r first, then s (`crates/groth16-core/src/lib.rs:152-153`).  Feed this RNG with those 8 limbs.
pub struct FixedLimbsRng { pub limbs: Vec<u64>, pub pos: usize }
impl FixedLimbsRng {
    pub fn for_r_s(r: &Fr, s: &Fr) -> Self { let mut l = (r.0).0.to_vec(); l.extend_from_slice(&(s.0).0); FixedLimbsRng { limbs: l, pos: 0 } }
    pub fn next_u64(&mut self) -> u64 { let v = self.limbs[self.pos % self.limbs.len()]; self.pos += 1; v }
}

#[allow(dead_code)]
fn _unused(_: *const c_void) {}
