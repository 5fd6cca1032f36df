"""groth16_cuda -- Python host binding of libg16cuda.so (the B200 MSM engine).

This mirrors the Rust `groth16-cuda` crate (../rust/groth16-cuda) one to one: it only marshals
ark-layout limbs (numpy uint64) across the C ABI of include/g16_cuda.h.  Names follow the
reference seam:

    Context.multi_scalar_mult_g1 / _g2   <- Prover::multi_scalar_mult_g1/_g2
                                            (/root/reference/crates/groth16-core/src/lib.rs:275-300)
    Context.fixed_base_mul_g1 / _g2      <- `(gen * fr).into_affine()` blocks of
                                            CRS::generate_from_qap (crates/groth16-setup/src/lib.rs:185-241)
    Context.prove                        <- group part of Prover::prove (lib.rs:164-271)

There is no CPU fallback: importing works anywhere, but `Context()` raises `MSMError` unless the
CUDA library is present and a device is visible.
"""
from __future__ import annotations

import ctypes
import weakref
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(os.path.dirname(_PKG), "lib", "libg16cuda.so")

G16_OK, G16_ERR_INVALID, G16_ERR_CUDA, G16_ERR_NO_DEVICE, G16_ERR_OOM, G16_ERR_LENGTH = range(6)
G1_WORDS64, G2_WORDS64 = 12, 24
G1_PARTIAL_WORDS, G2_PARTIAL_WORDS = 48, 96
G1_AFFINE_WORDS, G2_AFFINE_WORDS = 25, 49

EXPORTS = [
    "g16_version", "g16_device_count", "g16_ctx_create", "g16_ctx_destroy", "g16_last_error",
    "g16_ctx_set_stream", "g16_ctx_synchronize", "g16_ctx_set_window_bits", "g16_ctx_set_h2d_pipeline_min", "g16_ctx_set_item_max",
    "g16_g1_bases_upload", "g16_g2_bases_upload", "g16_g1_bases_from_device", "g16_g2_bases_from_device",
    "g16_bases_free", "g16_bases_len", "g16_bases_precompute",
    "g16_g1_msm", "g16_g2_msm", "g16_g1_msm_oneshot", "g16_g2_msm_oneshot",
    "g16_g1_msm_device", "g16_g2_msm_device", "g16_g1_msm_async", "g16_g2_msm_async",
    "g16_g1_combine_partials_device", "g16_g2_combine_partials_device",
    "g16_g1_fixed_base_mul", "g16_g2_fixed_base_mul",
    "g16_g1_fixed_base_mul_device", "g16_g2_fixed_base_mul_device",
    "g16_pk_upload", "g16_pk_free", "g16_prove", "g16_pk_precompute", "g16_pk_precompute_bits", "g16_quotient_h", "g16_quotient_h_device",
    "g16_r1cs_upload", "g16_r1cs_free", "g16_r1cs_domain_size", "g16_r1cs_domain_evals", "g16_r1cs_eval_at",
    "g16_setup_crs", "g16_prove_r1cs",
    "g16_g1_serialize", "g16_g2_serialize", "g16_g1_deserialize", "g16_g2_deserialize",
    "g16_proof_serialize", "g16_proof_deserialize",
]


class MSMError(RuntimeError):
    """Counterpart of GrothError::MSMError(String) (crates/groth16-core/src/lib.rs:74-76)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"MSM computation error: {msg} (code {code})")
        self.code = code


class _PkHost(ctypes.Structure):
    _fields_ = [
        ("alpha_g1", ctypes.c_void_p), ("beta_g1", ctypes.c_void_p), ("delta_g1", ctypes.c_void_p),
        ("beta_g2", ctypes.c_void_p), ("delta_g2", ctypes.c_void_p),
        ("a_g1", ctypes.c_void_p), ("a_g1_inf", ctypes.c_void_p), ("a_len", ctypes.c_size_t),
        ("b_g1", ctypes.c_void_p), ("b_g1_inf", ctypes.c_void_p), ("b1_len", ctypes.c_size_t),
        ("b_g2", ctypes.c_void_p), ("b_g2_inf", ctypes.c_void_p), ("b2_len", ctypes.c_size_t),
        ("ic_g1", ctypes.c_void_p), ("ic_g1_inf", ctypes.c_void_p), ("ic_len", ctypes.c_size_t),
        ("h_g1", ctypes.c_void_p), ("h_g1_inf", ctypes.c_void_p), ("h_len", ctypes.c_size_t),
        ("num_public", ctypes.c_size_t),
    ]


class _Csr(ctypes.Structure):
    _fields_ = [("row_ptr", ctypes.c_void_p), ("col", ctypes.c_void_p), ("val", ctypes.c_void_p)]


class _CrsHost(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "alpha_g1", "beta_g1", "delta_g1", "beta_g2", "gamma_g2", "delta_g2",
        "a_g1", "a_g1_inf", "b_g1", "b_g1_inf", "b_g2", "b_g2_inf", "ic_g1", "ic_g1_inf",
        "vk_ic_g1", "vk_ic_g1_inf", "h_g1", "h_g1_inf")]


_libs = {}


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    path = path or DEFAULT_LIB
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise MSMError(G16_ERR_NO_DEVICE, f"{path} not found: build it with `python zero-knowledge-proofs_b200/build.py` "
                                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.g16_version.restype = ctypes.c_char_p
    lib.g16_last_error.restype = ctypes.c_char_p
    lib.g16_last_error.argtypes = [vp]
    lib.g16_device_count.restype = ci
    lib.g16_ctx_create.argtypes = [vp, ci, ctypes.POINTER(vp)]
    lib.g16_ctx_destroy.argtypes = [vp]
    lib.g16_ctx_destroy.restype = None
    lib.g16_ctx_set_stream.argtypes = [vp, vp]
    lib.g16_ctx_synchronize.argtypes = [vp]
    lib.g16_ctx_set_window_bits.argtypes = [vp, ctypes.c_uint]
    lib.g16_ctx_set_h2d_pipeline_min.argtypes = [vp, sz]
    lib.g16_ctx_set_item_max.argtypes = [vp, ctypes.c_uint]
    for g in ("g1", "g2"):
        getattr(lib, f"g16_{g}_bases_upload").argtypes = [vp, vp, vp, sz, ctypes.POINTER(vp)]
        getattr(lib, f"g16_{g}_bases_from_device").argtypes = [vp, vp, sz, ctypes.POINTER(vp)]
        getattr(lib, f"g16_{g}_msm").argtypes = [vp, vp, vp, sz, vp, vp]
        getattr(lib, f"g16_{g}_msm_oneshot").argtypes = [vp, vp, vp, vp, sz, vp, vp]
        getattr(lib, f"g16_{g}_msm_device").argtypes = [vp, vp, vp, sz, vp, vp]
        getattr(lib, f"g16_{g}_msm_async").argtypes = [vp, vp, vp, sz, vp, vp]
        getattr(lib, f"g16_{g}_combine_partials_device").argtypes = [vp, vp, sz, vp]
        getattr(lib, f"g16_{g}_fixed_base_mul").argtypes = [vp, vp, vp, sz, vp, vp]
        getattr(lib, f"g16_{g}_fixed_base_mul_device").argtypes = [vp, vp, vp, sz, vp]
    lib.g16_bases_precompute.argtypes = [vp, vp, ctypes.c_uint, sz, ctypes.POINTER(ctypes.c_uint)]
    lib.g16_bases_free.argtypes = [vp]
    lib.g16_bases_free.restype = None
    lib.g16_bases_len.argtypes = [vp]
    lib.g16_bases_len.restype = sz
    lib.g16_pk_upload.argtypes = [vp, ctypes.POINTER(_PkHost), ctypes.POINTER(vp)]
    lib.g16_pk_precompute.argtypes = [vp, vp]
    lib.g16_pk_precompute_bits.argtypes = [vp, vp, ctypes.c_uint]
    lib.g16_pk_free.argtypes = [vp]
    lib.g16_pk_free.restype = None
    lib.g16_prove.argtypes = [vp, vp, vp, sz, vp, sz, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.g16_quotient_h.argtypes = [vp, vp, vp, vp, sz, vp]
    lib.g16_quotient_h_device.argtypes = [vp, vp, sz, vp, vp]
    lib.g16_r1cs_upload.argtypes = [vp, sz, sz, ctypes.POINTER(_Csr), ctypes.POINTER(_Csr), ctypes.POINTER(_Csr),
                                    ctypes.POINTER(vp)]
    lib.g16_r1cs_free.argtypes = [vp]
    lib.g16_r1cs_free.restype = None
    lib.g16_r1cs_domain_size.argtypes = [vp]
    lib.g16_r1cs_domain_size.restype = sz
    lib.g16_r1cs_domain_evals.argtypes = [vp, vp, vp, sz, vp, vp, vp]
    lib.g16_r1cs_eval_at.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.g16_setup_crs.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz, ctypes.POINTER(_CrsHost), ctypes.POINTER(vp)]
    lib.g16_prove_r1cs.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp, vp, vp, vp, vp, vp]
    u8 = ctypes.c_uint8
    for g in ("g1", "g2"):
        getattr(lib, f"g16_{g}_serialize").argtypes = [vp, vp, vp, sz, ci, vp]
        getattr(lib, f"g16_{g}_deserialize").argtypes = [vp, vp, sz, ci, ci, vp, vp, vp]
    lib.g16_proof_serialize.argtypes = [vp, vp, u8, vp, u8, vp, u8, ci, vp]
    lib.g16_proof_deserialize.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp, vp]
    _libs[path] = lib
    return lib


def _u64(a, width=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if width is not None:
        a = a.reshape(-1, width)
    return a


def _ptr(a) -> Optional[int]:
    return None if a is None else a.ctypes.data


class Bases:
    """Device-resident base points (g16_bases)."""

    def __init__(self, ctx: "Context", handle: int, group: int, keepalive=None):
        self.ctx, self.handle, self.group, self._keep = ctx, handle, group, keepalive
        ctx._children.add(self)

    def __len__(self):
        return int(self.ctx.lib.g16_bases_len(self.handle))

    def precompute(self, window_bits: int = 0, budget_bytes: int = 0) -> int:
        """One-time table of multiples 2^(c w) P_i (see g16_bases_precompute); returns c."""
        used = ctypes.c_uint(0)
        self.ctx._check(self.ctx.lib.g16_bases_precompute(self.ctx.handle, self.handle, window_bits, budget_bytes,
                                                          ctypes.byref(used)))
        return int(used.value)

    def free(self):
        if self.handle:
            self.ctx.lib.g16_bases_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ProvingKeyDevice:
    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.handle = ctx, handle
        ctx._children.add(self)

    def free(self):
        if self.handle:
            self.ctx.lib.g16_pk_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class R1CSDevice:
    """Device-resident sparse constraint system (g16_r1cs): the three matrices of `R1CS<F>`
    (crates/groth16-r1cs) in CSR form, by constraint and by variable."""

    def __init__(self, ctx: "Context", handle: int, num_constraints: int, num_variables: int):
        self.ctx, self.handle = ctx, handle
        ctx._children.add(self)
        self.num_constraints, self.num_variables = num_constraints, num_variables

    @property
    def domain_size(self) -> int:
        return int(self.ctx.lib.g16_r1cs_domain_size(self.handle))

    def free(self):
        if self.handle:
            self.ctx.lib.g16_r1cs_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    def __init__(self, devices: Optional[Sequence[int]] = None, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        self._children = weakref.WeakSet()   # handles that point into this context: freed before it is destroyed
        h = ctypes.c_void_p()
        if devices:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self.lib.g16_ctx_create(arr, len(devices), ctypes.byref(h))
        else:
            rc = self.lib.g16_ctx_create(None, 0, ctypes.byref(h))
        if rc != G16_OK:
            raise MSMError(rc, self.lib.g16_last_error(None).decode())
        self.handle = h.value

    # ---- plumbing
    def _check(self, rc: int):
        if rc != G16_OK:
            raise MSMError(rc, self.lib.g16_last_error(self.handle).decode())

    def close(self):
        if getattr(self, "handle", None):
            for child in list(self._children):
                child.free()
            self.lib.g16_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.g16_ctx_set_stream(self.handle, cuda_stream))

    def synchronize(self):
        self._check(self.lib.g16_ctx_synchronize(self.handle))

    def set_window_bits(self, c: int):
        self._check(self.lib.g16_ctx_set_window_bits(self.handle, c))

    def set_h2d_pipeline_min(self, min_scalars: int):
        """Host-scalar MSMs of at least this many scalars per device pipeline their H2D copy (0 = default 2^19)."""
        self._check(self.lib.g16_ctx_set_h2d_pipeline_min(self.handle, min_scalars))

    def set_item_max(self, item_max: int):
        """Tuning: longest serial run of additions per thread in the bucket accumulation (0 = chosen per call)."""
        self._check(self.lib.g16_ctx_set_item_max(self.handle, item_max))

    # ---- bases
    def _upload(self, g: str, xy, inf) -> Bases:
        width = G1_WORDS64 if g == "g1" else G2_WORDS64
        xy = _u64(xy, width)
        n = xy.shape[0]
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
            if inf.shape[0] != n:
                raise MSMError(G16_ERR_LENGTH, "infinity flags length mismatch")
        h = ctypes.c_void_p()
        self._check(getattr(self.lib, f"g16_{g}_bases_upload")(self.handle, _ptr(xy), _ptr(inf), n, ctypes.byref(h)))
        return Bases(self, h.value, 1 if g == "g1" else 2)

    def g1_bases_upload(self, xy, inf=None) -> Bases:
        return self._upload("g1", xy, inf)

    def g2_bases_upload(self, xy, inf=None) -> Bases:
        return self._upload("g2", xy, inf)

    def bases_from_device(self, g: str, dev_ptr: int, n: int, keepalive=None) -> Bases:
        h = ctypes.c_void_p()
        self._check(getattr(self.lib, f"g16_{g}_bases_from_device")(self.handle, dev_ptr, n, ctypes.byref(h)))
        return Bases(self, h.value, 1 if g == "g1" else 2, keepalive)

    # ---- MSM with resident bases
    def _msm(self, g: str, bases: Bases, scalars) -> Tuple[np.ndarray, int]:
        width = G1_WORDS64 if g == "g1" else G2_WORDS64
        scalars = _u64(scalars, 4)
        out = np.zeros(width, dtype=np.uint64)
        inf = np.zeros(1, dtype=np.uint8)
        self._check(getattr(self.lib, f"g16_{g}_msm")(self.handle, bases.handle, _ptr(scalars), scalars.shape[0],
                                                     _ptr(out), _ptr(inf)))
        return out, int(inf[0])

    def g1_msm(self, bases: Bases, scalars):
        return self._msm("g1", bases, scalars)

    def g2_msm(self, bases: Bases, scalars):
        return self._msm("g2", bases, scalars)

    # ---- the reference seam: fresh (scalar, point) lists per call
    def _oneshot(self, g: str, xy, inf, scalars):
        width = G1_WORDS64 if g == "g1" else G2_WORDS64
        xy = _u64(xy, width)
        scalars = _u64(scalars, 4)
        if xy.shape[0] != scalars.shape[0]:
            # ark: `msm` returns Err(min_len) -> GrothError::MSMError (lib.rs:283,297)
            raise MSMError(G16_ERR_LENGTH, f"G{1 if g == 'g1' else 2} MSM failed: {min(xy.shape[0], scalars.shape[0])}")
        if inf is not None:
            inf = np.ascontiguousarray(inf, dtype=np.uint8)
        out = np.zeros(width, dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        self._check(getattr(self.lib, f"g16_{g}_msm_oneshot")(self.handle, _ptr(xy), _ptr(inf), _ptr(scalars),
                                                             xy.shape[0], _ptr(out), _ptr(oinf)))
        return out, int(oinf[0])

    def multi_scalar_mult_g1(self, scalars, points_xy, points_inf=None):
        """Prover::multi_scalar_mult_g1: sum s_i P_i as an affine point; empty input -> identity."""
        return self._oneshot("g1", points_xy, points_inf, scalars)

    def multi_scalar_mult_g2(self, scalars, points_xy, points_inf=None):
        return self._oneshot("g2", points_xy, points_inf, scalars)

    # ---- device-resident variants (pointers are plain ints, e.g. torch.Tensor.data_ptr())
    def msm_device(self, g: str, bases: Bases, dev_scalars: int, n: int, dev_out_affine: int = 0, dev_out_partial: int = 0):
        self._check(getattr(self.lib, f"g16_{g}_msm_device")(self.handle, bases.handle, dev_scalars, n,
                                                            dev_out_affine or None, dev_out_partial or None))

    def msm_async(self, g: str, bases: Bases, host_scalars_ptr: int, n: int, dev_out_affine: int = 0,
                  dev_out_partial: int = 0):
        """HOST scalars (raw pointer, e.g. a pinned buffer) -> device result, asynchronous on the ctx stream."""
        self._check(getattr(self.lib, f"g16_{g}_msm_async")(self.handle, bases.handle, host_scalars_ptr, n,
                                                           dev_out_affine or None, dev_out_partial or None))

    def combine_partials_device(self, g: str, dev_partials: int, k: int, dev_out_affine: int):
        self._check(getattr(self.lib, f"g16_{g}_combine_partials_device")(self.handle, dev_partials, k, dev_out_affine))

    # ---- fixed base
    def _fixed(self, g: str, base_xy, scalars, out=None):
        width = G1_WORDS64 if g == "g1" else G2_WORDS64
        base_xy = _u64(base_xy).reshape(width)
        scalars = _u64(scalars, 4)
        n = scalars.shape[0]
        if out is None:
            out, inf = np.zeros((n, width), dtype=np.uint64), np.zeros(n, dtype=np.uint8)
        else:   # caller-owned result buffers (e.g. pinned host memory), the C ABI's own convention
            out, inf = out
            assert out.dtype == np.uint64 and out.shape == (n, width) and out.flags.c_contiguous
            assert inf.dtype == np.uint8 and inf.shape == (n,)
        self._check(getattr(self.lib, f"g16_{g}_fixed_base_mul")(self.handle, _ptr(base_xy), _ptr(scalars), n,
                                                                _ptr(out), _ptr(inf)))
        return out, inf

    def fixed_base_mul_g1(self, base_xy, scalars, out=None):
        """[(base * s).into_affine() for s in scalars] (crates/groth16-setup/src/lib.rs:185-241).
        out = (xy[n, 12] uint64, inf[n] uint8) to receive the result in caller-owned memory."""
        return self._fixed("g1", base_xy, scalars, out)

    def fixed_base_mul_g2(self, base_xy, scalars, out=None):
        return self._fixed("g2", base_xy, scalars, out)

    def fixed_base_mul_device(self, g: str, base_xy, dev_scalars: int, n: int, dev_out: int):
        width = G1_WORDS64 if g == "g1" else G2_WORDS64
        base_xy = _u64(base_xy).reshape(width)
        self._check(getattr(self.lib, f"g16_{g}_fixed_base_mul_device")(self.handle, _ptr(base_xy), dev_scalars, n, dev_out))

    # ---- quotient polynomial
    def quotient_h(self, a_evals, b_evals, c_evals, out=None) -> np.ndarray:
        """H = (A*B - C) / (x^n - 1) from domain evaluations (QAP::compute_quotient_polynomial); n x 4 u64.
        out: caller-owned n x 4 uint64 result buffer (e.g. pinned host memory: pinned buffers are copied at PCIe rate)."""
        a, b, c = (_u64(x, 4) for x in (a_evals, b_evals, c_evals))
        n = a.shape[0]
        if b.shape[0] != n or c.shape[0] != n:
            raise MSMError(G16_ERR_LENGTH, "evaluation vectors differ in length")
        h = np.zeros((n, 4), dtype=np.uint64) if out is None else out
        assert h.dtype == np.uint64 and h.shape == (n, 4) and h.flags.c_contiguous
        self._check(self.lib.g16_quotient_h(self.handle, _ptr(a), _ptr(b), _ptr(c), n, _ptr(h)))
        return h

    def quotient_h_device(self, dev_abc: int, n: int, dev_h: int, dev_bad_rows: int):
        """Device pointers (e.g. torch.Tensor.data_ptr()): abc = A, B, C evaluations back to back (overwritten)."""
        self._check(self.lib.g16_quotient_h_device(self.handle, dev_abc, n, dev_h, dev_bad_rows))

    # ---- prove
    def pk_upload(self, pk: dict) -> ProvingKeyDevice:
        """pk: dict with alpha_g1, beta_g1, delta_g1 (12 u64), beta_g2, delta_g2 (24 u64), a_g1, b_g1, ic_g1, h_g1
        (n x 12), b_g2 (n x 24), optional *_inf byte arrays, num_public."""
        keep = []

        def arr(name, width):
            a = _u64(pk[name], width)
            keep.append(a)
            return a

        def flags(name, n):
            f = pk.get(name + "_inf")
            if f is None:
                return None
            f = np.ascontiguousarray(f, dtype=np.uint8)
            assert f.shape[0] == n
            keep.append(f)
            return f

        s = _PkHost()
        for name, width in (("alpha_g1", 12), ("beta_g1", 12), ("delta_g1", 12), ("beta_g2", 24), ("delta_g2", 24)):
            setattr(s, name, _ptr(arr(name, width)))
        for name, width, ln in (("a_g1", 12, "a_len"), ("b_g1", 12, "b1_len"), ("b_g2", 24, "b2_len"),
                                ("ic_g1", 12, "ic_len"), ("h_g1", 12, "h_len")):
            a = arr(name, width)
            setattr(s, name, _ptr(a) if a.shape[0] else None)
            setattr(s, name + "_inf", _ptr(flags(name, a.shape[0])))
            setattr(s, ln, a.shape[0])
        s.num_public = int(pk["num_public"])
        h = ctypes.c_void_p()
        self._check(self.lib.g16_pk_upload(self.handle, ctypes.byref(s), ctypes.byref(h)))
        return ProvingKeyDevice(self, h.value)

    def pk_precompute(self, pk: ProvingKeyDevice, scalar_bits: int = 0):
        """One-time tables of multiples for the five resident arrays (g16_pk_precompute / g16_pk_precompute_bits).
        scalar_bits: the assignment and H coefficients are promised to be below 2^scalar_bits (0 = full width; 64 is
        what the reference's truncation yields) -- a tuning hint only, any scalar stays valid."""
        self._check(self.lib.g16_pk_precompute_bits(self.handle, pk.handle, scalar_bits))

    def prove(self, pk: ProvingKeyDevice, assignment_fr, h_coeffs, r, s):
        """Group part of Prover::prove.  Returns ((a_xy, a_inf), (b_xy, b_inf), (c_xy, c_inf))."""
        assignment_fr = _u64(assignment_fr, 4)
        h_coeffs = _u64(h_coeffs if h_coeffs is not None else np.zeros((0, 4)), 4)
        r = _u64(r).reshape(4)
        s = _u64(s).reshape(4)
        a = np.zeros(12, dtype=np.uint64); b = np.zeros(24, dtype=np.uint64); c = np.zeros(12, dtype=np.uint64)
        fl = np.zeros(3, dtype=np.uint8)
        self._check(self.lib.g16_prove(self.handle, pk.handle, _ptr(assignment_fr), assignment_fr.shape[0],
                                       _ptr(h_coeffs) if h_coeffs.shape[0] else None, h_coeffs.shape[0],
                                       _ptr(r), _ptr(s), _ptr(a), fl.ctypes.data, _ptr(b), fl.ctypes.data + 1,
                                       _ptr(c), fl.ctypes.data + 2))
        return (a, int(fl[0])), (b, int(fl[1])), (c, int(fl[2]))

    # ---- wire format: ark CanonicalSerialize / CanonicalDeserialize (Zcash encoding of ark-bls12-381)
    def serialize_points(self, group: str, points_xy, points_inf=None, compressed: bool = True) -> bytes:
        """Concatenated encodings of a `Vec<G1Affine>` / `Vec<G2Affine>` body (no length prefix)."""
        width = G1_WORDS64 if group == "g1" else G2_WORDS64
        xy = _u64(points_xy, width)
        n = xy.shape[0]
        inf = None if points_inf is None else np.ascontiguousarray(points_inf, dtype=np.uint8)
        per = (48 if compressed else 96) * (1 if group == "g1" else 2)
        out = np.zeros(n * per, dtype=np.uint8)
        f = getattr(self.lib, f"g16_{group}_serialize")
        self._check(f(self.handle, _ptr(xy) if n else None, _ptr(inf), n, int(compressed), _ptr(out) if n else None))
        return out.tobytes()

    def deserialize_points(self, group: str, data: bytes, compressed: bool = True, validate: bool = True,
                           return_status: bool = False):
        """Inverse of serialize_points.  Raises MSMError ("InvalidData ..." / "UnexpectedFlags ...") like ark's
        SerializationError unless return_status, in which case (xy, inf, status) comes back for every element."""
        width = G1_WORDS64 if group == "g1" else G2_WORDS64
        per = (48 if compressed else 96) * (1 if group == "g1" else 2)
        if len(data) % per:
            raise MSMError(G16_ERR_LENGTH, "InvalidData: truncated input")
        n = len(data) // per
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        xy = np.zeros((n, width), dtype=np.uint64)
        inf = np.zeros(n, dtype=np.uint8)
        status = np.zeros(n, dtype=np.uint8)
        f = getattr(self.lib, f"g16_{group}_deserialize")
        rc = f(self.handle, _ptr(buf) if n else None, n, int(compressed), int(validate), _ptr(xy) if n else None,
               _ptr(inf), _ptr(status))
        if return_status and rc in (G16_OK, G16_ERR_INVALID):
            return xy, inf, status
        self._check(rc)
        return xy, inf

    def proof_serialize(self, a, b, c, compressed: bool = True) -> bytes:
        """`Proof::serialize_compressed` / `serialize_uncompressed` (crates/groth16-core/src/lib.rs:27-36): a, b, c as
        (xy, inf) pairs, the shape Context.prove returns.  192 / 384 bytes."""
        (a_xy, a_inf), (b_xy, b_inf), (c_xy, c_inf) = a, b, c
        a_xy, b_xy, c_xy = _u64(a_xy).reshape(12), _u64(b_xy).reshape(24), _u64(c_xy).reshape(12)
        out = np.zeros(192 if compressed else 384, dtype=np.uint8)
        self._check(self.lib.g16_proof_serialize(self.handle, _ptr(a_xy), int(a_inf), _ptr(b_xy), int(b_inf), _ptr(c_xy),
                                                 int(c_inf), int(compressed), _ptr(out)))
        return out.tobytes()

    def proof_deserialize(self, data: bytes, compressed: bool = True, validate: bool = True):
        if len(data) != (192 if compressed else 384):
            raise MSMError(G16_ERR_LENGTH, "InvalidData: a proof is 192 bytes compressed, 384 uncompressed")
        buf = np.frombuffer(data, dtype=np.uint8).copy()
        a = np.zeros(12, dtype=np.uint64); b = np.zeros(24, dtype=np.uint64); c = np.zeros(12, dtype=np.uint64)
        fl = np.zeros(3, dtype=np.uint8)
        self._check(self.lib.g16_proof_deserialize(self.handle, _ptr(buf), int(compressed), int(validate), _ptr(a),
                                                   fl.ctypes.data, _ptr(b), fl.ctypes.data + 1, _ptr(c), fl.ctypes.data + 2))
        return (a, int(fl[0])), (b, int(fl[1])), (c, int(fl[2]))

    # ---- sparse R1CS: setup and prove for real circuits
    def r1cs_upload(self, num_constraints: int, num_variables: int, a, b, c) -> R1CSDevice:
        """a, b, c: (row_ptr[num_constraints + 1] uint32, col[nnz] uint32, val[nnz x 4] uint64 Montgomery Fr)."""
        keep, structs = [], []
        for row_ptr, col, val in (a, b, c):
            row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
            col = np.ascontiguousarray(col, dtype=np.uint32)
            val = _u64(val, 4)
            if row_ptr.shape[0] != num_constraints + 1 or col.shape[0] != val.shape[0] or \
                    (num_constraints and int(row_ptr[-1]) != col.shape[0]):
                raise MSMError(G16_ERR_LENGTH, "CSR arrays are inconsistent")
            keep += [row_ptr, col, val]
            st = _Csr()
            st.row_ptr, st.col, st.val = _ptr(row_ptr), _ptr(col) if col.shape[0] else None, _ptr(val) if col.shape[0] else None
            structs.append(st)
        h = ctypes.c_void_p()
        self._check(self.lib.g16_r1cs_upload(self.handle, num_constraints, num_variables, ctypes.byref(structs[0]),
                                             ctypes.byref(structs[1]), ctypes.byref(structs[2]), ctypes.byref(h)))
        return R1CSDevice(self, h.value, num_constraints, num_variables)

    def r1cs_domain_evals(self, r1cs: R1CSDevice, assignment):
        """(<A-row i, w>, <B-row i, w>, <C-row i, w>) on the domain; three n x 4 uint64 arrays."""
        w = _u64(assignment, 4)
        n = r1cs.domain_size
        out = [np.zeros((n, 4), dtype=np.uint64) for _ in range(3)]
        self._check(self.lib.g16_r1cs_domain_evals(self.handle, r1cs.handle, _ptr(w), w.shape[0], *[_ptr(o) for o in out]))
        return out

    def r1cs_eval_at(self, r1cs: R1CSDevice, s):
        """(A_j(s), B_j(s), C_j(s)) for every variable j; three num_variables x 4 uint64 arrays."""
        s = _u64(s).reshape(4)
        out = [np.zeros((r1cs.num_variables, 4), dtype=np.uint64) for _ in range(3)]
        self._check(self.lib.g16_r1cs_eval_at(self.handle, r1cs.handle, _ptr(s), *[_ptr(o) for o in out]))
        return out

    def setup_crs(self, r1cs: R1CSDevice, params: dict, num_public: int, want_host: bool = True, want_device_pk: bool = False):
        """CRS::generate_from_qap.  params: alpha, beta, gamma, delta, s as 4 x uint64 Montgomery limbs.
        Returns (crs dict of host arrays or None, ProvingKeyDevice or None)."""
        p = {k: _u64(params[k]).reshape(4) for k in ("alpha", "beta", "gamma", "delta", "s")}
        nv, n = r1cs.num_variables, r1cs.domain_size
        crs, st = None, None
        if want_host:
            n_ic = max(0, nv - num_public - 1)
            crs = {"num_public": num_public}
            st = _CrsHost()
            for name, w in (("alpha_g1", 12), ("beta_g1", 12), ("delta_g1", 12), ("beta_g2", 24), ("gamma_g2", 24), ("delta_g2", 24)):
                crs[name] = np.zeros(w, dtype=np.uint64)
                setattr(st, name, _ptr(crs[name]))
            for name, w, ln in (("a_g1", 12, nv), ("b_g1", 12, nv), ("b_g2", 24, nv), ("ic_g1", 12, n_ic),
                                ("vk_ic_g1", 12, min(num_public + 1, nv)), ("h_g1", 12, n)):
                crs[name] = np.zeros((ln, w), dtype=np.uint64)
                crs[name + "_inf"] = np.zeros(ln, dtype=np.uint8)
                setattr(st, name, _ptr(crs[name]) if ln else None)
                setattr(st, name + "_inf", _ptr(crs[name + "_inf"]) if ln else None)
        h = ctypes.c_void_p()
        self._check(self.lib.g16_setup_crs(self.handle, r1cs.handle, _ptr(p["alpha"]), _ptr(p["beta"]), _ptr(p["gamma"]),
                                           _ptr(p["delta"]), _ptr(p["s"]), num_public,
                                           ctypes.byref(st) if st is not None else None,
                                           ctypes.byref(h) if want_device_pk else None))
        return crs, (ProvingKeyDevice(self, h.value) if want_device_pk else None)

    def prove_r1cs(self, pk: ProvingKeyDevice, r1cs: R1CSDevice, assignment, r, s):
        """Prover::prove from the un-truncated witness.  Returns ((a_xy, a_inf), (b_xy, b_inf), (c_xy, c_inf))."""
        w = _u64(assignment, 4)
        r = _u64(r).reshape(4)
        s = _u64(s).reshape(4)
        a = np.zeros(12, dtype=np.uint64); b = np.zeros(24, dtype=np.uint64); c = np.zeros(12, dtype=np.uint64)
        fl = np.zeros(3, dtype=np.uint8)
        self._check(self.lib.g16_prove_r1cs(self.handle, pk.handle, r1cs.handle, _ptr(w), w.shape[0], _ptr(r), _ptr(s),
                                            _ptr(a), fl.ctypes.data, _ptr(b), fl.ctypes.data + 1, _ptr(c), fl.ctypes.data + 2))
        return (a, int(fl[0])), (b, int(fl[1])), (c, int(fl[2]))
