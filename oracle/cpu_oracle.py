"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  The library is the C restatement of the reference's CPU path (ark-ec 0.4.2
msm_bigint_wnaf etc., see oracle/cpu_msm.c); it shares the packed limb layout of
include/g16_cuda.h so that parity tests pass identical buffers to both sides.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ora_max_threads.restype = ctypes.c_int
        _lib.ora_msm_window.restype = ctypes.c_int
        _lib.ora_msm_window.argtypes = [ctypes.c_size_t]
    return _lib


def _p64(a):
    return a.ctypes.data_as(_u64p)


def _p8(a):
    return None if a is None else a.ctypes.data_as(_u8p)


def max_threads() -> int:
    """Host threads the CPU arm may use: the CPUs this process is allowed to run on.  Deliberately NOT
    omp_get_max_threads(): launchers such as torch.distributed.run export OMP_NUM_THREADS=1, which made
    the baseline depend on how bench.py was started.  Every caller passes the count
    explicitly (num_threads clause), so the environment variable has no effect."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def _msm(name, width, bases, inf, scalars, threads):
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, width)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = bases.shape[0]
    assert scalars.shape[0] == n
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8)
        assert inf.shape[0] == n
    out = np.zeros(width, dtype=np.uint64)
    out_inf = np.zeros(1, dtype=np.uint8)
    fn = getattr(lib(), name)
    if name.endswith("naive"):
        rc = fn(_p64(bases), _p8(inf), _p64(scalars), ctypes.c_size_t(n), _p64(out), _p8(out_inf))
    else:
        rc = fn(_p64(bases), _p8(inf), _p64(scalars), ctypes.c_size_t(n), ctypes.c_int(threads),
                _p64(out), _p8(out_inf))
    if rc != 0:
        raise RuntimeError(f"{name} failed rc={rc}")
    return out, int(out_inf[0])


def g1_msm(bases, inf, scalars_mont, threads=1):
    """ark msm_bigint_wnaf + into_affine.  bases: n x 12 u64, scalars: n x 4 u64 (Montgomery)."""
    return _msm("ora_g1_msm", 12, bases, inf, scalars_mont, threads)


def g2_msm(bases, inf, scalars_mont, threads=1):
    return _msm("ora_g2_msm", 24, bases, inf, scalars_mont, threads)


def g1_msm_naive(bases, inf, scalars_mont):
    return _msm("ora_g1_msm_naive", 12, bases, inf, scalars_mont, 1)


def g2_msm_naive(bases, inf, scalars_mont):
    return _msm("ora_g2_msm_naive", 24, bases, inf, scalars_mont, 1)


def _fixed(name, width, base_xy, scalars, threads):
    base_xy = np.ascontiguousarray(base_xy, dtype=np.uint64).reshape(width)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    n = scalars.shape[0]
    out = np.zeros((n, width), dtype=np.uint64)
    out_inf = np.zeros(n, dtype=np.uint8)
    rc = getattr(lib(), name)(_p64(base_xy), _p64(scalars), ctypes.c_size_t(n), ctypes.c_int(threads),
                              _p64(out), _p8(out_inf))
    if rc != 0:
        raise RuntimeError(f"{name} failed rc={rc}")
    return out, out_inf


def g1_fixed_base_mul(base_xy, scalars_mont, threads=1):
    """(base * s_i).into_affine() per element, crates/groth16-setup/src/lib.rs:185-241."""
    return _fixed("ora_g1_fixed_base_mul", 12, base_xy, scalars_mont, threads)


def g2_fixed_base_mul(base_xy, scalars_mont, threads=1):
    return _fixed("ora_g2_fixed_base_mul", 24, base_xy, scalars_mont, threads)


def _binop(name, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 6)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 6)
    r = np.zeros_like(a)
    getattr(lib(), name)(_p64(a), _p64(b), _p64(r), ctypes.c_size_t(a.shape[0]))
    return r


def fq_mul(a, b): return _binop("ora_fq_mul", a, b)
def fq_add(a, b): return _binop("ora_fq_add", a, b)
def fq_sub(a, b): return _binop("ora_fq_sub", a, b)


def fq_inv(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 6)
    r = np.zeros_like(a)
    lib().ora_fq_inv(_p64(a), _p64(r), ctypes.c_size_t(a.shape[0]))
    return r


def fr_from_mont(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    r = np.zeros_like(a)
    lib().ora_fr_from_mont(_p64(a), _p64(r), ctypes.c_size_t(a.shape[0]))
    return r


def fr_to_mont(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    r = np.zeros_like(a)
    lib().ora_fr_to_mont(_p64(a), _p64(r), ctypes.c_size_t(a.shape[0]))
    return r


def dot_mod_r(a_mont, b_mont, threads: int = 0):
    """sum a_i b_i mod r of two Montgomery Fr vectors; returns the sum in Montgomery form (4 x uint64)."""
    a = np.ascontiguousarray(a_mont, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b_mont, dtype=np.uint64).reshape(-1, 4)
    assert a.shape == b.shape
    out = np.zeros(4, dtype=np.uint64)
    lib().ora_dot_mod_r(_p64(a), _p64(b), ctypes.c_size_t(a.shape[0]), ctypes.c_int(threads or max_threads()), _p64(out))
    return out


def gen_scalars(seed: int, n: int, bits: int = 255):
    """n scalars uniform in [0, min(r, 2^bits)) from SplitMix64(seed), Montgomery form."""
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().ora_gen_scalars(ctypes.c_uint64(seed), ctypes.c_int(bits), ctypes.c_size_t(n), _p64(out))
    return out


def msm_window(n: int) -> int:
    return lib().ora_msm_window(n)


def make_digits(scalar_limbs, w: int):
    a = np.ascontiguousarray(scalar_limbs, dtype=np.uint64).reshape(4)
    cnt = (255 + w - 1) // w
    d = np.zeros(cnt, dtype=np.int64)
    lib().ora_make_digits(_p64(a), ctypes.c_int(w), d.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return d
