"""Build libg16cuda.so (sm_100a) in-tree: one nvcc -c per translation unit, in parallel, then link.

    python zero-knowledge-proofs_b200/build.py [--force] [--jobs N] [--verbose]

The library lands in zero-knowledge-proofs_b200/lib/libg16cuda.so (git-ignored; it travels to the
GPU box with the gpurun snapshot).  Objects are cached under lib/obj and rebuilt when any header or
the unit itself is newer.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libg16cuda.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def source_hash() -> str:
    """sha256 over csrc/*.cu, csrc/*.cuh and include/*.h (names + contents, sorted): baked into g16_version()."""
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                   glob.glob(os.path.join(INCLUDE, "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _newest_header() -> float:
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(unit: str, force: bool, verbose: bool):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJDIR, unit.replace(".cu", ".o"))
    log = obj + ".log"
    digest = source_hash()
    stamp = obj + ".hash"
    if unit == "version.cu":
        # cheap and always exact: rebuilt whenever any source changed since the last build
        fresh = os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest
        if not force and fresh:
            return unit, 0.0, "cached"
    elif not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), _newest_header()):
        return unit, 0.0, "cached"
    t0 = time.time()
    cmd = [NVCC] + ARCH + COMMON + ["-c", src, "-o", obj]
    if unit == "version.cu":
        cmd += [f'-DG16_SOURCE_HASH="{digest}"']
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {unit}:\n{res.stdout[-4000:]}")
    if verbose:
        print(res.stdout)
    if unit == "version.cu":
        with open(stamp, "w") as f:
            f.write(digest)
    return unit, time.time() - t0, "built"


def build(force: bool = False, jobs: int | None = None, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    units = sorted(os.path.basename(p) for p in glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    results = []
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for r in ex.map(lambda u: _compile(u, force, verbose), units):
            results.append(r)
    rebuilt = any(r[2] == "built" for r in results)
    if rebuilt or not os.path.exists(LIB):
        objs = [os.path.join(OBJDIR, u.replace(".cu", ".o")) for u in units]
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        subprocess.run(cmd, check=True)
    for unit, dt, what in results:
        if what == "built":
            print(f"[build] {unit}: {dt:.1f}s")
    build_tools(force)
    return LIB


def build_tools(force: bool = False) -> str:
    """lib/imad_peak: the integer-pipe roofline microbenchmark bench.py runs before the timed region;
    lib/dfma_peak, lib/dpf_mul_bench: the FP64-pipe companions quoted in DESIGN.md."""
    first = None
    for name in ("imad_peak", "dfma_peak", "dpf_mul_bench"):
        src = os.path.join(HERE, "tools", name + ".cu")
        out = os.path.join(LIBDIR, name)
        if force or not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), _newest_header()):
            subprocess.run([NVCC] + ARCH + ["-O3", "-std=c++17", "-lineinfo", "-o", out, src], check=True)
        first = first or out
    return first


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    t = time.time()
    print(build(a.force, a.jobs, a.verbose), f"({time.time() - t:.1f}s)")
