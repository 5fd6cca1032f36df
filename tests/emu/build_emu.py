"""TEST INFRASTRUCTURE ONLY -- host emulation build of the kernel bodies.

This container has no GPU.  To debug the pipeline logic before spending GPU time, the very
same sources (zero-knowledge-proofs_b200/csrc/*.cu) are compiled here with g++ and -DG16_EMU:
kernel bodies run in a host loop and the PTX carry-chain primitives are replaced by their C
emulation (g16_defs.cuh).  The resulting tests/emu/_build/libg16emu.so is loaded ONLY by
tests/test_emu_*.py; the product library (lib/libg16cuda.so) contains no such path and the
Python binding never looks for this file.
"""
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "zero-knowledge-proofs_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libg16emu.so")


def build(force: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(os.path.dirname(OUT), os.path.basename(s) + ".o")
        objs.append(o)
        extra = os.environ.get("G16_EMU_EXTRA_FLAGS", "").split()   # A/B of compile-time variants (e.g. -DG16_FQ2_DUAL=1)
        procs.append(subprocess.Popen(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-DG16_EMU=1"] + extra + ["-x", "c++", "-c", s, "-o", o]))
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emu build failed")
    subprocess.run(["/usr/bin/g++", "-shared", "-o", OUT] + objs, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
