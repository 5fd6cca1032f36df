#!/usr/bin/env python
"""A/B variants of libg16cuda.so for kernel experiments (NOT shipped, NOT loaded by the product binding).

    python zero-knowledge-proofs_b200/tools/lab_build.py            # builds every variant below into lib/lab/<name>.so
    python zero-knowledge-proofs_b200/tools/bench_stages.py --group g2 --log-n 20 --lib zero-knowledge-proofs_b200/lib/lab/<name>.so

A variant recompiles the listed translation units with extra -D flags and links them with the standard objects of
lib/obj (build.py must have run).  The knobs are plain compile-time constants with the shipped values as defaults
(csrc/msm_kernels.cuh): the library itself has one code path and reads no environment variable."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
import build as b  # noqa: E402

LAB = os.path.join(b.LIBDIR, "lab")
ENGINE_UNITS = ("api.cu", "api_r1cs.cu", "api_wire.cu")   # translation units that include engine.cuh


def engine(*flags):
    return {u: list(flags) for u in ENGINE_UNITS}


VARIANTS = {
    # name: {unit: [flags]}   (shipped: accumulate 64 threads x 6 blocks (G1); item floor 16; reduction split 15 / 15)
    "g1_b128": {"k_acc_g1.cu": ["-DG16_ACC_BLOCK=128", "-DG16_ACC_MIN_BLOCKS_G1=3"]},
    "g1_b32_mb12": {"k_acc_g1.cu": ["-DG16_ACC_BLOCK=32", "-DG16_ACC_MIN_BLOCKS_G1=12"]},
    "g1_mb5": {"k_acc_g1.cu": ["-DG16_ACC_MIN_BLOCKS_G1=5"]},
    "g1_mb7": {"k_acc_g1.cu": ["-DG16_ACC_MIN_BLOCKS_G1=7"]},
    "g1_b128_mb3": {"k_acc_g1.cu": ["-DG16_ACC_BLOCK=128", "-DG16_ACC_MIN_BLOCKS_G1=3"]},
    "g2_mb2_b128": {"k_acc_g2.cu": ["-DG16_ACC_BLOCK=128", "-DG16_ACC_MIN_BLOCKS_G2=2"]},
    "g2_b32": {"k_acc_g2.cu": ["-DG16_ACC_BLOCK=32", "-DG16_ACC_MIN_BLOCKS_G2=8"]},
    "g2_b128": {"k_acc_g2.cu": ["-DG16_ACC_BLOCK=128"]},
    "g2_acc_fq2_calls": {"k_acc_g2.cu": ["-DG16_COLD_FQ2=1"]},
    "g2_acc_fq2_calls_mb6": {"k_acc_g2.cu": ["-DG16_COLD_FQ2=1", "-DG16_ACC_MIN_BLOCKS_G2=6"]},
    # G2 accumulate on lane pairs (csrc/pair_g2.cuh) instead of one thread per item; Karatsuba instead of the two-product Fq2 multiplication
    # y3 = r (q - x3) - y1 ppp as two multiplications instead of one two- / four-product multiplication
    "g1_acc_no_mul_diff": {"k_acc_g1.cu": ["-DG16_MUL_DIFF=0"]},
    "g2_acc_no_mul_diff": {"k_acc_g2.cu": ["-DG16_MUL_DIFF=0"]},
    "g2_pair": {"k_acc_g2.cu": ["-DG16_G2_ACC_THREAD=0"]},
    "g2_pair_mb4": {"k_acc_g2.cu": ["-DG16_G2_ACC_THREAD=0", "-DG16_PAIR_MIN_BLOCKS=4"]},
    "g2_acc_karatsuba": {"k_acc_g2.cu": ["-DG16_FQ2_DUAL=0"]},
    "g2_red_dual": {"k_red_g2.cu": ["-UG16_FQ2_DUAL", "-DG16_FQ2_DUAL=1"]},
    "g1_red_two_loops": {"k_red_g1.cu": ["-DG16_RED_TWO_LOOPS_G1=1"]},
    "g1_red_two_loops_mb6": {"k_red_g1.cu": ["-DG16_RED_TWO_LOOPS_G1=1", "-DG16_RED_MIN_BLOCKS_G1=6"]},
    "g1_red_mb6": {"k_red_g1.cu": ["-DG16_RED_MIN_BLOCKS_G1=6"]},
    "g2_tile_k4": {"k_red_g2.cu": ["-DG16_TILE_K_G2=4"]},
    "g2_tile_k16": {"k_red_g2.cu": ["-DG16_TILE_K_G2=16"]},
    "g1_tile_k8": {"k_red_g1.cu": ["-DG16_TILE_K_G1=8"]},
    "g2_red_one_loop": {"k_red_g2.cu": ["-DG16_RED_TWO_LOOPS_G2=0"]},
    "g2_red_mb5": {"k_red_g2.cu": ["-DG16_RED_MIN_BLOCKS_G2=5"]},
    "g2_red_mb6": {"k_red_g2.cu": ["-DG16_RED_MIN_BLOCKS_G2=6"]},
    "g2_red_mb8": {"k_red_g2.cu": ["-DG16_RED_MIN_BLOCKS_G2=8"]},
    "g2_red_mb12": {"k_red_g2.cu": ["-DG16_RED_MIN_BLOCKS_G2=12"]},
    "g2_fb_dual": {"k_fbmul_g2.cu": ["-DG16_FQ2_DUAL=1"], "k_fbtab_g2.cu": ["-DG16_FQ2_DUAL=1"]},
    "g2_fb_inline": {"k_fbmul_g2.cu": ["-UG16_COLD"]},
    "g2_comb_fq2_calls": {"k_comb_g2.cu": ["-DG16_COLD_FQ2=1", "-UG16_COLD"]},
    "prove_no_split_tail": {"api.cu": ["-DG16_PROVE_SPLIT_TAIL=0"]},
    "item_floor8": engine("-DG16_ITEM_FLOOR=8"),
    "red_14_15": engine("-DG16_RED_GROUPS_LOG2=14"),
    "red_16_15": engine("-DG16_RED_GROUPS_LOG2=16"),
}


def build_variant(name, units):
    os.makedirs(os.path.join(LAB, "obj"), exist_ok=True)
    std_units = sorted(os.path.basename(p) for p in __import__("glob").glob(os.path.join(b.CSRC, "*.cu")))
    objs = []
    for u in std_units:
        if u in units:
            obj = os.path.join(LAB, "obj", f"{name}_{u.replace('.cu', '.o')}")
            cmd = [b.NVCC] + b.ARCH + b.COMMON + units[u] + ["-c", os.path.join(b.CSRC, u), "-o", obj]
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            open(obj + ".log", "w").write(res.stdout)
            if res.returncode:
                raise RuntimeError(res.stdout[-3000:])
            stats = [l for l in res.stdout.splitlines() if "BucketAccumulate" in l or "Used" in l or "spill" in l]
            # the lines that follow the BucketAccumulate entry
            for i, l in enumerate(res.stdout.splitlines()):
                if "Compiling entry function" in l and ("BucketAccumulate" in l or "accumulate_pair" in l):
                    print(f"[{name}]", " | ".join(x.strip() for x in res.stdout.splitlines()[i + 2:i + 4]))
            objs.append(obj)
        else:
            objs.append(os.path.join(b.OBJDIR, u.replace(".cu", ".o")))
    out = os.path.join(LAB, name + ".so")
    subprocess.run([b.NVCC] + b.ARCH + ["-shared", "-o", out] + objs, check=True)
    return out


def main():
    b.build()
    names = sys.argv[1:] or list(VARIANTS)
    with cf.ThreadPoolExecutor(max_workers=min(len(names), os.cpu_count() or 4)) as ex:
        for out in ex.map(lambda n: build_variant(n, VARIANTS[n]), names):
            print(out)


if __name__ == "__main__":
    main()
