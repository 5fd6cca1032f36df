#!/usr/bin/env python
"""BASELINE config 5 (setup part): fixed-base batch scalar multiplication of the generator, the group work
of CRS::generate_from_qap (/root/reference/crates/groth16-setup/src/lib.rs:185-241), on one GPU.

    python zero-knowledge-proofs_b200/tools/bench_setup.py --log-n 22 [--steps 3]

Scalars: "ref" = uniform < 2^64 (what the reference's low-64-bit truncation produces), "full" = uniform < r.
Device-resident timing (scalars and outputs in HBM) with CUDA events; the CPU figure is the C port of ark's
double-and-add + into_affine on all host threads over a bounded sample; a sample of the GPU output is
compared bit-for-bit with it.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "zero-knowledge-proofs_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=1 << 13)
    ap.add_argument("--lib", default=None, help="alternative build of the library (A/B experiments)")
    ap.add_argument("--groups", default="g1,g2")
    a = ap.parse_args()
    import torch
    import bls12_381 as bls
    import cpu_oracle as oracle
    import groth16_cuda
    oracle.build()
    n = 1 << a.log_n
    dev = torch.device("cuda:0")
    ctx = groth16_cuda.Context([0], lib_path=a.lib)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    gens = {"g1": np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64),
            "g2": np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64)}
    th = oracle.max_threads()
    for group, width in (("g1", 24), ("g2", 48)):
        if group not in a.groups.split(","):
            continue
        for dist, bits in (("ref_faithful_u64", 64), ("full_width", 255)):
            k = oracle.gen_scalars(0x5e70 + bits, n, bits)
            d_k = torch.from_numpy(k.view(np.int64)).to(dev)
            d_out = torch.empty((n, width), dtype=torch.int32, device=dev)
            for _ in range(2):
                ctx.fixed_base_mul_device(group, gens[group], d_k.data_ptr(), n, d_out.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                ctx.fixed_base_mul_device(group, gens[group], d_k.data_ptr(), n, d_out.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            m = min(n, a.cpu_sample)
            t0 = time.perf_counter()
            exp, einf = (oracle.g1_fixed_base_mul if group == "g1" else oracle.g2_fixed_base_mul)(gens[group], k[:m], threads=th)
            cpu_s = time.perf_counter() - t0
            got = d_out[:m].cpu().numpy().view(np.uint32).view(np.uint64).reshape(m, width // 2)
            ok = bool((got == exp).all())
            print(json.dumps({"metric": "fixed_base_points_per_sec", "group": group, "scalars": dist, "n": n, "gpu_ms": ms,
                              "gpu_points_per_s": n / (ms * 1e-3), "cpu_points_per_s": m / cpu_s, "cpu_threads": th,
                              "cpu_sample": m, "bit_exact_sample": ok, "speedup": (n / (ms * 1e-3)) / (m / cpu_s)}), flush=True)
            assert ok


if __name__ == "__main__":
    main()
