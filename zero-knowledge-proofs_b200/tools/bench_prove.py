#!/usr/bin/env python
"""BASELINE config 3: Groth16 prove (4 x G1 + 1 x G2 MSM schedule of Prover::prove,
/root/reference/crates/groth16-core/src/lib.rs:164-271) on synthetic ProvingKey-shaped arrays.

The reference's own R1CS -> QAP -> prove pipeline is Theta(constraints x variables) dense and cannot reach
2^20 constraints, so the arrays are synthesised: a_g1, b_g1, ic_g1, h_g1 = distinct random
G1 points, b_g2 = random G2 points (k_i * G built on the GPU), N = n = 2^log_n variables / H coefficients,
one public input.  Two scalar distributions: "ref" (every scalar < 2^64, what the reference's truncation
produces) and "full" (uniform < r).  Prints one JSON line per distribution with the GPU prove time, the CPU
time of the same five MSMs on the host cores (C port of ark's Pippenger; MSM-only, stated as such) and a
bit-exact comparison of the proof.

    python zero-knowledge-proofs_b200/tools/bench_prove.py --log-n 20 [--steps 3] [--no-cpu]
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
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--no-precompute-bits", action="store_true", help="keep the full-width tables for the 64-bit scalars")
    ap.add_argument("--timeline", action="store_true", help="print the per-lane stage timeline of one prove")
    ap.add_argument("--lib", default=None, help="alternative build of the library (A/B experiments, tools/lab_build.py)")
    args = ap.parse_args()
    import torch
    import bls12_381 as bls
    import cpu_oracle as oracle
    import groth16_cuda
    oracle.build()
    n = 1 << args.log_n
    dev = torch.device("cuda:0")
    ctx = groth16_cuda.Context([0], lib_path=args.lib)
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    g2 = np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64)

    def gen_points(group, seed, count):
        k = oracle.gen_scalars(seed, count)
        d_k = torch.from_numpy(k.view(np.int64)).to(dev)
        width = 24 if group == "g1" else 48
        d_p = torch.empty((count, width), dtype=torch.int32, device=dev)
        ctx.fixed_base_mul_device(group, g1 if group == "g1" else g2, d_k.data_ptr(), count, d_p.data_ptr())
        ctx.synchronize()
        return d_p.cpu().numpy().view(np.uint32).view(np.uint64).reshape(count, width // 2)

    t0 = time.time()
    pk = {"num_public": 1}
    for i, name in enumerate(("a_g1", "b_g1", "ic_g1", "h_g1")):
        pk[name] = gen_points("g1", 0xc0de00 + i, n)
    pk["b_g2"] = gen_points("g2", 0xc0de10, n)
    singles = gen_points("g1", 0xc0de20, 3)
    pk["alpha_g1"], pk["beta_g1"], pk["delta_g1"] = singles[0], singles[1], singles[2]
    singles2 = gen_points("g2", 0xc0de21, 2)
    pk["beta_g2"], pk["delta_g2"] = singles2[0], singles2[1]
    gen_s = time.time() - t0
    t0 = time.time()
    dev_pk = ctx.pk_upload(pk)
    if not args.no_precompute:
        ctx.pk_precompute(dev_pk)
    upload_s = time.time() - t0
    r = oracle.gen_scalars(0xaa, 1)[0]
    s = oracle.gen_scalars(0xbb, 1)[0]
    one = np.array(bls.fr_to_mont(1), dtype=np.uint64)

    for dist, bits in (("full_width", 255), ("ref_faithful_u64", 64)):
        if bits == 64 and not args.no_precompute_bits and not args.no_precompute:
            ctx.pk_precompute(dev_pk, scalar_bits=64)   # tables for the scalars the reference's truncation yields
        w = oracle.gen_scalars(0x1000 + bits, n, bits)
        w[0] = one                                    # assignment[0] is the constant 1
        h = oracle.gen_scalars(0x2000 + bits, n - 1, bits)
        for _ in range(2):
            proof = ctx.prove(dev_pk, w, h, r, s)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            proof = ctx.prove(dev_pk, w, h, r, s)
        gpu_ms = (time.perf_counter() - t0) / args.steps * 1e3
        if args.timeline:   # one more prove with the per-lane stage marks
            import ctypes
            ctx.lib.g16_ctx_enable_stage_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
            ctx.lib.g16_ctx_prove_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            ctx.lib.g16_ctx_enable_stage_timing(ctx.handle, 1)
            ctx.prove(dev_pk, w, h, r, s)
            tl = (ctypes.c_float * 35)()
            ctx._check(ctx.lib.g16_ctx_prove_timeline(ctx.handle, tl))
            ctx.lib.g16_ctx_enable_stage_timing(ctx.handle, 0)
            names = ["pi_A", "pi_B(G2)", "H", "pi_B'", "pi_C"]
            for lane in range(5):
                print("timeline", dist, names[lane], [round(float(tl[7 * lane + k]), 2) for k in range(7)], flush=True)
        line = {"metric": "groth16_prove_ms", "config": f"synthetic ProvingKey, N = n = 2^{args.log_n}, 1 public input, {dist}",
                "gpu_ms": gpu_ms, "n_gpus": 1, "msms": "4 x G1 + 1 x G2 (+ ad-hoc terms)", "pk_generate_s": gen_s,
                "pk_upload_s": upload_s}
        if not args.no_cpu:
            th = oracle.max_threads()
            t0 = time.perf_counter()
            cat = np.concatenate
            zero12 = np.zeros(12, np.uint64)
            # the same sums on the CPU (lib.rs:164-265), all host threads
            a, ai = oracle.g1_msm(cat([pk["alpha_g1"][None], pk["delta_g1"][None], pk["a_g1"]]), None, cat([one[None], r[None], w]), th)
            b, bi = oracle.g2_msm(cat([pk["beta_g2"][None], pk["delta_g2"][None], pk["b_g2"]]), None, cat([one[None], s[None], w]), th)
            hs, hi = oracle.g1_msm(pk["h_g1"][: n - 1], None, h, th)
            b1, b1i = oracle.g1_msm(cat([pk["beta_g1"][None], pk["b_g1"]]), None, cat([one[None], w]), th)
            cpts = cat([pk["ic_g1"][: n - 2], hs[None], a[None], b1[None]])
            cinf = np.zeros(cpts.shape[0], np.uint8); cinf[-3:] = [hi, ai, b1i]
            c, ci = oracle.g1_msm(cpts, cinf, cat([w[2:], one[None], s[None], r[None]]), th)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            ok = ((proof[0][0] == a).all() and (proof[1][0] == b).all() and (proof[2][0] == c).all()
                  and (proof[0][1], proof[1][1], proof[2][1]) == (ai, bi, ci))
            line.update({"cpu_msm_only_ms": cpu_ms, "cpu_threads": th, "cpu_kind": "port (C restatement of ark-ec 0.4.2 msm)",
                         "bit_exact_vs_cpu": bool(ok), "speedup": cpu_ms / gpu_ms})
            assert ok, "GPU proof differs from the CPU oracle"
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
