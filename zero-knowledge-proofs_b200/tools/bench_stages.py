#!/usr/bin/env python
"""Stage-by-stage CUDA-event timing of one MSM (G1 or G2), device-resident inputs.
    python zero-knowledge-proofs_b200/tools/bench_stages.py --group g2 --log-n 20 [--no-precompute]"""
import argparse, ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "zero-knowledge-proofs_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="g2")
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--bits", type=int, default=255)
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--precompute-bits", type=int, default=0)
    ap.add_argument("--lib", default=None, help="alternative build of the library (A/B experiments)")
    a = ap.parse_args()
    import torch, bls12_381 as bls, cpu_oracle as oracle, groth16_cuda
    oracle.build()
    n = 1 << a.log_n
    dev = torch.device("cuda:0")
    ctx = groth16_cuda.Context([0], lib_path=a.lib)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    lib = ctx.lib
    lib.g16_ctx_enable_stage_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.g16_ctx_last_stage_ms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    gen = np.array((bls.g1_to_mont(bls.G1_GEN) if a.group == "g1" else bls.g2_to_mont(bls.G2_GEN))[0], dtype=np.uint64)
    width = 24 if a.group == "g1" else 48
    d_k = torch.from_numpy(oracle.gen_scalars(1, n).view(np.int64)).to(dev)
    d_s = torch.from_numpy(oracle.gen_scalars(2, n, a.bits).view(np.int64)).to(dev)
    d_p = torch.empty((n, width), dtype=torch.int32, device=dev)
    t0 = time.perf_counter()
    ctx.fixed_base_mul_device(a.group, gen, d_k.data_ptr(), n, d_p.data_ptr())
    torch.cuda.synchronize()
    fb_ms = (time.perf_counter() - t0) * 1e3
    bases = ctx.bases_from_device(a.group, d_p.data_ptr(), n, keepalive=d_p)
    pre = 0
    t0 = time.perf_counter()
    if not a.no_precompute:
        pre = bases.precompute(a.precompute_bits)
    torch.cuda.synchronize()
    pre_ms = (time.perf_counter() - t0) * 1e3
    out = torch.zeros(49, dtype=torch.int32, device=dev)
    lib.g16_ctx_enable_stage_timing(ctx.handle, 1)
    for _ in range(2):
        ctx.msm_device(a.group, bases, d_s.data_ptr(), n, out.data_ptr(), 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ctx.msm_device(a.group, bases, d_s.data_ptr(), n, out.data_ptr(), 0)
    e1.record()
    torch.cuda.synchronize()
    st = (ctypes.c_float * 6)(); plan = (ctypes.c_uint * 3)()
    lib.g16_ctx_last_stage_ms(ctx.handle, st, plan)
    print(json.dumps({"group": a.group, "log_n": a.log_n, "scalar_bits": a.bits, "ms": e0.elapsed_time(e1) / a.steps,
                      "points_per_s": n / (e0.elapsed_time(e1) / a.steps * 1e-3), "precompute_c": pre, "precompute_ms": pre_ms,
                      "fixed_base_ms": fb_ms, "fixed_base_points_per_s": n / (fb_ms * 1e-3), "plan": list(plan),
                      "stage_ms": dict(zip(["digits", "scan_items", "scatter", "accumulate", "reduce", "combine"], [round(float(x), 3) for x in st]))}))


if __name__ == "__main__":
    main()
