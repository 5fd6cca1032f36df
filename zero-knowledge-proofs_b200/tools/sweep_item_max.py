#!/usr/bin/env python
"""Sweep of the work-item limit (g16_ctx_set_item_max) for small and medium MSMs: ms per MSM and the accumulate stage.
    python zero-knowledge-proofs_b200/tools/sweep_item_max.py --group g1 --log-n 16 17 18 19 20"""
import argparse, ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "zero-knowledge-proofs_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="g1")
    ap.add_argument("--log-n", type=int, nargs="+", default=[16, 17, 18, 19, 20])
    ap.add_argument("--item-max", type=int, nargs="+", default=[0, 6, 8, 10, 11, 12, 14, 16, 20, 22, 24, 28, 32, 43, 64, 128])
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch, bls12_381 as bls, cpu_oracle as oracle, groth16_cuda
    oracle.build()
    dev = torch.device("cuda:0")
    ctx = groth16_cuda.Context([0])
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    lib = ctx.lib
    lib.g16_ctx_enable_stage_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.g16_ctx_last_stage_ms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    gen = np.array((bls.g1_to_mont(bls.G1_GEN) if a.group == "g1" else bls.g2_to_mont(bls.G2_GEN))[0], dtype=np.uint64)
    width = 24 if a.group == "g1" else 48
    for log_n in a.log_n:
        n = 1 << log_n
        d_k = torch.from_numpy(oracle.gen_scalars(1, n).view(np.int64)).to(dev)
        d_s = torch.from_numpy(oracle.gen_scalars(2, n).view(np.int64)).to(dev)
        d_p = torch.empty((n, width), dtype=torch.int32, device=dev)
        ctx.fixed_base_mul_device(a.group, gen, d_k.data_ptr(), n, d_p.data_ptr())
        bases = ctx.bases_from_device(a.group, d_p.data_ptr(), n, keepalive=d_p)
        bases.precompute(0)
        out = torch.zeros(width + 1, dtype=torch.int32, device=dev)
        ref = None
        row = {}
        for im in a.item_max:
            ctx.set_item_max(im)
            lib.g16_ctx_enable_stage_timing(ctx.handle, 1)
            for _ in range(3):
                ctx.msm_device(a.group, bases, d_s.data_ptr(), n, out.data_ptr(), 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                ctx.msm_device(a.group, bases, d_s.data_ptr(), n, out.data_ptr(), 0)
            e1.record()
            torch.cuda.synchronize()
            st = (ctypes.c_float * 6)(); plan = (ctypes.c_uint * 3)()
            lib.g16_ctx_last_stage_ms(ctx.handle, st, plan)
            got = out.cpu().numpy().tobytes()
            ref = ref or got
            assert got == ref, f"item_max {im} changes the result"
            row[im] = (round(e0.elapsed_time(e1) / a.steps, 3), round(float(st[3]), 3))
        print(json.dumps({"group": a.group, "log_n": log_n, "plan": list(plan), "ms_total_and_accumulate_by_item_max": row}), flush=True)
        bases.free()


if __name__ == "__main__":
    main()
