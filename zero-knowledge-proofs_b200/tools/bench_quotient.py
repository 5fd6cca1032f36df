#!/usr/bin/env python
"""Quotient polynomial H = (A*B - C)/Z on the GPU (g16_quotient_h): timing through the C ABI with HOST
buffers (3 x n x 32 B in, n x 32 B out, pinned), and device-resident through g16_quotient_h_device, n = 2^log_n.

    python zero-knowledge-proofs_b200/tools/bench_quotient.py --log-n 20

Inputs: random a, b evaluations, c = a*b on the domain (a satisfied system).  The result is checked by the
polynomial identity A(x0)*B(x0) - C(x0) = H(x0) * (x0^n - 1) at a random point x0, with A, B, C, H evaluated
on the host from their domain values / coefficients (barycentric formula / Horner) in exact integers.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "zero-knowledge-proofs_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import bls12_381 as bls, cpu_oracle as oracle, groth16_cuda, groth16_ref as ref
    oracle.build()
    n = 1 << a.log_n
    ctx = groth16_cuda.Context([0])
    ae = oracle.gen_scalars(0xa0, n)
    be = oracle.gen_scalars(0xb0, n)
    # c = a * b (canonical product, back to Montgomery) via python ints on a sample-free vector path
    av = oracle.fr_from_mont(ae); bv = oracle.fr_from_mont(be)
    to_int = lambda x: [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in x]
    ai, bi = to_int(av), to_int(bv)
    ci = [x * y % bls.R for x, y in zip(ai, bi)]
    ce = oracle.fr_to_mont(np.array([bls.int_to_limbs64(v, 4) for v in ci], dtype=np.uint64))
    import torch
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).pin_memory().numpy().view(np.uint64)   # noqa: E731
    ae, be, ce = pin(ae), pin(be), pin(ce)
    h = pin(np.zeros((n, 4), dtype=np.uint64))
    ctx.quotient_h(ae, be, ce, out=h)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ctx.quotient_h(ae, be, ce, out=h)
    ms = (time.perf_counter() - t0) / a.steps * 1e3
    # device-resident: CUDA events around g16_quotient_h_device on torch's stream (the input block is restored by a
    # device-to-device copy outside the event pairs, the transform overwrites it)
    dev = torch.device("cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    src = torch.from_numpy(np.concatenate([ae, be, ce]).view(np.int64)).to(dev)
    work = torch.empty_like(src)
    d_h = torch.empty((n, 4), dtype=torch.int64, device=dev)
    d_bad = torch.zeros(1, dtype=torch.int32, device=dev)
    dev_ms = []
    for it in range(a.steps + 2):
        work.copy_(src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.quotient_h_device(work.data_ptr(), n, d_h.data_ptr(), d_bad.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            dev_ms.append(e0.elapsed_time(e1))
    assert int(d_bad.item()) == 0 and (d_h.cpu().numpy().view(np.uint64) == h).all(), "device-resident result differs"
    dev_ms = float(np.mean(dev_ms))
    # 7 transforms x log2(n) stages x n/2 butterflies, one Fr multiplication each; HBM traffic of the fused passes
    passes = -(-a.log_n // 7)
    hbm_bytes = 7 * passes * n * 64 + 3 * n * 32 * 2 + n * 32 * 3
    # identity check at a random point (barycentric evaluation of A, B, C from domain values)
    hi = to_int(oracle.fr_from_mont(h))
    x0 = 0x1234567890abcdef1234567890abcdef % bls.R
    dom = ref.Domain(n)
    w, zx = dom.group_gen, (pow(x0, n, bls.R) - 1) % bls.R
    ninv = pow(n, -1, bls.R)
    def bary(vals):
        # P(x0) = (x0^n - 1)/n * sum_i vals_i * w^i / (x0 - w^i)
        acc, wi = 0, 1
        dens = []
        for i in range(n):
            dens.append((x0 - wi) % bls.R); wi = wi * w % bls.R
        # batch inverse
        pref = [1] * (n + 1)
        for i, d in enumerate(dens): pref[i + 1] = pref[i] * d % bls.R
        inv_all = pow(pref[n], -1, bls.R)
        wi_list = [1] * n
        for i in range(1, n): wi_list[i] = wi_list[i - 1] * w % bls.R
        for i in range(n - 1, -1, -1):
            di = inv_all * pref[i] % bls.R
            inv_all = inv_all * dens[i] % bls.R
            acc += vals[i] * wi_list[i] % bls.R * di
        return acc % bls.R * zx % bls.R * ninv % bls.R
    lhs = (bary(ai) * bary(bi) - bary(ci)) % bls.R
    rhs = ref.poly_eval(hi, x0) * zx % bls.R
    ok = lhs == rhs
    print(json.dumps({"metric": "quotient_h_ms", "n": n, "gpu_ms_device_resident": dev_ms, "gpu_ms_host_to_host": ms,
                      "identity_check_at_random_point": bool(ok), "h2d_bytes": 3 * n * 32, "d2h_bytes": n * 32,
                      "fr_mul_per_s": 7 * a.log_n * (n // 2) / (dev_ms * 1e-3),
                      "hbm": {"algorithmic_bytes": hbm_bytes, "achieved_GBps": hbm_bytes / (dev_ms * 1e-3) / 1e9,
                              "passes_per_transform": passes}}), flush=True)
    assert ok


if __name__ == "__main__":
    main()
