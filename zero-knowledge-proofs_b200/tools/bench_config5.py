#!/usr/bin/env python
"""BASELINE config 5: Groth16 setup + prove at 2^log_n constraints on the G GPUs of one box, through the
reference-facing C ABI with HOST buffers (one process, multi-device context: the path a Rust prover takes).

  setup   the group work of CRS::generate_from_qap (/root/reference/crates/groth16-setup/src/lib.rs:185-241):
          5 x n G1 + n G2 fixed-base multiplications of the generators, index-range shards, no collective
          ("5 x 2^24 G1 + 2^24 G2")
  prove   the 4 x G1 + 1 x G2 MSM schedule of Prover::prove (crates/groth16-core/src/lib.rs:164-271) on the
          arrays setup produced, every array sharded by index range, 192/384-byte partial sums folded on GPU 0

    python zero-knowledge-proofs_b200/tools/bench_config5.py --gpus 8 --log-n 24 [--steps 3] [--check-exponent]

Scalars: "ref_faithful_u64" (uniform < 2^64, what the reference's truncation produces) and "full_width".
Checks at full size: a bounded sample of every setup array against the CPU oracle; with --check-exponent the
whole proof in the exponent (every base is k_i * G with known k_i, so proof.a = (k_alpha + sum w_i ka_i +
r k_delta) * G etc. -- an exact O(n) big-integer computation, independent of the GPU path).  The compressed proof
(192 bytes, ark encoding, produced by g16_proof_serialize) is printed so that runs at different GPU counts can be
compared byte for byte."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (os.path.join(ROOT, "zero-knowledge-proofs_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402


def mont_ints(limbs):
    """n x 4 u64 Montgomery limbs -> python ints (still Montgomery: value * 2^256 mod r)."""
    raw = np.ascontiguousarray(limbs).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--sample", type=int, default=1 << 10)
    ap.add_argument("--check-exponent", action="store_true")
    ap.add_argument("--no-precompute", action="store_true")
    ap.add_argument("--pageable", action="store_true", help="plain numpy host buffers instead of pinned host memory")
    ap.add_argument("--lib", default=None, help="library path (tests only: the host-emulation build)")
    args = ap.parse_args()
    import ctypes
    import bls12_381 as bls
    import cpu_oracle as oracle
    import groth16_cuda
    oracle.build()
    n = 1 << args.log_n
    ctx = groth16_cuda.Context(list(range(args.gpus)), lib_path=args.lib)
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    g2 = np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64)
    th = oracle.max_threads()
    R = bls.R
    if args.pageable or args.lib:
        pin = lambda a: a                                    # noqa: E731
        host_kind = "pageable"
    else:
        import torch                                         # pinned host memory only (torch owns the allocator)

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64 else np.ascontiguousarray(a))
            return t.pin_memory().numpy().view(a.dtype)
        host_kind = "pinned"

    # ---- setup: fixed-base multiplications, host scalars -> host points ---------------------------------
    names = ("a_g1", "b_g1", "ic_g1", "h_g1", "vk_ic_g1")
    for dist, bits in (("ref_faithful_u64", 64), ("full_width", 255)):
        ks = {nm: pin(oracle.gen_scalars(0xc5e700 + 16 * i + bits, n, bits)) for i, nm in enumerate(names + ("b_g2",))}
        outs = {nm: (pin(np.zeros((n, 24 if nm == "b_g2" else 12), dtype=np.uint64)), pin(np.zeros(n, dtype=np.uint8)))
                for nm in names + ("b_g2",)}
        ctx.fixed_base_mul_g1(g1, ks["a_g1"], out=outs["a_g1"])   # warm-up at full size: tables, workspaces
        ctx.fixed_base_mul_g2(g2, ks["b_g2"], out=outs["b_g2"])
        t0 = time.perf_counter()
        pts = {}
        for nm in names:
            pts[nm] = ctx.fixed_base_mul_g1(g1, ks[nm], out=outs[nm])
        t_g1 = time.perf_counter() - t0
        pts["b_g2"] = ctx.fixed_base_mul_g2(g2, ks["b_g2"], out=outs["b_g2"])
        t_all = time.perf_counter() - t0
        m = min(n, args.sample)
        ok = True
        t0 = time.perf_counter()
        for nm in names + ("b_g2",):
            f = oracle.g2_fixed_base_mul if nm == "b_g2" else oracle.g1_fixed_base_mul
            exp, einf = f(g2 if nm == "b_g2" else g1, ks[nm][:m], threads=th)
            ok &= bool((pts[nm][0][:m] == exp).all() and (pts[nm][1][:m] == einf).all())
            # and the last elements (the last shard)
            exp, einf = f(g2 if nm == "b_g2" else g1, ks[nm][-8:], threads=th)
            ok &= bool((pts[nm][0][-8:] == exp).all() and (pts[nm][1][-8:] == einf).all())
        cpu_s = time.perf_counter() - t0
        cpu_rate = (6 * (m + 8)) / cpu_s
        print(json.dumps({"metric": "groth16_setup_group_ms", "config": f"5 x 2^{args.log_n} G1 + 2^{args.log_n} G2 fixed-base, {dist}",
                          "n_gpus": args.gpus, "ms": t_all * 1e3, "g1_ms": t_g1 * 1e3, "g2_ms": (t_all - t_g1) * 1e3,
                          "points_per_s": 6 * n / t_all, "through": f"C ABI, host scalars in, host points out ({host_kind} host memory)",
                          "h2d_bytes": 6 * n * 32, "d2h_bytes": n * (5 * 97 + 193),
                          "cpu_points_per_s_mixed": cpu_rate, "cpu_threads": th, "bit_exact_sample": ok,
                          "sample": f"first {m} + last 8 of each array"}), flush=True)
        assert ok, "setup sample differs from the CPU oracle"

    # ---- prove on the arrays of the last (full-width) setup ---------------------------------------------
    singles_k = oracle.gen_scalars(0xc5e7ff, 5)
    singles, _ = ctx.fixed_base_mul_g1(g1, singles_k[:3])
    singles2, _ = ctx.fixed_base_mul_g2(g2, singles_k[3:5])
    pk = {"num_public": 1, "alpha_g1": singles[0], "beta_g1": singles[1], "delta_g1": singles[2],
          "beta_g2": singles2[0], "delta_g2": singles2[1]}
    for nm in ("a_g1", "b_g1", "ic_g1", "h_g1", "b_g2"):
        pk[nm] = pts[nm][0]
    pk["ic_g1"] = pk["ic_g1"][: n - 2]
    t0 = time.perf_counter()
    dev_pk = ctx.pk_upload(pk)
    upload_s = time.perf_counter() - t0
    pre_s = 0.0
    if not args.no_precompute:
        t0 = time.perf_counter()
        ctx.lib.g16_pk_precompute.argtypes = [ctypes.c_void_p] * 2
        ctx._check(ctx.lib.g16_pk_precompute(ctx.handle, dev_pk.handle))
        pre_s = time.perf_counter() - t0
    r = oracle.gen_scalars(0xaa, 1)[0]
    s = oracle.gen_scalars(0xbb, 1)[0]
    one = np.array(bls.fr_to_mont(1), dtype=np.uint64)
    for dist, bits in (("ref_faithful_u64", 64), ("full_width", 255)):
        w = oracle.gen_scalars(0x1000 + bits, n, bits)
        w[0] = one
        w = pin(w)
        h = pin(oracle.gen_scalars(0x2000 + bits, n - 1, bits))
        for _ in range(2):
            proof = ctx.prove(dev_pk, w, h, r, s)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            proof = ctx.prove(dev_pk, w, h, r, s)
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        line = {"metric": "groth16_prove_ms", "config": f"ProvingKey from the setup above, N = n = 2^{args.log_n}, 1 public input, {dist}",
                "n_gpus": args.gpus, "gpu_ms": ms, "msms": "4 x G1 + 1 x G2 (+ ad-hoc terms)", "pk_upload_s": upload_s,
                "pk_precompute_s": pre_s, "through": f"g16_prove: host assignment + H coefficients in ({host_kind} host memory), proof out",
                "h2d_bytes_per_prove": (2 * n - 1) * 32, "d2h_bytes_per_prove": 51 * 4,
                "proof_compressed": ctx.proof_serialize(*proof).hex() if args.gpus == 1 else None}
        if args.gpus > 1:
            # the wire-format entry points are single-device: serialise on a one-GPU context
            c1 = groth16_cuda.Context([0], lib_path=args.lib)
            line["proof_compressed"] = c1.proof_serialize(*proof).hex()
            c1.close()
        if args.check_exponent:
            t0 = time.perf_counter()
            RI = bls.FR_RINV
            wi, hi_ = mont_ints(w), mont_ints(h)
            kint = {nm: mont_ints(ks[nm]) for nm in ("a_g1", "b_g1", "ic_g1", "h_g1", "b_g2")}
            sk = [v * RI % R for v in mont_ints(singles_k)]
            ri, si = int(mont_ints(r[None])[0]) * RI % R, int(mont_ints(s[None])[0]) * RI % R
            dot = lambda x, y: sum(map(int.__mul__, x, y)) * RI * RI % R   # noqa: E731  (both operands Montgomery)
            e_a = (sk[0] + dot(wi, kint["a_g1"]) + ri * sk[2]) % R
            e_b = (sk[3] + dot(wi, kint["b_g2"]) + si * sk[4]) % R
            e_b1 = (sk[1] + dot(wi, kint["b_g1"])) % R
            e_c = (dot(wi[2:], kint["ic_g1"][: n - 2]) + dot(hi_, kint["h_g1"][: n - 1]) + si * e_a + ri * e_b1) % R
            exp = bls.proof_bytes(bls.G1.mul(bls.G1_GEN, e_a), bls.G2.mul(bls.G2_GEN, e_b), bls.G1.mul(bls.G1_GEN, e_c)).hex()
            line["exponent_check"] = {"bit_exact": exp == line["proof_compressed"], "seconds": time.perf_counter() - t0,
                                      "what": "proof recomputed in the exponent with exact big integers (O(n)), compressed bytes compared"}
            assert exp == line["proof_compressed"], "proof differs from the exponent computation"
        print(json.dumps(line), flush=True)
    dev_pk.free()
    ctx.close()


if __name__ == "__main__":
    main()
