#!/usr/bin/env python
"""BASELINE config 3 with a REAL constraint system: setup + prove for a synthetic sparse R1CS of 2^log_m
constraints, entirely on one B200 (g16_r1cs_upload -> g16_setup_crs -> g16_pk_precompute -> g16_prove_r1cs).

The reference's dense R1CS -> QAP path (Theta(constraints x variables)) cannot build this key at
all; the sparse kernels (csrc/r1cs_kernels.cuh) give the same values in O(non-zeros).  Circuit: variables
[1, public, 62 seeds, one product per constraint]; constraint i is (c1 x_i1 + c2 x_i2) * (c3 x_i3 + c4 x_i4) =
y_i with 64-bit seeds and small coefficients; every fourth A row also touches the constant (a 2^(log_m - 2)-entry
column, the block-summed path).  With --check the proof is compared with the five MSMs of the C oracle on the
exported key (scalars truncated on the host exactly as the reference does) and the quotient is tied to the
constraint system by the polynomial identity A*B - C = H*Z at a random point in exact arithmetic.

    python zero-knowledge-proofs_b200/tools/bench_r1cs.py --log-m 20 [--steps 3] [--check]
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


def limbs_of(values):
    """python ints (< 2^256) -> n x 4 uint64 canonical limbs"""
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in values), dtype=np.uint64).reshape(-1, 4).copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-m", type=int, default=20)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-precompute", action="store_true")
    args = ap.parse_args()
    import ctypes
    import bls12_381 as bls
    import cpu_oracle as oracle
    import groth16_cuda
    oracle.build()
    R = bls.R
    m = 1 << args.log_m
    n_seed, n_pub = 62, 1
    base = 1 + n_pub + n_seed
    nv = base + m
    rng = np.random.default_rng(0xc1c0 + args.log_m)
    t0 = time.time()
    seeds = [1] + [int(x) for x in rng.integers(1, 1 << 63, size=base - 1, dtype=np.uint64)]
    idx = rng.integers(1, base, size=(m, 4), dtype=np.int64)
    idx[::4, 0] = 0                                  # the constant variable in every fourth A row
    idx[:, 1] = np.where(idx[:, 1] == idx[:, 0], (idx[:, 0] % (base - 1)) + 1, idx[:, 1])   # distinct columns per row
    idx[:, 3] = np.where(idx[:, 3] == idx[:, 2], (idx[:, 2] % (base - 1)) + 1, idx[:, 3])
    coef = rng.integers(1, 17, size=(m, 4), dtype=np.int64)
    w = list(seeds)
    il, cl = idx.tolist(), coef.tolist()
    for (i1, i2, i3, i4), (c1, c2, c3, c4) in zip(il, cl):
        w.append((c1 * seeds[i1] + c2 * seeds[i2]) * (c3 * seeds[i3] + c4 * seeds[i4]) % R)
    w_mont = oracle.fr_to_mont(limbs_of(w))
    small = oracle.fr_to_mont(limbs_of(range(17)))   # Montgomery forms of the coefficients 0..16
    row2 = np.arange(0, 2 * m + 1, 2, dtype=np.uint32)
    row1 = np.arange(0, m + 1, dtype=np.uint32)
    A = (row2, idx[:, :2].reshape(-1).astype(np.uint32), small[coef[:, :2].reshape(-1)])
    B = (row2, idx[:, 2:].reshape(-1).astype(np.uint32), small[coef[:, 2:].reshape(-1)])
    C = (row1, (base + np.arange(m)).astype(np.uint32), small[np.ones(m, dtype=np.int64)])
    gen_s = time.time() - t0

    ctx = groth16_cuda.Context([0])
    t0 = time.perf_counter()
    r1cs = ctx.r1cs_upload(m, nv, A, B, C)
    upload_s = time.perf_counter() - t0
    params = {k: oracle.gen_scalars(0x5e70 + i, 1)[0] for i, k in enumerate(("alpha", "beta", "gamma", "delta", "s"))}
    ctx.setup_crs(r1cs, params, n_pub, want_host=False, want_device_pk=True)[1].free()      # warm-up (tables, workspaces)
    t0 = time.perf_counter()
    crs, pk = ctx.setup_crs(r1cs, params, n_pub, want_host=args.check, want_device_pk=True)
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    if not args.no_precompute:
        ctx.lib.g16_pk_precompute.argtypes = [ctypes.c_void_p] * 2
        ctx._check(ctx.lib.g16_pk_precompute(ctx.handle, pk.handle))
        ctx.synchronize()
    pre_s = time.perf_counter() - t0
    r = oracle.gen_scalars(0xaa, 1)[0]
    s = oracle.gen_scalars(0xbb, 1)[0]
    for _ in range(2):
        proof = ctx.prove_r1cs(pk, r1cs, w_mont, r, s)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        proof = ctx.prove_r1cs(pk, r1cs, w_mont, r, s)
    prove_ms = (time.perf_counter() - t0) / args.steps * 1e3
    line = {"metric": "groth16_setup_prove_r1cs", "config": f"synthetic sparse R1CS, 2^{args.log_m} constraints, {nv} variables, "
            f"{5 * m} non-zeros, 1 public input, reference semantics (64-bit truncations)", "n_gpus": 1,
            "circuit_generate_s": gen_s, "r1cs_upload_s": upload_s, "setup_crs_s": setup_s, "pk_precompute_s": pre_s,
            "prove_ms": prove_ms, "prove_includes": "H2D witness, sparse A/B/C row products, Witness::validate, 7 NTTs (quotient), "
            "truncations, 4 x G1 + 1 x G2 MSM, D2H proof"}
    if args.check:
        th = oracle.max_threads()
        M64 = (1 << 64) - 1
        one = np.array(bls.fr_to_mont(1), dtype=np.uint64)
        # the reference's host-side inputs: truncated assignment (lib.rs:156-161) and truncated H (lib.rs:203-208)
        w_t = oracle.fr_to_mont(limbs_of([x & M64 for x in w]))
        a_ev, b_ev, c_ev = ctx.r1cs_domain_evals(r1cs, w_mont)
        h_mont = ctx.quotient_h(a_ev, b_ev, c_ev)
        h_can = oracle.fr_from_mont(h_mont)
        h_int = [int.from_bytes(x.tobytes(), "little") for x in h_can]
        # (1) polynomial identity at a random point ties rows, columns and quotient together
        x = 0x1234567 * 0x89abcdef % R
        va, vb, vc = (oracle.fr_from_mont(v) for v in ctx.r1cs_eval_at(r1cs, np.array(bls.fr_to_mont(x), dtype=np.uint64)))
        dot = lambda v: sum(wi * int.from_bytes(vi.tobytes(), "little") for wi, vi in zip(w, v)) % R
        Hx = 0
        for cval in reversed(h_int):
            Hx = (Hx * x + cval) % R
        identity = (dot(va) * dot(vb) - dot(vc)) % R == Hx * (pow(x, r1cs.domain_size, R) - 1) % R
        # (2) the five MSMs on the CPU over the exported key
        h_t = oracle.fr_to_mont(limbs_of([c & M64 for c in h_int]))
        cat = np.concatenate
        t0 = time.perf_counter()
        a, ai = oracle.g1_msm(cat([crs["alpha_g1"][None], crs["delta_g1"][None], crs["a_g1"]]), cat([[0, 0], crs["a_g1_inf"]]).astype(np.uint8), cat([one[None], r[None], w_t]), th)
        b, bi = oracle.g2_msm(cat([crs["beta_g2"][None], crs["delta_g2"][None], crs["b_g2"]]), cat([[0, 0], crs["b_g2_inf"]]).astype(np.uint8), cat([one[None], s[None], w_t]), th)
        hs, hi = oracle.g1_msm(crs["h_g1"], crs["h_g1_inf"], h_t, th)
        b1, b1i = oracle.g1_msm(cat([crs["beta_g1"][None], crs["b_g1"]]), cat([[0], crs["b_g1_inf"]]).astype(np.uint8), cat([one[None], w_t]), th)
        cpts = cat([crs["ic_g1"], hs[None], a[None], b1[None]])
        cinf = cat([crs["ic_g1_inf"], [hi, ai, b1i]]).astype(np.uint8)
        c, ci = oracle.g1_msm(cpts, cinf, cat([w_t[n_pub + 1:], one[None], s[None], r[None]]), th)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        ok = ((proof[0][0] == a).all() and (proof[1][0] == b).all() and (proof[2][0] == c).all()
              and (proof[0][1], proof[1][1], proof[2][1]) == (ai, bi, ci))
        line.update({"identity_AB_minus_C_eq_HZ_at_random_point": bool(identity), "cpu_msm_only_ms": cpu_ms, "cpu_threads": th,
                     "bit_exact_vs_cpu": bool(ok)})
        assert identity and ok
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
