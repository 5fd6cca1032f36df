"""BASELINE config 1: toy circuits, fixed SetupParams and fixed (r, s).  The engine's prove schedule
(g16_prove through the C ABI) must give the same proof bytes as the reference-semantics model
(oracle/groth16_ref.py), and the model's verifier decides whether that proof verifies."""
import numpy as np

import bls12_381 as bls
import groth16_ref as ref
import prove_model as pm


def pk_to_arrays(pk):
    def g1s(pts):
        xy = np.array([bls.g1_to_mont(p)[0] for p in pts], dtype=np.uint64).reshape(-1, 12)
        inf = np.array([bls.g1_to_mont(p)[1] for p in pts], dtype=np.uint8)
        return xy, inf
    def g2s(pts):
        xy = np.array([bls.g2_to_mont(p)[0] for p in pts], dtype=np.uint64).reshape(-1, 24)
        inf = np.array([bls.g2_to_mont(p)[1] for p in pts], dtype=np.uint8)
        return xy, inf
    d = {"num_public": pk["num_public"]}
    for k in ("alpha_g1", "beta_g1", "delta_g1"):
        d[k] = np.array(bls.g1_to_mont(pk[k])[0], dtype=np.uint64)
    for k in ("beta_g2", "delta_g2"):
        d[k] = np.array(bls.g2_to_mont(pk[k])[0], dtype=np.uint64)
    for k in ("a_g1", "b_g1", "ic_g1", "h_g1"):
        d[k], d[k + "_inf"] = g1s(pk[k])
    d["b_g2"], d["b_g2_inf"] = g2s(pk["b_g2"])
    return d


def fr_arr(vals):
    return np.array([bls.fr_to_mont(v) for v in vals], dtype=np.uint64).reshape(-1, 4)


def gpu_prove(ctx, pk, assignment, r, s):
    """Host side of Prover::prove up to the MSMs (truncation + quotient, lib.rs:149-208) from the model,
    group part on the engine."""
    w, h = ref.prover_inputs(pk, assignment)
    dev_pk = ctx.pk_upload(pk_to_arrays(pk))
    (a, ai), (b, bi), (c, ci) = ctx.prove(dev_pk, fr_arr(w), fr_arr(h) if h else None, fr_arr([r])[0], fr_arr([s])[0])
    dev_pk.free()
    return (bls.g1_from_mont(list(a), ai), bls.g2_from_mont(list(b), bi), bls.g1_from_mont(list(c), ci))


CASES = [
    # (name, circuit, params, expected verifier verdict under reference semantics -- SURVEY.md App. C)
    ("mul_Pverify", ref.circuit_mul, ref.P_VERIFY, True),
    ("mul_Prand", ref.circuit_mul, ref.P_RAND, False),
    ("cubic_Pverify", ref.circuit_cubic, ref.P_VERIFY, False),
    ("cubic_Prand", ref.circuit_cubic, ref.P_RAND, False),
]


def check_config1(ctx, golden=None):
    out = {}
    for name, circuit, params, verdict in CASES:
        constraints, nvars, w, npub = circuit()
        qap = ref.QAP(constraints, nvars)
        pk, vk = ref.setup(qap, params, npub)
        expect = ref.prove(pk, w, ref.FIXED_R, ref.FIXED_S)
        got = gpu_prove(ctx, pk, w, ref.FIXED_R, ref.FIXED_S)
        assert ref.proof_to_bytes(got) == ref.proof_to_bytes(expect), name
        assert len(ref.proof_to_bytes(got)) == 192 and len(ref.proof_to_bytes(got, compressed=False)) == 384
        assert ref.verify(vk, got, w[1:npub + 1]) is verdict, name
        if verdict:   # a wrong public input must not verify (test_invalid_proof, lib.rs:483-511)
            assert ref.verify(vk, got, [w[1] + 1]) is False
        # the verifier's own MSM (Verifier::verify, lib.rs:330-340: ic[0] + sum x_i ic[i+1], a 2-term call of
        # multi_scalar_mult_g1) through the engine: same point as the model's
        pub = [ref.t64(x) for x in w[1:npub + 1]]
        terms = [(1, vk["ic_g1"][0])] + [(x, vk["ic_g1"][i + 1]) for i, x in enumerate(pub) if x]
        exp_ic = bls.G1.msm_naive([p for _, p in terms], [k for k, _ in terms])
        pts = np.array([bls.g1_to_mont(p)[0] for _, p in terms], dtype=np.uint64).reshape(-1, 12)
        inf = np.array([bls.g1_to_mont(p)[1] for _, p in terms], dtype=np.uint8)
        ic_xy, ic_inf = ctx.multi_scalar_mult_g1(fr_arr([k for k, _ in terms]), pts, inf)
        assert bls.g1_from_mont(list(ic_xy), ic_inf) == exp_ic, name
        out[name] = ref.proof_to_bytes(got).hex()
        if golden is not None:
            assert golden[name] == out[name], name
    return out


def check_random_key(ctx, ctx_single, oracle, gens, n, seed, num_public=2):
    """Random ProvingKey-shaped arrays (k_i * G) of n variables: the proof of `ctx` (any sharding) equals the proof of
    the one-device schedule of `ctx_single`, including arrays shorter than the assignment and H of a different length."""
    import helpers
    pk = {"num_public": num_public}
    for i, name in enumerate(("a_g1", "b_g1", "ic_g1", "h_g1")):
        cnt = {"ic_g1": n - num_public - 1, "h_g1": n + 3}.get(name, n)
        pk[name], pk[name + "_inf"] = helpers.make_points(oracle, gens, "g1", 0x9e00 + 16 * seed + i, cnt)
    pk["b_g2"], pk["b_g2_inf"] = helpers.make_points(oracle, gens, "g2", 0x9e80 + seed, n)
    s1, _ = helpers.make_points(oracle, gens, "g1", 0x9f00 + seed, 3)
    s2, _ = helpers.make_points(oracle, gens, "g2", 0x9f80 + seed, 2)
    pk.update(alpha_g1=s1[0], beta_g1=s1[1], delta_g1=s1[2], beta_g2=s2[0], delta_g2=s2[1])
    w = oracle.gen_scalars(0xa000 + seed, n)
    w[0] = np.array(bls.fr_to_mont(1), dtype=np.uint64)
    w[3] = 0                                            # a zero scalar (the reference filters these out)
    h = oracle.gen_scalars(0xa100 + seed, n - 1)
    r, s = oracle.gen_scalars(0xa200 + seed, 2)
    proofs = []
    for c in (ctx, ctx_single):
        dev_pk = c.pk_upload(pk)
        proofs.append(c.prove(dev_pk, w, h, r, s))
        proofs.append(c.prove(dev_pk, w, None, r, s))  # no H term
        dev_pk.free()
    for got, exp in ((proofs[0], proofs[2]), (proofs[1], proofs[3])):
        for (gx, gi), (ex, ei) in zip(got, exp):
            assert gi == ei and (gx == ex).all()
    # ... and both equal the reference's own five MSMs (lib.rs:179,197,220,255,264) on the CPU (C port of ark's Pippenger)
    assert pm.proofs_equal(proofs[2], pm.five_msms_cpu(pk, w, h, r, s)), "one-device proof differs from the CPU five-MSM model"
    assert pm.proofs_equal(proofs[3], pm.five_msms_cpu(pk, w, None, r, s)), "proof without H differs from the CPU model"


def synthetic_key(ctx, oracle, gens, n, seed, num_public=1, bits=255):
    """ProvingKey-shaped host arrays k_i * G (built by the engine's fixed-base kernels, spot-checked against the
    oracle) together with their exponents."""
    k = pm.synthetic_key_exponents(n, seed, num_public, bits)
    pk = {"num_public": num_public}
    for name in ("a_g1", "b_g1", "ic_g1", "h_g1"):
        pk[name], pk[name + "_inf"] = ctx.fixed_base_mul_g1(gens[0], k[name])
    pk["b_g2"], pk["b_g2_inf"] = ctx.fixed_base_mul_g2(gens[1], k["b_g2"])
    for name in ("alpha_g1", "beta_g1", "delta_g1"):
        pk[name] = oracle.g1_fixed_base_mul(gens[0], k[name][None])[0][0]
    for name in ("beta_g2", "delta_g2"):
        pk[name] = oracle.g2_fixed_base_mul(gens[1], k[name][None])[0][0]
    idx = [0, 1, n // 3, n - num_public - 2]
    for name in ("a_g1", "ic_g1", "h_g1"):
        exp, einf = oracle.g1_fixed_base_mul(gens[0], k[name][idx])
        assert (pk[name][idx] == exp).all() and (pk[name + "_inf"][idx] == einf).all(), name
    exp, einf = oracle.g2_fixed_base_mul(gens[1], k["b_g2"][idx])
    assert (pk["b_g2"][idx] == exp).all()
    return pk, k


def check_prove_in_exponent(ctx, oracle, gens, log_n, seed, precompute, bits_list=(255, 64), cpu_msms=False):
    """BASELINE config 3 / 5 shape: N = n = 2^log_n variables and H coefficients, one public input.  The proof of the
    engine's schedule (g16_prove) must equal the exact computation in the exponent; with cpu_msms also the five MSMs of
    the CPU model (sizes the C port finishes in seconds)."""
    n = 1 << log_n
    pk, k = synthetic_key(ctx, oracle, gens, n, seed)
    dev_pk = ctx.pk_upload(pk)
    if precompute:
        ctx.pk_precompute(dev_pk)
    r, s = oracle.gen_scalars(seed + 0xaa, 2)
    for bits in bits_list:          # full width, and what the reference's 64-bit truncation produces
        if precompute and bits == 64:
            # tables rebuilt for the promised scalar width (g16_pk_precompute_bits); r, s and the (1, r) / (1, s) prefixes
            # stay full width, and one assignment entry breaks the promise on purpose: a hint, not a contract
            ctx.pk_precompute(dev_pk, scalar_bits=64)
        w = oracle.gen_scalars(seed + 0x1000 + bits, n, bits)
        w[0] = pm.ONE
        w[7] = 0
        if precompute and bits == 64 and n > 16:
            w[11] = oracle.gen_scalars(seed + 0x3000, 1)[0]
        h = oracle.gen_scalars(seed + 0x2000 + bits, n - 1, bits)
        got = ctx.prove(dev_pk, w, h, r, s)
        assert pm.proofs_equal(got, pm.proof_in_exponent(k, 1, w, h, r, s, gens)), f"2^{log_n} proof, {bits}-bit scalars"
        if cpu_msms:
            assert pm.proofs_equal(got, pm.five_msms_cpu(pk, w, h, r, s))
    dev_pk.free()
