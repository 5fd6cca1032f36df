"""CPU checks of oracle/prove_model.py (the checker of the prove-schedule parity tests): the five-MSM model through
the C port of ark's Pippenger against the reference-semantics big-integer model on the toy circuits, and against the
exact computation in the exponent on a synthetic key."""
import numpy as np

import bls12_381 as bls
import groth16_ref as ref
import prove_cases
import prove_model as pm


def test_five_msms_match_reference_model_on_toy_circuits():
    for name, circuit, params, _ in prove_cases.CASES:
        constraints, nvars, w, npub = circuit()
        pk, _vk = ref.setup(ref.QAP(constraints, nvars), params, npub)
        expect = ref.prove(pk, w, ref.FIXED_R, ref.FIXED_S)
        wt, h = ref.prover_inputs(pk, w)
        fr = prove_cases.fr_arr
        (a, ai), (b, bi), (c, ci) = pm.five_msms_cpu(prove_cases.pk_to_arrays(pk), fr(wt), fr(h) if h else None,
                                                     fr([ref.FIXED_R])[0], fr([ref.FIXED_S])[0], threads=2)
        got = (bls.g1_from_mont(list(a), ai), bls.g2_from_mont(list(b), bi), bls.g1_from_mont(list(c), ci))
        assert ref.proof_to_bytes(got) == ref.proof_to_bytes(expect), name


def test_exponent_model_matches_five_msms(oracle, gens):
    n, npub = 57, 2
    k = pm.synthetic_key_exponents(n, 0x5e7, npub)
    pk = {"num_public": npub}
    for name in ("a_g1", "b_g1", "ic_g1", "h_g1"):
        pk[name], pk[name + "_inf"] = oracle.g1_fixed_base_mul(gens[0], k[name], threads=oracle.max_threads())
    pk["b_g2"], pk["b_g2_inf"] = oracle.g2_fixed_base_mul(gens[1], k["b_g2"], threads=oracle.max_threads())
    for name in ("alpha_g1", "beta_g1", "delta_g1"):
        pk[name] = oracle.g1_fixed_base_mul(gens[0], k[name][None])[0][0]
    for name in ("beta_g2", "delta_g2"):
        pk[name] = oracle.g2_fixed_base_mul(gens[1], k[name][None])[0][0]
    w = oracle.gen_scalars(0x77, n)
    w[0] = pm.ONE
    w[5] = 0
    h = oracle.gen_scalars(0x78, n - 1)
    r, s = oracle.gen_scalars(0x79, 2)
    for hh in (h, None):
        assert pm.proofs_equal(pm.five_msms_cpu(pk, w, hh, r, s), pm.proof_in_exponent(k, npub, w, hh, r, s, gens))
    # dot_mod_r against Python integers
    a = oracle.fr_from_mont(w); b = oracle.fr_from_mont(k["a_g1"])
    val = lambda v: sum(int(v[i]) << (64 * i) for i in range(4))
    e = sum(val(x) * val(y) for x, y in zip(a, b)) % bls.R
    assert list(oracle.dot_mod_r(w, k["a_g1"], threads=3)) == list(np.array(bls.fr_to_mont(e), dtype=np.uint64))
