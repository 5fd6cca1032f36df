"""CPU-only: the reference-semantics model (oracle/groth16_ref.py) reproduces the behaviour predicted
for the reference in SURVEY.md App. C, its proofs are pinned as golden bytes, and the engine's prove
schedule (host-emulated kernels) reproduces those bytes."""
import json
import os

import pytest

import groth16_ref as ref
import prove_cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_proofs.json")


def test_domain_and_truncation():
    import bls12_381 as bls
    d = ref.Domain(4)
    assert d.size == 4 and pow(d.group_gen, 4, bls.R) == 1 and pow(d.group_gen, 2, bls.R) != 1
    assert ref.Domain(1).size == 1 and ref.Domain(3).size == 4
    assert ref.t64((1 << 64) + 5) == 5 and ref.t64(bls.R - 1) == (bls.R - 1) & ((1 << 64) - 1)
    c, nv, w, npub = ref.circuit_cubic()
    q = ref.QAP(c, nv)
    assert q.n == 4 and q.degree() == 4                      # qap.degree() == domain size (lib.rs:285-294)
    a, b, cc = q.evaluate_at(q.domain.group_gen, w)
    assert a * b % bls.R == cc                               # Witness::validate (lib.rs:112-131)
    assert len(q.quotient(w)) == 3                           # App. C: 3 non-zero H coefficients
    with pytest.raises(ValueError):
        q.quotient([1, 35, 3, 9, 27, 31])


def test_emu_engine_reproduces_reference_proofs(emu_ctx):
    golden = json.load(open(GOLDEN)) if os.path.exists(GOLDEN) else None
    out = prove_cases.check_config1(emu_ctx, golden)
    if golden is None:   # first run writes the fixture (committed afterwards)
        json.dump(out, open(GOLDEN, "w"), indent=1)
