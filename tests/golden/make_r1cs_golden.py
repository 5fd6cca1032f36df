"""Generates tests/golden/r1cs_proofs.json from the reference-semantics model alone (oracle/groth16_ref.py):
compressed proof bytes for the circuits of tests/r1cs_cases.py, fixed SetupParams and fixed (r, s).
    python tests/golden/make_r1cs_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "zero-knowledge-proofs_b200")):
    sys.path.insert(0, p)

import groth16_ref as ref  # noqa: E402
import r1cs_cases as rc  # noqa: E402

out = {}
for ci, mk in enumerate(rc.CIRCUITS):
    constraints, nvars, w, npub = mk()
    qap = ref.QAP(constraints, nvars)
    for pname, params in (("Pverify", ref.P_VERIFY), ("Prand", ref.P_RAND)):
        pk, vk = ref.setup(qap, params, npub)
        proof = ref.prove(pk, w, ref.FIXED_R, ref.FIXED_S)
        out[f"circuit{ci}_{pname}"] = ref.proof_to_bytes(proof).hex()
json.dump(out, open(os.path.join(HERE, "r1cs_proofs.json"), "w"), indent=1)
print(len(out), "proofs written")
