"""Generates tests/golden/*.json from the exact Python big-integer model (oracle/bls12_381.py).

The reference holds no golden vector for this path, and its Rust sources cannot be
run here, so these fixtures are produced by the independent big-int implementation of the published
curve arithmetic, with naive double-and-add (no bucket method involved).  Everything else -- the C
port of ark's Pippenger, the host-emulated kernels and the CUDA engine -- is checked against them.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
from bls12_381 import *  # noqa


def hx(limbs):
    return [f"{int(v):016x}" for v in limbs]


def case_g1(name, ks, scalars, special=None):
    pts = [G1.mul(G1_GEN, k) for k in ks]
    if special:
        special(pts)
    res = G1.msm_naive(pts, scalars)
    return {
        "name": name, "group": "g1",
        "points": [hx(g1_to_mont(p)[0]) for p in pts], "inf": [g1_to_mont(p)[1] for p in pts],
        "scalars": [hx(fr_to_mont(s)) for s in scalars],
        "result": hx(g1_to_mont(res)[0]), "result_inf": g1_to_mont(res)[1],
        "result_compressed": g1_compress(res).hex(),
    }


def case_g2(name, ks, scalars):
    pts = [G2.mul(G2_GEN, k) for k in ks]
    res = G2.msm_naive(pts, scalars)
    return {
        "name": name, "group": "g2",
        "points": [hx(g2_to_mont(p)[0]) for p in pts], "inf": [g2_to_mont(p)[1] for p in pts],
        "scalars": [hx(fr_to_mont(s)) for s in scalars],
        "result": hx(g2_to_mont(res)[0]), "result_inf": g2_to_mont(res)[1],
        "result_compressed": g2_compress(res).hex(),
    }


def main():
    rng = SplitMix64(0x601d)
    cases = []
    # tiny sizes hit by Verifier::verify (crates/groth16-core/src/lib.rs:340)
    for n in (1, 2, 3, 4, 5):
        cases.append(case_g1(f"g1_random_n{n}", [random_fr(rng) for _ in range(n)], [random_fr(rng) for _ in range(n)]))
    cases.append(case_g1("g1_random_n40", [random_fr(rng) for _ in range(40)], [random_fr(rng) for _ in range(40)]))
    # edge scalars: 0, 1, r-1, 2^64-1 (the reference truncates to 64 bit), full width
    ks = [random_fr(rng) for _ in range(6)]
    cases.append(case_g1("g1_edge_scalars", ks, [0, 1, R - 1, (1 << 64) - 1, 1 << 254, random_fr(rng)]))
    # small CRS-like points k*G with repeated and opposite bases, an infinity base (k = 0)
    cases.append(case_g1("g1_repeated_bases", [5, 5, 5, R - 5, 7, 0, 7, 1], [3, 3, random_fr(rng), 9, 1, random_fr(rng), R - 1, 2]))
    # result is the identity
    cases.append(case_g1("g1_identity_result", [11, 11, 22], [4, R - 6, 1]))
    cases.append(case_g1("g1_all_zero_scalars", [random_fr(rng) for _ in range(4)], [0, 0, 0, 0]))
    for n in (1, 3, 12):
        cases.append(case_g2(f"g2_random_n{n}", [random_fr(rng) for _ in range(n)], [random_fr(rng) for _ in range(n)]))
    cases.append(case_g2("g2_repeated_bases", [5, 5, R - 5, 0, 9], [3, 3, 9, random_fr(rng), R - 1]))
    with open(os.path.join(HERE, "msm_cases.json"), "w") as f:
        json.dump(cases, f, indent=0)

    # fixed-base cases (CRS generation, crates/groth16-setup/src/lib.rs:185-241)
    sc = [0, 1, 2, R - 1, (1 << 64) - 1, 0xdeadbeef, random_fr(rng), random_fr(rng, 64)]
    fb = {
        "scalars": [hx(fr_to_mont(s)) for s in sc],
        "g1": [hx(g1_to_mont(G1.mul(G1_GEN, s))[0]) for s in sc], "g1_inf": [int(s % R == 0) for s in sc],
        "g2": [hx(g2_to_mont(G2.mul(G2_GEN, s))[0]) for s in sc], "g2_inf": [int(s % R == 0) for s in sc],
    }
    with open(os.path.join(HERE, "fixed_base_cases.json"), "w") as f:
        json.dump(fb, f, indent=0)

    # public known answers
    kat = {
        "g1_gen_compressed": g1_compress(G1_GEN).hex(),
        "g2_gen_compressed": g2_compress(G2_GEN).hex(),
        "g1_2g_x": f"{G1.mul(G1_GEN, 2)[0]:096x}",
        "fq_R": f"{FQ_R:096x}", "fq_R2": f"{FQ_R2:096x}", "fq_ninv64": f"{FQ_NINV64:016x}",
        "fr_R": f"{FR_R:064x}", "fr_R2": f"{FR_R2:064x}", "fr_ninv64": f"{FR_NINV64:016x}",
    }
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote", len(cases), "msm cases")


if __name__ == "__main__":
    main()
