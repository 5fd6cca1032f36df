"""Shared input builders for the parity tests (oracle side only)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def limbs(rows):
    """list of hex-limb lists -> uint64 array"""
    return np.array([[int(v, 16) for v in r] for r in rows], dtype=np.uint64)


def golden_msm_cases(group=None):
    out = []
    for c in load_json("msm_cases.json"):
        if group and c["group"] != group:
            continue
        out.append(dict(
            name=c["name"], group=c["group"], points=limbs(c["points"]), inf=np.array(c["inf"], dtype=np.uint8),
            scalars=limbs(c["scalars"]), result=limbs([c["result"]])[0], result_inf=c["result_inf"],
            compressed=c["result_compressed"]))
    return out


def make_points(oracle, gens, group, seed, n, bits=255):
    """n points k_i * G with k_i from SplitMix64(seed), via the CPU oracle."""
    k = oracle.gen_scalars(seed, n, bits)
    g1, g2 = gens
    th = oracle.max_threads()
    if group == "g1":
        return oracle.g1_fixed_base_mul(g1, k, threads=th)
    return oracle.g2_fixed_base_mul(g2, k, threads=th)


def adversarial(oracle, gens, group, seed, n):
    """points/scalars with ~1% infinities, duplicates, negated duplicates and edge scalars."""
    import bls12_381 as bls
    pts, inf = make_points(oracle, gens, group, seed, n)
    sc = oracle.gen_scalars(seed + 1, n)
    rng = np.random.default_rng(seed)
    width = pts.shape[1]
    half = width // 2
    m = max(1, n // 100)
    idx = rng.permutation(n)
    for i in idx[:m]:                      # infinity bases with non-zero scalars
        pts[i] = 0; inf[i] = 1
    for i in idx[m:2 * m]:                 # duplicates of another base
        j = idx[(3 * m + int(i)) % n]
        pts[i] = pts[j]; inf[i] = inf[j]
    for i in idx[2 * m:3 * m]:             # negated duplicates: y -> q - y (Montgomery form is linear)
        j = idx[(5 * m + int(i)) % n]
        pts[i] = pts[j]; inf[i] = inf[j]
        if not inf[i]:
            zero = np.zeros((half // 6, 6), dtype=np.uint64)
            pts[i, half:] = oracle.fq_sub(zero, pts[i, half:].reshape(-1, 6)).reshape(-1)
    edge = [0, 1, bls.R - 1, (1 << 64) - 1, 2]
    for t, i in enumerate(idx[3 * m:3 * m + 2 * len(edge)]):
        sc[i] = np.array(bls.fr_to_mont(edge[t % len(edge)]), dtype=np.uint64)
    return pts, inf, sc
