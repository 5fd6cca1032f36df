"""CPU model of the group part of `Prover::prove` on array-shaped keys -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/ and bench.py's cpu_baseline / --impl reference legs may import this.

Restates the MSM schedule of /root/reference/crates/groth16-core/src/lib.rs:164-271 on packed ark-layout
arrays (the layout of include/g16_cuda.h), in two independent ways:

  five_msms_cpu       the five `multi_scalar_mult_*` calls (lib.rs:179,197,220,255,264) through the C port of
                      ark-ec 0.4.2's Pippenger (oracle/cpu_msm.c) -- the reference's own CPU path, also the CPU
                      baseline of the prove benchmark;
  proof_in_exponent   when every base is a known multiple k_i * G of the generator, the proof is
                      A = (k_alpha + <w, ka> + r k_delta) G1, ... : an exact O(n) computation in Fr that does not
                      share a line of group code with either MSM implementation (size independent: used at 2^20
                      and 2^24, where the CPU MSMs take seconds to minutes).
PARITY UNPINNED by the reference (no golden vectors upstream).
"""
from __future__ import annotations

import numpy as np

import bls12_381 as bls
import cpu_oracle as oracle

ONE = np.array(bls.fr_to_mont(1), dtype=np.uint64)


def five_msms_cpu(pk: dict, w, h, r, s, threads: int = 0):
    """pk: dict of numpy arrays as groth16_cuda.Context.pk_upload takes (optional *_inf flags); w: num_vars x 4
    Montgomery Fr (already truncated `assignment_fr`, lib.rs:156-161); h: coefficients or None.
    Returns ((a_xy, a_inf), (b_xy, b_inf), (c_xy, c_inf)) exactly as Context.prove does."""
    th = threads or oracle.max_threads()
    w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
    nv = w.shape[0]
    npub = int(pk["num_public"])

    def arr(name, width):
        xy = np.ascontiguousarray(pk[name], dtype=np.uint64).reshape(-1, width)
        inf = pk.get(name + "_inf")
        inf = np.zeros(xy.shape[0], np.uint8) if inf is None else np.ascontiguousarray(inf, dtype=np.uint8)
        return xy, inf

    def single(name):
        p = np.ascontiguousarray(pk[name], dtype=np.uint64).reshape(1, -1)
        return p, np.array([0 if p.any() else 1], dtype=np.uint8)

    def msm(g, singles, single_scalars, name, width, scalars):
        xy, inf = arr(name, width)
        m = min(xy.shape[0], scalars.shape[0])
        pts = np.concatenate([p for p, _ in singles] + [xy[:m]])
        fl = np.concatenate([f for _, f in singles] + [inf[:m]])
        sc = np.concatenate([np.asarray(x, np.uint64).reshape(1, 4) for x in single_scalars] + [scalars[:m]])
        return (oracle.g1_msm if g == 1 else oracle.g2_msm)(pts, fl, sc, threads=th)

    r = np.asarray(r, np.uint64).reshape(4)
    s = np.asarray(s, np.uint64).reshape(4)
    a, ai = msm(1, [single("alpha_g1"), single("delta_g1")], [ONE, r], "a_g1", 12, w)                 # lib.rs:164-179
    b, bi = msm(2, [single("beta_g2"), single("delta_g2")], [ONE, s], "b_g2", 24, w)                  # lib.rs:182-197
    if h is not None and len(h):
        hs, hi = msm(1, [], [], "h_g1", 12, np.ascontiguousarray(h, dtype=np.uint64).reshape(-1, 4))  # lib.rs:200-220
    else:
        hs, hi = np.zeros(12, np.uint64), 1
    b1, b1i = msm(1, [single("beta_g1")], [ONE], "b_g1", 12, w)                                       # lib.rs:243-255
    ic_xy, ic_inf = arr("ic_g1", 12)
    priv = w[npub + 1:]
    m = min(ic_xy.shape[0], priv.shape[0])
    cpts = np.concatenate([ic_xy[:m], hs[None], a[None], b1[None]])
    cinf = np.concatenate([ic_inf[:m], np.array([hi, ai, b1i], np.uint8)])
    csc = np.concatenate([priv[:m], ONE[None], s[None], r[None]])
    c, ci = oracle.g1_msm(cpts, cinf, csc, threads=th)                                                # lib.rs:223-264
    assert nv > npub
    return (a, ai), (b, bi), (c, ci)


def _int(mont4) -> int:
    v = oracle.fr_from_mont(np.asarray(mont4, np.uint64).reshape(1, 4))[0]
    return sum(int(v[i]) << (64 * i) for i in range(4))


def proof_in_exponent(k: dict, num_public: int, w, h, r, s, gens, threads: int = 0):
    """k: Montgomery Fr exponents of every key element (`a_g1`, `b_g1`, `b_g2`, `ic_g1`, `h_g1`: n x 4;
    `alpha_g1`, `beta_g1`, `delta_g1`, `beta_g2`, `delta_g2`: 4).  Returns the proof as Context.prove does."""
    th = threads or oracle.max_threads()
    w = np.ascontiguousarray(w, dtype=np.uint64).reshape(-1, 4)
    R = bls.R

    def dot(name, scalars):
        kk = np.ascontiguousarray(k[name], dtype=np.uint64).reshape(-1, 4)
        m = min(kk.shape[0], scalars.shape[0])
        return _int(oracle.dot_mod_r(kk[:m], scalars[:m], th)) if m else 0

    ri, si = _int(r), _int(s)
    ea = (_int(k["alpha_g1"]) + dot("a_g1", w) + ri * _int(k["delta_g1"])) % R
    eb = (_int(k["beta_g2"]) + dot("b_g2", w) + si * _int(k["delta_g2"])) % R
    eh = dot("h_g1", np.ascontiguousarray(h, dtype=np.uint64).reshape(-1, 4)) if h is not None and len(h) else 0
    eb1 = (_int(k["beta_g1"]) + dot("b_g1", w)) % R
    ec = (dot("ic_g1", w[num_public + 1:]) + eh + si * ea + ri * eb1) % R
    g1, g2 = gens
    mont = lambda e: np.array([bls.fr_to_mont(e)], dtype=np.uint64)
    a, ai = oracle.g1_fixed_base_mul(g1, mont(ea))
    b, bi = oracle.g2_fixed_base_mul(g2, mont(eb))
    c, ci = oracle.g1_fixed_base_mul(g1, mont(ec))
    return (a[0], int(ai[0])), (b[0], int(bi[0])), (c[0], int(ci[0]))


def synthetic_key_exponents(n: int, seed: int, num_public: int = 1, bits: int = 255):
    """Exponents of a ProvingKey-shaped key with n variables (the reference's dense QAP cannot reach
    2^20 constraints, so keys are synthesised as k_i * G): a, b1, b2 of n elements, ic of n - num_public - 1,
    h of n, plus the five single points."""
    k = {}
    for i, (name, cnt) in enumerate((("a_g1", n), ("b_g1", n), ("b_g2", n), ("ic_g1", n - num_public - 1), ("h_g1", n))):
        k[name] = oracle.gen_scalars(seed + 0x10 * i, cnt, bits)
    singles = oracle.gen_scalars(seed + 0x100, 5, bits)
    for name, v in zip(("alpha_g1", "beta_g1", "delta_g1", "beta_g2", "delta_g2"), singles):
        k[name] = v
    return k


def proofs_equal(p, q) -> bool:
    return all(pi == qi and (np.asarray(px) == np.asarray(qx)).all() for (px, pi), (qx, qi) in zip(p, q))
