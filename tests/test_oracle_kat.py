"""Pins the oracle (CPU-only).  The reference has no golden vectors for the MSM path,
so the pins are: public BLS12-381 known answers, fixtures from the exact big-int model
(tests/golden/make_golden.py), and agreement between the two independent CPU implementations
(Python big-int, naive double-and-add  vs  C port of ark's msm_bigint_wnaf)."""
import numpy as np
import pytest

import helpers


def test_public_known_answers(bls):
    kat = helpers.load_json("kat.json")
    # widely published encodings of the generators (Zcash/IETF format used by ark-bls12-381 0.4.0)
    assert kat["g1_gen_compressed"].startswith("97f1d3a73197d7942695638c4fa9ac0f")
    assert kat["g2_gen_compressed"].startswith("93e02b6052719f607dacd3a088274f65")
    assert bls.g1_compress(bls.G1_GEN).hex() == kat["g1_gen_compressed"]
    assert bls.g2_compress(bls.G2_GEN).hex() == kat["g2_gen_compressed"]
    assert f"{bls.G1.mul(bls.G1_GEN, 2)[0]:096x}" == kat["g1_2g_x"]
    assert bls.G1.on_curve(bls.G1_GEN) and bls.G2.on_curve(bls.G2_GEN)
    assert bls.G1.mul(bls.G1_GEN, bls.R) is None and bls.G2.mul(bls.G2_GEN, bls.R) is None
    assert bls.g1_compress(None)[0] == 0xC0 and bls.g1_decompress(bls.g1_compress(bls.G1_GEN)) == bls.G1_GEN
    # Montgomery constants of ark-ff
    assert f"{bls.FQ_R:096x}" == kat["fq_R"] and f"{bls.FQ_NINV64:016x}" == "89f3fffcfffcfffd"
    assert f"{bls.FR_R:064x}" == kat["fr_R"] and f"{bls.FR_NINV64:016x}" == "fffffffeffffffff"
    assert len(bls.proof_bytes(bls.G1_GEN, bls.G2_GEN, None)) == 192
    assert len(bls.proof_bytes(bls.G1_GEN, bls.G2_GEN, None, compressed=False)) == 384


def test_c_field_matches_bigint(oracle, bls):
    rng = np.random.default_rng(1)
    vals = [0, 1, bls.Q - 1, bls.FQ_R, bls.Q >> 1] + [int.from_bytes(rng.bytes(48), "little") % bls.Q for _ in range(200)]
    a = np.array([bls.int_to_limbs64(v, 6) for v in vals], dtype=np.uint64)
    b = a[::-1].copy()
    for name, op in (("fq_mul", lambda x, y: x * y * bls.FQ_RINV), ("fq_add", lambda x, y: x + y), ("fq_sub", lambda x, y: x - y)):
        r = getattr(oracle, name)(a, b)
        for i, (x, y) in enumerate(zip(vals, vals[::-1])):
            assert bls.limbs64_to_int(r[i]) == op(x, y) % bls.Q, name
    inv = oracle.fq_inv(a[1:])
    for i, x in enumerate(vals[1:]):
        assert bls.limbs64_to_int(inv[i]) == pow(x * bls.FQ_RINV, -1, bls.Q) * bls.FQ_R % bls.Q
    s = np.array([bls.fr_to_mont(v % bls.R) for v in vals], dtype=np.uint64)
    back = oracle.fr_from_mont(s)
    for i, v in enumerate(vals):
        assert bls.limbs64_to_int(back[i]) == v % bls.R


def test_ark_window_and_digits(oracle, bls):
    # c = 3 if n < 32 else ceil(log2 n) * 69 / 100 + 2   
    assert [oracle.msm_window(n) for n in (1, 31, 32, 1 << 16, 1 << 20, 1 << 24, 1 << 26)] == [3, 3, 5, 13, 15, 18, 19]
    rng = bls.SplitMix64(3)
    for w in (3, 5, 13, 16):
        for _ in range(20):
            k = bls.random_fr(rng)
            d = oracle.make_digits(bls.int_to_limbs64(k, 4), w)
            assert sum(int(x) << (w * i) for i, x in enumerate(d)) == k
            assert all(-(1 << (w - 1)) <= int(x) < (1 << (w - 1)) for x in d[:-1])


@pytest.mark.parametrize("case", helpers.golden_msm_cases(), ids=lambda c: c["name"])
def test_c_msm_matches_golden(oracle, bls, case):
    fn = oracle.g1_msm if case["group"] == "g1" else oracle.g2_msm
    naive = oracle.g1_msm_naive if case["group"] == "g1" else oracle.g2_msm_naive
    for f in (fn, naive):
        out, inf = f(case["points"], case["inf"], case["scalars"])
        assert inf == case["result_inf"]
        assert (out == case["result"]).all()
    if case["group"] == "g1":
        assert bls.g1_compress(bls.g1_from_mont(list(case["result"]), case["result_inf"])).hex() == case["compressed"]
    else:
        assert bls.g2_compress(bls.g2_from_mont(list(case["result"]), case["result_inf"])).hex() == case["compressed"]


def test_c_fixed_base_matches_golden(oracle, gens):
    fb = helpers.load_json("fixed_base_cases.json")
    sc = helpers.limbs(fb["scalars"])
    out, inf = oracle.g1_fixed_base_mul(gens[0], sc)
    assert (out == helpers.limbs(fb["g1"])).all() and list(inf) == fb["g1_inf"]
    out, inf = oracle.g2_fixed_base_mul(gens[1], sc)
    assert (out == helpers.limbs(fb["g2"])).all() and list(inf) == fb["g2_inf"]


def test_c_pippenger_vs_python_pippenger(oracle, bls, gens):
    n = 300
    pts, inf = helpers.make_points(oracle, gens, "g1", 0xba5e0000, n)
    sc = oracle.gen_scalars(0x5eed0000, n)
    out, oinf = oracle.g1_msm(pts, inf, sc)
    py_pts = [bls.g1_from_mont(list(p), i) for p, i in zip(pts, inf)]
    py_sc = [bls.fr_from_mont(list(s)) for s in sc]
    assert bls.g1_from_mont(list(out), oinf) == bls.G1.msm_pippenger(py_pts, py_sc)
    out_t, _ = oracle.g1_msm(pts, inf, sc, threads=4)   # windows in parallel: same element
    assert (out_t == out).all()


def test_c_adversarial_vs_naive(oracle, gens):
    for group in ("g1", "g2"):
        n = 400 if group == "g1" else 120
        pts, inf, sc = helpers.adversarial(oracle, gens, group, 77, n)
        f, naive = (oracle.g1_msm, oracle.g1_msm_naive) if group == "g1" else (oracle.g2_msm, oracle.g2_msm_naive)
        a, ai = f(pts, inf, sc)
        b, bi = naive(pts, inf, sc)
        assert ai == bi and (a == b).all()


def test_wire_format_oracle(bls):
    """oracle/bls12_381.py encoder / decoder against the committed fixture and the public generator encodings."""
    fx = helpers.load_json("wire_cases.json")
    kat = helpers.load_json("kat.json")
    assert fx["g1"][1]["compressed"] == kat["g1_gen_compressed"] and fx["g2"][1]["compressed"] == kat["g2_gen_compressed"]
    for group, from_mont, enc_c, enc_u in (("g1", bls.g1_from_mont, bls.g1_compress, bls.g1_uncompressed),
                                           ("g2", bls.g2_from_mont, bls.g2_compress, bls.g2_uncompressed)):
        for c in fx[group]:
            p = from_mont([int(v, 16) for v in c["limbs"]], c["inf"])
            assert enc_c(p).hex() == c["compressed"] and enc_u(p).hex() == c["uncompressed"]
            assert bls.point_deserialize(group, bytes.fromhex(c["compressed"]), True) == p
            assert bls.point_deserialize(group, bytes.fromhex(c["uncompressed"]), False) == p
    # rejections: wrong compression flag, non-canonical x, x without a point, a curve point outside the subgroup
    g = bytes.fromhex(kat["g1_gen_compressed"])
    with pytest.raises(bls.WireError, match="UnexpectedFlags"):
        bls.point_deserialize("g1", g, compressed=False)
    with pytest.raises(bls.WireError, match="InvalidData"):
        bls.point_deserialize("g1", bytes([0x9f]) + b"\xff" * 47)
    rejected = accepted_unchecked = 0
    x = int.from_bytes(bytes([g[0] & 0x1f]) + g[1:], "big")
    for d in range(1, 12):
        b = (x + d).to_bytes(48, "big")
        b = bytes([b[0] | 0x80]) + b[1:]
        try:
            bls.point_deserialize("g1", b, validate=True)
        except bls.WireError:
            rejected += 1
        try:
            p = bls.point_deserialize("g1", b, validate=False)
            assert bls.G1.on_curve(p)
            accepted_unchecked += 1
        except bls.WireError:
            pass
    assert rejected == 11 and 0 < accepted_unchecked < 11     # on the curve about half the time, never in the subgroup
    r2 = bls.fq2_sqrt((3, 5))
    assert r2 is None or bls.Fq2Ops.sqr(r2) == (3, 5)
    sq = bls.Fq2Ops.sqr((123456789, 987654321))
    r2 = bls.fq2_sqrt(sq)
    assert bls.Fq2Ops.sqr(r2) == sq
