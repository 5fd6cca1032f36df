"""Wire-format checks shared by the emulation tests and the GPU tests: the engine's batch encoders / decoders
(csrc/wire_kernels.cuh, through the C ABI) against the big-integer statement of ark-bls12-381's encoding in
oracle/bls12_381.py, plus the published encodings of the generators (tests/golden/kat.json)."""
import numpy as np
import pytest

import bls12_381 as bls
import helpers


def _points(group, ks):
    curve, gen = (bls.G1, bls.G1_GEN) if group == "g1" else (bls.G2, bls.G2_GEN)
    return [None if k == 0 else curve.mul(gen, k) for k in ks]


def _to_limbs(group, pts):
    to = bls.g1_to_mont if group == "g1" else bls.g2_to_mont
    w = 12 if group == "g1" else 24
    xy = np.zeros((len(pts), w), dtype=np.uint64)
    inf = np.zeros(len(pts), dtype=np.uint8)
    for i, p in enumerate(pts):
        if p is None:
            inf[i] = 1
        else:
            xy[i] = to(p)[0]
    return xy, inf


def _enc(group, p, compressed):
    if group == "g1":
        return bls.g1_compress(p) if compressed else bls.g1_uncompressed(p)
    return bls.g2_compress(p) if compressed else bls.g2_uncompressed(p)


KS = [0, 1, 2, 3, 5, bls.R - 1, bls.R - 2, 0xdeadbeef, (1 << 200) + 12345, 0, 77]


def check_known_answers(ctx):
    kat = helpers.load_json("kat.json")
    xy, inf = _to_limbs("g1", [bls.G1_GEN])
    assert ctx.serialize_points("g1", xy, inf).hex() == kat["g1_gen_compressed"]
    xy, inf = _to_limbs("g2", [bls.G2_GEN])
    assert ctx.serialize_points("g2", xy, inf).hex() == kat["g2_gen_compressed"]


def check_roundtrip(ctx, group, ks=KS):
    pts = _points(group, ks)
    xy, inf = _to_limbs(group, pts)
    for compressed in (True, False):
        exp = b"".join(_enc(group, p, compressed) for p in pts)
        got = ctx.serialize_points(group, xy, inf, compressed=compressed)
        assert got == exp, (group, compressed)
        # the oracle's decoder agrees with its encoder ...
        per = len(exp) // len(pts)
        for i, p in enumerate(pts):
            assert bls.point_deserialize(group, exp[i * per:(i + 1) * per], compressed) == p
        # ... and the engine decodes to ark's in-memory limbs
        for validate in (True, False):
            back_xy, back_inf = ctx.deserialize_points(group, got, compressed=compressed, validate=validate)
            assert (back_inf == inf).all() and (back_xy == xy).all(), (group, compressed, validate)
    assert ctx.serialize_points(group, xy[:0], inf[:0]) == b""
    e_xy, e_inf = ctx.deserialize_points(group, b"")
    assert e_xy.shape[0] == 0


def check_rejects(ctx, group):
    """every rejection ark's reader makes: flag mismatch, non-canonical coordinate, x without a point, point outside
    the subgroup, and (uncompressed, validate) a point off the curve -- element by element against the oracle."""
    import groth16_cuda
    p = _points(group, [7])[0]
    k = 1 if group == "g1" else 2
    good_c, good_u = _enc(group, p, True), _enc(group, p, False)
    cases_c = [good_c]
    cases_c.append(bytes([good_c[0] & 0x7f]) + good_c[1:])                      # compressed flag missing
    cases_c.append(bytes([0x9f]) + b"\xff" * (48 * k - 1))                      # coordinate >= q
    cases_c.append(bytes([good_c[0] ^ 0x20]) + good_c[1:])                      # other root: -P, still valid
    # x values around P.x: some have no point, some are on the curve but outside the subgroup
    x_int = int.from_bytes(bytes([good_c[0] & 0x1f]) + good_c[1:], "big")
    for d in range(1, 9):
        b = (x_int + d).to_bytes(48 * k, "big")
        cases_c.append(bytes([b[0] | 0x80]) + b[1:])
    cases_c.append(bytes([0xc0]) + bytes(48 * k - 1))                           # identity
    cases_c.append(bytes([0xc0]) + b"\x01" * (48 * k - 1))                      # identity flag wins (ark returns zero)
    cases_u = [good_u, bytes([good_u[0] | 0x80]) + good_u[1:], bytes([0x40]) + bytes(96 * k - 1)]
    off = bytearray(good_u); off[-1] ^= 1
    cases_u.append(bytes(off))                                                  # y changed: off the curve
    for compressed, cases in ((True, cases_c), (False, cases_u)):
        for validate in (True, False):
            exp_status, exp_pts = [], []
            for c in cases:
                try:
                    exp_pts.append(bls.point_deserialize(group, c, compressed, validate))
                    exp_status.append(0)
                except bls.WireError as e:
                    exp_pts.append(None)
                    exp_status.append(1 if e.kind == "InvalidData" else 2)
            xy, inf, status = ctx.deserialize_points(group, b"".join(cases), compressed=compressed, validate=validate,
                                                     return_status=True)
            assert list(status) == exp_status, (group, compressed, validate, list(status), exp_status)
            exp_xy, exp_inf = _to_limbs(group, exp_pts)
            assert (inf == exp_inf).all() and (xy == exp_xy).all(), (group, compressed, validate)
            if any(exp_status):
                with pytest.raises(groth16_cuda.MSMError) as ei:
                    ctx.deserialize_points(group, b"".join(cases), compressed=compressed, validate=validate)
                first = next(s for s in exp_status if s)
                assert ("InvalidData" if first == 1 else "UnexpectedFlags") in str(ei.value)
        assert 1 in exp_status or not compressed


def check_proof(ctx):
    """Proof = a || b || c (crates/groth16-core/src/lib.rs:27-36): the committed config-1 proofs."""
    import groth16_cuda
    a, c = _points("g1", [11, 0xabcdef])
    b = _points("g2", [13])[0]
    la, ia = _to_limbs("g1", [a]); lb, ib = _to_limbs("g2", [b]); lc, ic = _to_limbs("g1", [c])
    for compressed in (True, False):
        exp = bls.proof_bytes(a, b, c, compressed)
        got = ctx.proof_serialize((la[0], ia[0]), (lb[0], ib[0]), (lc[0], ic[0]), compressed=compressed)
        assert got == exp and len(got) == (192 if compressed else 384)
        (ra, rai), (rb, rbi), (rc, rci) = ctx.proof_deserialize(got, compressed=compressed)
        assert (ra == la[0]).all() and (rb == lb[0]).all() and (rc == lc[0]).all() and (rai, rbi, rci) == (0, 0, 0)
        bad = bytearray(got); bad[len(got) // 4 + 1] ^= 0x55       # inside b
        with pytest.raises(groth16_cuda.MSMError) as ei:
            ctx.proof_deserialize(bytes(bad), compressed=compressed)
        assert "proof.b" in str(ei.value)
    # identity members
    got = ctx.proof_serialize((la[0] * 0, 1), (lb[0], 0), (lc[0] * 0, 1))
    assert got == bls.proof_bytes(None, b, None)
    with pytest.raises(groth16_cuda.MSMError):
        ctx.proof_deserialize(got[:100])


def check_golden(ctx):
    """committed fixture (tests/golden/wire_cases.json, made by make_wire_golden.py): bytes <-> ark limbs."""
    fx = helpers.load_json("wire_cases.json")
    for group in ("g1", "g2"):
        xy = helpers.limbs([c["limbs"] for c in fx[group]])
        inf = np.array([c["inf"] for c in fx[group]], dtype=np.uint8)
        for mode, compressed in (("compressed", True), ("uncompressed", False)):
            data = bytes.fromhex("".join(c[mode] for c in fx[group]))
            assert ctx.serialize_points(group, xy, inf, compressed=compressed) == data
            bxy, binf = ctx.deserialize_points(group, data, compressed=compressed)
            assert (bxy == xy).all() and (binf == inf).all()
