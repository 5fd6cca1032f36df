"""Parity checks shared by the GPU tests (real library, -m gpu) and the CPU-side emulation tests.
Every check calls the engine through the C ABI (groth16_cuda.Context) and compares bit-for-bit with
the oracle on the same inputs."""
import ctypes

import numpy as np

import helpers


def check_golden_msm(ctx):
    for case in helpers.golden_msm_cases():
        f = ctx.multi_scalar_mult_g1 if case["group"] == "g1" else ctx.multi_scalar_mult_g2
        out, inf = f(case["scalars"], case["points"], case["inf"])
        assert inf == case["result_inf"], case["name"]
        assert (out == case["result"]).all(), case["name"]


def check_golden_fixed_base(ctx, gens):
    fb = helpers.load_json("fixed_base_cases.json")
    sc = helpers.limbs(fb["scalars"])
    out, inf = ctx.fixed_base_mul_g1(gens[0], sc)
    assert (out == helpers.limbs(fb["g1"])).all() and list(inf) == fb["g1_inf"]
    out, inf = ctx.fixed_base_mul_g2(gens[1], sc)
    assert (out == helpers.limbs(fb["g2"])).all() and list(inf) == fb["g2_inf"]


def check_empty(ctx):
    # Prover::multi_scalar_mult_g1 returns the identity for an empty list (lib.rs:276-278)
    out, inf = ctx.multi_scalar_mult_g1(np.zeros((0, 4), np.uint64), np.zeros((0, 12), np.uint64))
    assert inf == 1 and not out.any()
    out, inf = ctx.multi_scalar_mult_g2(np.zeros((0, 4), np.uint64), np.zeros((0, 24), np.uint64))
    assert inf == 1 and not out.any()


def check_random_msm(ctx, oracle, gens, group, n, seed, windows=(0,), pre=()):
    pts, inf = helpers.make_points(oracle, gens, group, 0xba5e0000 + seed, n)
    sc = oracle.gen_scalars(0x5eed0000 + seed, n)
    exp, einf = (oracle.g1_msm if group == "g1" else oracle.g2_msm)(pts, inf, sc, threads=oracle.max_threads())
    bases = ctx.g1_bases_upload(pts, inf) if group == "g1" else ctx.g2_bases_upload(pts, inf)
    for c in windows:
        ctx.set_window_bits(c)
        out, oinf = (ctx.g1_msm if group == "g1" else ctx.g2_msm)(bases, sc)
        assert oinf == einf and (out == exp).all(), (group, n, c)
    ctx.set_window_bits(0)
    # resident bases with precomputed multiples 2^(c w) P: same group element
    for pc_bits in pre:
        used = bases.precompute(pc_bits)
        assert used == (pc_bits or used) and 8 <= used <= 24
        out, oinf = (ctx.g1_msm if group == "g1" else ctx.g2_msm)(bases, sc)
        assert oinf == einf and (out == exp).all(), (group, n, "precompute", pc_bits)
    # prefix of the resident bases (n' < len(bases))
    m = max(1, n // 3)
    exp, einf = (oracle.g1_msm if group == "g1" else oracle.g2_msm)(pts[:m], inf[:m], sc[:m])
    out, oinf = (ctx.g1_msm if group == "g1" else ctx.g2_msm)(bases, sc[:m])
    assert oinf == einf and (out == exp).all()
    bases.free()


def check_adversarial(ctx, oracle, gens, group, n, seed):
    pts, inf, sc = helpers.adversarial(oracle, gens, group, seed, n)
    exp, einf = (oracle.g1_msm if group == "g1" else oracle.g2_msm)(pts, inf, sc, threads=oracle.max_threads())
    f = ctx.multi_scalar_mult_g1 if group == "g1" else ctx.multi_scalar_mult_g2
    out, oinf = f(sc, pts, inf)
    assert oinf == einf and (out == exp).all()


def check_skewed_scalars(ctx, oracle, gens, n, seed):
    """reference-faithful distributions: 64-bit truncated scalars, booleans, one repeated value."""
    import bls12_381 as bls
    pts, inf = helpers.make_points(oracle, gens, "g1", seed, n, bits=64)   # CRS-like k*G, small k
    rng = np.random.default_rng(seed)
    one = np.array(bls.fr_to_mont(1), dtype=np.uint64)
    for kind in ("u64", "bool", "const"):
        if kind == "u64":
            sc = oracle.gen_scalars(seed + 9, n, bits=64)
        elif kind == "bool":
            sc = np.zeros((n, 4), dtype=np.uint64)
            sc[rng.random(n) < 0.5] = one
        else:
            # every scalar equal: each window piles all n points into ONE bucket (exercises bucket splitting)
            sc = np.tile(np.array(bls.fr_to_mont(0x1234567), dtype=np.uint64), (n, 1))
        exp, einf = oracle.g1_msm(pts, inf, sc, threads=oracle.max_threads())
        out, oinf = ctx.multi_scalar_mult_g1(sc, pts, inf)
        assert oinf == einf and (out == exp).all(), kind


def check_length_mismatch(ctx):
    import groth16_cuda
    try:
        ctx.multi_scalar_mult_g1(np.zeros((3, 4), np.uint64), np.zeros((2, 12), np.uint64))
    except groth16_cuda.MSMError as e:
        assert e.code == groth16_cuda.G16_ERR_LENGTH
    else:
        raise AssertionError("length mismatch must raise MSMError (ark: Err(min_len))")


def check_fixed_base_random(ctx, oracle, gens, n, seed):
    for bits in (64, 255):
        sc = oracle.gen_scalars(seed + bits, n, bits)
        sc[0] = 0
        out, inf = ctx.fixed_base_mul_g1(gens[0], sc)
        exp, einf = oracle.g1_fixed_base_mul(gens[0], sc, threads=oracle.max_threads())
        assert (out == exp).all() and (inf == einf).all()
    m = max(1, n // 4)
    sc = oracle.gen_scalars(seed + 1, m)
    out, inf = ctx.fixed_base_mul_g2(gens[1], sc)
    exp, einf = oracle.g2_fixed_base_mul(gens[1], sc, threads=oracle.max_threads())
    assert (out == exp).all() and (inf == einf).all()
    # a base other than the generator
    base, _ = oracle.g1_fixed_base_mul(gens[0], oracle.gen_scalars(seed + 2, 1))
    out, inf = ctx.fixed_base_mul_g1(base[0], sc)
    exp, einf = oracle.g1_fixed_base_mul(base[0], sc, threads=oracle.max_threads())
    assert (out == exp).all() and (inf == einf).all()


def check_debug_field(ctx, oracle, n=2000):
    import bls12_381 as bls
    lib = ctx.lib
    lib.g16_debug_fq_op.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_size_t]
    rng = np.random.default_rng(5)
    edge = [0, 1, bls.Q - 1, bls.Q - 2, bls.FQ_R, 2, 1 << 380, bls.Q >> 1]
    vals = edge + [int.from_bytes(rng.bytes(48), "little") % bls.Q for _ in range(n)]
    a = np.array([bls.int_to_limbs64(v, 6) for v in vals], dtype=np.uint64)
    b = np.ascontiguousarray(a[::-1])
    # all edge x edge pairs first
    ea = np.array([bls.int_to_limbs64(x, 6) for x in edge for _ in edge], dtype=np.uint64)
    eb = np.array([bls.int_to_limbs64(y, 6) for _ in edge for y in edge], dtype=np.uint64)
    a = np.concatenate([ea, a]); b = np.concatenate([eb, b])
    out = np.zeros_like(a)
    for op, ref in ((0, oracle.fq_mul), (1, oracle.fq_add), (2, oracle.fq_sub)):
        assert lib.g16_debug_fq_op(ctx.handle, op, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.shape[0]) == 0
        assert (out == ref(a, b)).all(), op
    assert lib.g16_debug_fq_op(ctx.handle, 4, a.ctypes.data, None, out.ctypes.data, a.shape[0]) == 0
    assert (out == oracle.fq_mul(a, a)).all()
    assert lib.g16_debug_fq_op(ctx.handle, 5, a.ctypes.data, None, out.ctypes.data, a.shape[0]) == 0
    assert (out == oracle.fq_sub(np.zeros_like(a), a)).all()
    # two products, one reduction (Fp::mul_dual): (x y + x^2 (x + y)) R^-1
    assert lib.g16_debug_fq_op(ctx.handle, 6, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.shape[0]) == 0
    assert (out == oracle.fq_add(oracle.fq_mul(a, b), oracle.fq_mul(oracle.fq_mul(a, a), oracle.fq_add(a, b)))).all()
    xy, x2s = oracle.fq_mul(a, b), oracle.fq_mul(oracle.fq_mul(a, a), oracle.fq_add(a, b))
    assert lib.g16_debug_fq_op(ctx.handle, 7, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.shape[0]) == 0   # four products
    assert (out == oracle.fq_sub(oracle.fq_add(oracle.fq_add(xy, x2s), oracle.fq_mul(b, b)), oracle.fq_mul(a, a))).all()
    assert lib.g16_debug_fq_op(ctx.handle, 8, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.shape[0]) == 0   # difference of products
    assert (out == oracle.fq_sub(xy, x2s)).all()
    nz = a[a.any(axis=1)][:64]
    outi = np.zeros_like(nz)
    assert lib.g16_debug_fq_op(ctx.handle, 3, nz.ctypes.data, None, outi.ctypes.data, nz.shape[0]) == 0
    assert (outi == oracle.fq_inv(nz)).all()
    lib.g16_debug_fr_from_mont.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    s = oracle.gen_scalars(3, 500)
    s[0] = 0
    so = np.zeros_like(s)
    assert lib.g16_debug_fr_from_mont(ctx.handle, s.ctypes.data, so.ctypes.data, s.shape[0]) == 0
    assert (so == oracle.fr_from_mont(s)).all()


def check_debug_group_add(ctx, oracle, gens):
    """P + Q through the XYZZ mixed addition for: generic, P = Q, P = -Q, P = O, Q = O, both O."""
    lib = ctx.lib
    for g, width, msm in (("g1", 12, oracle.g1_msm), ("g2", 24, oracle.g2_msm)):
        fn = getattr(lib, f"g16_debug_{g}_add")
        fn.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_size_t]
        n = 24
        p, pinf = helpers.make_points(oracle, gens, g, 31, n)
        q, qinf = helpers.make_points(oracle, gens, g, 32, n)
        half = width // 2
        q[0:4] = p[0:4]                                                   # doubling
        q[4:8] = p[4:8]
        q[4:8, half:] = oracle.fq_sub(np.zeros((4 * half // 6, 6), np.uint64), p[4:8, half:].reshape(-1, 6)).reshape(4, half)  # P + (-P)
        p[8:10] = 0; pinf[8:10] = 1                                       # O + Q
        q[10:12] = 0; qinf[10:12] = 1                                     # P + O
        p[12] = 0; pinf[12] = 1; q[12] = 0; qinf[12] = 1                  # O + O
        out = np.zeros((n, width), dtype=np.uint64)
        oinf = np.zeros(n, dtype=np.uint8)
        rc = fn(ctx.handle, p.ctypes.data, pinf.ctypes.data, q.ctypes.data, qinf.ctypes.data, out.ctypes.data,
                oinf.ctypes.data, n)
        assert rc == 0
        import bls12_381 as bls
        one = np.array([bls.fr_to_mont(1)] * 2, dtype=np.uint64)
        for i in range(n):
            exp, einf = msm(np.stack([p[i], q[i]]), np.array([pinf[i], qinf[i]], dtype=np.uint8), one)
            assert einf == oinf[i] and (exp == out[i]).all(), (g, i)


def check_chunked_host_path(ctx, oracle, gens, n, seed):
    """Host-scalar MSMs above the pipeline threshold copy their scalars in three index ranges; every range is sorted
    and accumulated INTO THE SAME bucket array, which is reduced once (engine.cuh, H2D_PIPE_PARTS): same group
    element, for plain and for precomputed resident bases, G1 and G2, including a prefix of the bases, infinities,
    duplicates and skewed scalars whose buckets are split into slices (the add-to path of the slice merge)."""
    ctx.set_h2d_pipeline_min(64)
    try:
        for group, m in (("g1", n), ("g2", max(70, n // 4))):
            pts, inf, sc = helpers.adversarial(oracle, gens, group, seed, m)
            msm = oracle.g1_msm if group == "g1" else oracle.g2_msm
            exp, einf = msm(pts, inf, sc, threads=oracle.max_threads())
            bases = ctx.g1_bases_upload(pts, inf) if group == "g1" else ctx.g2_bases_upload(pts, inf)
            f = ctx.g1_msm if group == "g1" else ctx.g2_msm
            out, oinf = f(bases, sc)
            assert oinf == einf and (out == exp).all(), (group, "plain")
            bases.precompute(9)
            out, oinf = f(bases, sc)
            assert oinf == einf and (out == exp).all(), (group, "precomputed")
            k = m - 7
            exp, einf = msm(pts[:k], inf[:k], sc[:k])
            out, oinf = f(bases, sc[:k])
            assert oinf == einf and (out == exp).all(), (group, "prefix")
            bases.free()
        check_skewed_scalars(ctx, oracle, gens, max(200, n // 2), seed + 3)
        check_skewed_scalars(ctx, oracle, gens, 1500, seed + 4)      # > ITEM_MAX entries per bucket in every range
    finally:
        ctx.set_h2d_pipeline_min(0)   # back to the default
