"""Parity tests proper (`-m gpu`): the CUDA engine, called through the C ABI, against the CPU oracle
on identical inputs -- bit-exact, as SURVEY.md 8 requires for integer work."""
import os

import numpy as np
import pytest

import helpers
import parity_cases as pc

pytestmark = pytest.mark.gpu


def test_library_is_the_cuda_build(gpu_ctx):
    import groth16_cuda
    assert os.path.samefile(groth16_cuda.DEFAULT_LIB, gpu_ctx.lib._name)
    assert b"sm_100a" in gpu_ctx.lib.g16_version()
    assert gpu_ctx.lib.g16_device_count() >= 1


def test_c_caller_through_the_abi(oracle, gens, tmp_path):
    """A caller written in C (tests/c_harness/abi_harness.c, built against include/g16_cuda.h): one-shot seam and the
    resident + precomputed path return the oracle's bytes."""
    import subprocess
    import groth16_cuda
    import test_abi_exports
    n = 3000
    pts, inf, sc = helpers.adversarial(oracle, gens, "g1", 0xc0ffee, n)
    exp, einf = oracle.g1_msm(pts, inf, sc, threads=oracle.max_threads())
    path = tmp_path / "msm_case.bin"
    with open(path, "wb") as f:
        f.write(np.uint64(n).tobytes()); f.write(pts.tobytes()); f.write(inf.tobytes()); f.write(sc.tobytes())
        f.write(exp.tobytes()); f.write(np.uint8(einf).tobytes())
    exe = test_abi_exports.build_c_harness(groth16_cuda.DEFAULT_LIB, tmp_path)
    out = subprocess.run([exe, "msm", str(path)], capture_output=True, text=True)
    assert out.returncode == 0 and "msm ok" in out.stdout, out.stdout + out.stderr


def test_field_ops_on_device(gpu_ctx, oracle):
    pc.check_debug_field(gpu_ctx, oracle, n=200000)


def test_group_add_exceptional_cases(gpu_ctx, oracle, gens):
    pc.check_debug_group_add(gpu_ctx, oracle, gens)


def test_golden_msm(gpu_ctx):
    pc.check_golden_msm(gpu_ctx)
    pc.check_empty(gpu_ctx)
    pc.check_length_mismatch(gpu_ctx)


def test_golden_fixed_base(gpu_ctx, gens):
    pc.check_golden_fixed_base(gpu_ctx, gens)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 1000])
def test_g1_small_sizes(gpu_ctx, oracle, gens, n):
    pc.check_random_msm(gpu_ctx, oracle, gens, "g1", n, n)


def test_g1_window_sweep(gpu_ctx, oracle, gens):
    pc.check_random_msm(gpu_ctx, oracle, gens, "g1", 3000, 3, windows=(0, 2, 3, 5, 8, 11, 13, 16, 20), pre=(8, 12, 16, 0))


def test_g1_2_16_config2(gpu_ctx, oracle, gens):
    """BASELINE config 2: standalone G1 MSM, 2^16 random scalars/points."""
    pc.check_random_msm(gpu_ctx, oracle, gens, "g1", 1 << 16, 16, pre=(0,))


@pytest.mark.parametrize("n", [1, 2, 3, 40, 1 << 12])
def test_g2_sizes(gpu_ctx, oracle, gens, n):
    pc.check_random_msm(gpu_ctx, oracle, gens, "g2", n, 100 + n, windows=(0, 7) if n == 40 else (0,), pre=(0,) if n >= 40 else ())


def test_adversarial_sets(gpu_ctx, oracle, gens):
    pc.check_adversarial(gpu_ctx, oracle, gens, "g1", 1 << 13, 5)
    pc.check_adversarial(gpu_ctx, oracle, gens, "g2", 1 << 10, 6)


def test_reference_faithful_scalar_distributions(gpu_ctx, oracle, gens):
    pc.check_skewed_scalars(gpu_ctx, oracle, gens, 1 << 13, 7)


def test_fixed_base(gpu_ctx, oracle, gens):
    pc.check_fixed_base_random(gpu_ctx, oracle, gens, 1 << 12, 9)


def test_device_resident_path_and_partials(gpu_ctx, oracle, gens):
    """Index-range shards -> projective partials -> combine == one-shot result (the multi-GPU data path,
    exercised on one device)."""
    import torch
    n, shards = 6000, 3
    pts, inf = helpers.make_points(oracle, gens, "g1", 0xabc, n)
    sc = oracle.gen_scalars(0xdef, n)
    exp, einf = oracle.g1_msm(pts, inf, sc, threads=oracle.max_threads())
    dev = torch.device("cuda:0")
    gpu_ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    d_sc = torch.from_numpy(sc.view(np.int64)).to(dev)
    partials = torch.zeros((shards, 48), dtype=torch.int32, device=dev)
    keep = []
    for k in range(shards):
        lo, hi = n * k // shards, n * (k + 1) // shards
        b = gpu_ctx.g1_bases_upload(pts[lo:hi], inf[lo:hi])
        keep.append(b)
        gpu_ctx.msm_device("g1", b, d_sc[lo:hi].data_ptr(), hi - lo, 0, partials[k].data_ptr())
    out = torch.zeros(25, dtype=torch.int32, device=dev)
    gpu_ctx.combine_partials_device("g1", partials.data_ptr(), shards, out.data_ptr())
    torch.cuda.synchronize()
    host = out.cpu().numpy().view(np.uint32)
    assert int(host[24]) == einf
    assert (host[:24].view(np.uint64) == exp).all()


def test_msm_sharded_orders_its_own_streams(oracle, gens):
    """groth16_cuda.dist.msm_sharded on a DEFAULT context (private non-blocking stream, no set_stream by the caller):
    MSM -> NCCL all-gather -> fold must be ordered by the function itself.  One-rank NCCL group, gather
    forced; repeated so that a missing dependency would show as a stale or partial result."""
    import torch
    import torch.distributed as dist
    import groth16_cuda
    from groth16_cuda.dist import msm_sharded
    dev = torch.device("cuda:0")
    created = False
    if not dist.is_initialized():
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
        created = True
    ctx = groth16_cuda.Context([0])
    try:
        n = 1 << 15
        pts, inf = helpers.make_points(oracle, gens, "g1", 0x51de, n)
        bases = ctx.g1_bases_upload(pts, inf)
        partial = torch.zeros(48, dtype=torch.int32, device=dev)
        gathered = torch.zeros(48, dtype=torch.int32, device=dev)
        out = torch.zeros(25, dtype=torch.int32, device=dev)
        for rep in range(4):
            sc = oracle.gen_scalars(0x51df + rep, n)
            exp, einf = oracle.g1_msm(pts, inf, sc, threads=oracle.max_threads())
            d_sc = torch.from_numpy(sc.view(np.int64)).to(dev)
            msm_sharded(ctx, "g1", bases, d_sc.data_ptr(), n, partial, gathered, out, 1, always_gather=True)
            host = out.cpu().numpy().view(np.uint32)      # .cpu() synchronises torch's current stream only
            assert int(host[24]) == einf and (host[:24].view(np.uint64) == exp).all(), rep
        bases.free()
    finally:
        ctx.close()
        if created:
            dist.destroy_process_group()


def test_multi_device_context_matches_single(gpu_ctx, oracle, gens):
    import groth16_cuda
    ndev = gpu_ctx.lib.g16_device_count()
    devs = list(range(min(ndev, 8)))
    if len(devs) < 2:
        devs = [0, 0]          # two shards on one device: same host-side sharding/combination code
    ctx = groth16_cuda.Context(devices=devs)
    try:
        pc.check_random_msm(ctx, oracle, gens, "g1", 5000, 41)
        pc.check_random_msm(ctx, oracle, gens, "g2", 300, 42)
        pc.check_fixed_base_random(ctx, oracle, gens, 512, 43)
    finally:
        ctx.close()


def _expected_by_discrete_log(oracle, gens, s, k):
    """(sum s_i k_i mod r) * G on the CPU: the exact value of sum s_i (k_i G)."""
    exp, einf = oracle.g1_fixed_base_mul(gens[0], oracle.dot_mod_r(s, k)[None])
    return exp[0], int(einf[0])


@pytest.mark.parametrize("log_n", [20, 22, 24])
def test_large_msm_by_discrete_log(gpu_ctx, oracle, gens, log_n):
    """Size-independent exact check at sizes the CPU oracle cannot reach in seconds: with bases
    P_i = k_i G (built on the GPU by the fixed-base kernel), sum s_i P_i must equal (sum s_i k_i mod r) G,
    which the oracle computes with one scalar multiplication.  log_n = 24 is BASELINE config 4 / the bench
    workload: same seeds, and the precomputed leg runs the exact bench plan (c = 22, 12 windows, one shared
    bucket set, two-pass partitioned scatter)."""
    import torch
    n = 1 << log_n
    dev = torch.device("cuda:0")
    gpu_ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    k = oracle.gen_scalars(0xba5e0000 + log_n, n)
    s = oracle.gen_scalars(0x5eed0000 + log_n, n)
    d_k = torch.from_numpy(k.view(np.int64)).to(dev)
    d_s = torch.from_numpy(s.view(np.int64)).to(dev)
    d_pts = torch.empty((n, 24), dtype=torch.int32, device=dev)
    gpu_ctx.fixed_base_mul_device("g1", gens[0], d_k.data_ptr(), n, d_pts.data_ptr())
    # spot-check the generated bases against the oracle
    torch.cuda.synchronize()
    del d_k
    idx = [0, 1, n // 2, n - 1]
    got = d_pts[idx].cpu().numpy().view(np.uint32).view(np.uint64)
    exp, _ = oracle.g1_fixed_base_mul(gens[0], k[idx])
    assert (got == exp).all()
    bases = gpu_ctx.bases_from_device("g1", d_pts.data_ptr(), n, keepalive=d_pts)
    out = torch.zeros(25, dtype=torch.int32, device=dev)

    def run(d_scalars):
        gpu_ctx.msm_device("g1", bases, d_scalars.data_ptr(), n, out.data_ptr(), 0)
        torch.cuda.synchronize()
        host = out.cpu().numpy().view(np.uint32)
        return host[:24].view(np.uint64).copy(), int(host[24])

    exp1 = _expected_by_discrete_log(oracle, gens, s, k)
    got1 = run(d_s)
    assert got1[1] == exp1[1] and (got1[0] == exp1[0]).all()
    # a second scalar vector (linearity through the same identity)
    s2 = oracle.gen_scalars(0x77 + log_n, n)
    d_s2 = torch.from_numpy(s2.view(np.int64)).to(dev)
    exp2 = _expected_by_discrete_log(oracle, gens, s2, k)
    got2 = run(d_s2)
    assert got2[1] == exp2[1] and (got2[0] == exp2[0]).all()
    # the same sums over precomputed multiples (one shared bucket set; from 2^22 on the entry array exceeds L2 and
    # goes through the two-pass partitioned scatter, as do the per-window bucket sets above)
    c = bases.precompute(0)
    if log_n == 24:
        assert c == 22, f"bench plan at 2^24 is c = 22, got {c}"
    for d_sc, exp in ((d_s, exp1), (d_s2, exp2)):
        got = run(d_sc)
        assert got[1] == exp[1] and (got[0] == exp[0]).all()
    # host scalars through the reference-facing call (pinned memory, chunked H2D): same point
    h_s = torch.from_numpy(s.view(np.int64)).pin_memory()
    gpu_ctx.msm_async("g1", bases, h_s.data_ptr(), n, out.data_ptr(), 0)
    torch.cuda.synchronize()
    host = out.cpu().numpy().view(np.uint32)
    assert int(host[24]) == exp1[1] and (host[:24].view(np.uint64) == exp1[0]).all()
    bases.free()


def test_chunked_host_path(gpu_ctx, oracle, gens):
    pc.check_chunked_host_path(gpu_ctx, oracle, gens, 5000, 21)


def test_wire_format(gpu_ctx, oracle, gens):
    """ark CanonicalSerialize / CanonicalDeserialize of G1Affine, G2Affine and Proof on the device."""
    import wire_cases as wc
    wc.check_known_answers(gpu_ctx)
    wc.check_golden(gpu_ctx)
    wc.check_roundtrip(gpu_ctx, "g1")
    wc.check_roundtrip(gpu_ctx, "g2")
    wc.check_rejects(gpu_ctx, "g1")
    wc.check_rejects(gpu_ctx, "g2")
    wc.check_proof(gpu_ctx)
    # batch: encode -> decode (with the subgroup check) is the identity on 2^13 CRS-style points, and the
    # compressed form of the first elements matches the big-integer encoder
    import bls12_381 as bls
    for group, n in (("g1", 1 << 13), ("g2", 1 << 11)):
        pts, inf = helpers.make_points(oracle, gens, group, 0x3e51a1, n)
        pts[5] = 0; inf[5] = 1
        for compressed in (True, False):
            data = gpu_ctx.serialize_points(group, pts, inf, compressed=compressed)
            xy, back_inf = gpu_ctx.deserialize_points(group, data, compressed=compressed, validate=True)
            assert (xy == pts).all() and (back_inf == inf).all()
        per = 48 if group == "g1" else 96
        data = gpu_ctx.serialize_points(group, pts[:8], inf[:8])
        for i in range(8):
            p = (bls.g1_from_mont if group == "g1" else bls.g2_from_mont)([int(v) for v in pts[i]], int(inf[i]))
            exp = bls.g1_compress(p) if group == "g1" else bls.g2_compress(p)
            assert data[i * per:(i + 1) * per] == exp


def test_g2_large_two_pass_scatter(gpu_ctx, oracle, gens):
    """G2 MSM whose entry array exceeds L2 (2^21 pairs x 13 windows): two-pass partitioned scatter in front of the G2
    accumulation, checked through the discrete-log identity (bases k_i G2 built by the fixed-base kernel)."""
    import bls12_381 as bls
    import torch
    n = 1 << 21
    dev = torch.device("cuda:0")
    gpu_ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    k = oracle.gen_scalars(0xba5e2200, n, 64)
    s = oracle.gen_scalars(0x5eed2200, n, 64)
    d_k = torch.from_numpy(k.view(np.int64)).to(dev)
    d_s = torch.from_numpy(s.view(np.int64)).to(dev)
    d_pts = torch.empty((n, 48), dtype=torch.int32, device=dev)
    gpu_ctx.fixed_base_mul_device("g2", gens[1], d_k.data_ptr(), n, d_pts.data_ptr())
    bases = gpu_ctx.bases_from_device("g2", d_pts.data_ptr(), n, keepalive=d_pts)
    out = torch.zeros(49, dtype=torch.int32, device=dev)
    gpu_ctx.set_window_bits(20)          # 13 windows -> 27 M entries (109 MB) even for 64-bit scalars' non-zero digits
    try:
        gpu_ctx.msm_device("g2", bases, d_s.data_ptr(), n, out.data_ptr(), 0)
        torch.cuda.synchronize()
    finally:
        gpu_ctx.set_window_bits(0)
    host = out.cpu().numpy().view(np.uint32)
    # 64-bit operands: the dot product fits numpy object arithmetic quickly
    e = int(sum(int(a) * int(b) for a, b in zip(oracle.fr_from_mont(s)[:, 0], oracle.fr_from_mont(k)[:, 0])) % bls.R)
    exp, einf = oracle.g2_fixed_base_mul(gens[1], np.array([bls.fr_to_mont(e)], dtype=np.uint64))
    assert int(host[48]) == int(einf[0]) and (host[:48].view(np.uint64) == exp[0]).all()
    bases.free()
