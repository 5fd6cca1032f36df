"""CPU-only, world_size = 2 over gloo: the N > 1 data path (index-range shards -> projective partials ->
all-gather -> fold) with the host-emulated kernels standing in for the GPU.  Checks that two ranks
reproduce the single-rank / oracle result bit for bit, for G1 and G2."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    root = sys.argv[1]
    for p in (root, os.path.join(root, "oracle"), os.path.join(root, "zero-knowledge-proofs_b200"), os.path.join(root, "tests"), os.path.join(root, "tests", "emu")):
        sys.path.insert(0, p)
    import bls12_381 as bls, cpu_oracle as oracle, groth16_cuda, build_emu, helpers
    from groth16_cuda.dist import shard_range, msm_sharded, PARTIAL_WORDS, AFFINE_WORDS
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = groth16_cuda.Context(lib_path=build_emu.build())       # emulation: "device" pointers are host pointers
    gens = (np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64), np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64))
    for group, n in (("g1", 301), ("g2", 45)):
        pts, inf, sc = helpers.adversarial(oracle, gens, group, 900 + n, n)          # same on every rank
        lo, hi = shard_range(n, rank, world)
        bases = (ctx.g1_bases_upload if group == "g1" else ctx.g2_bases_upload)(pts[lo:hi], inf[lo:hi])
        sc_loc = torch.from_numpy(np.ascontiguousarray(sc[lo:hi]).view(np.int64))
        partial = torch.zeros(PARTIAL_WORDS[group], dtype=torch.int32)
        gathered = torch.zeros(PARTIAL_WORDS[group] * world, dtype=torch.int32)
        out = torch.zeros(AFFINE_WORDS[group], dtype=torch.int32)
        msm_sharded(ctx, group, bases, sc_loc.data_ptr(), hi - lo, partial, gathered, out, world)
        exp, einf = (oracle.g1_msm if group == "g1" else oracle.g2_msm)(pts, inf, sc)
        got = out.numpy().view(np.uint32)
        assert int(got[-1]) == einf, (group, rank)
        assert (got[:-1].view(np.uint64) == exp).all(), (group, rank)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_ranks_reproduce_oracle(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    build_emu.build()
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                         capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("ok") == 2
