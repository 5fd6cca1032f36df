"""BASELINE config 1 on the GPU: toy circuits, fixed SetupParams, fixed (r, s) -> proof bytes identical
to the reference-semantics model and to the committed golden fixture; verdict from the model's verifier."""
import json
import os

import pytest

import prove_cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_proofs.json")


def test_config1_proofs_bit_exact(gpu_ctx):
    prove_cases.check_config1(gpu_ctx, json.load(open(GOLDEN)))


def test_config1_proofs_multi_shard_context():
    import groth16_cuda
    ctx = groth16_cuda.Context(devices=[0, 0])     # two index-range shards (one GPU is enough for the code path)
    try:
        prove_cases.check_config1(ctx, json.load(open(GOLDEN)))
    finally:
        ctx.close()


def test_quotient_polynomial_ntt(gpu_ctx):
    import quotient_cases as qc
    qc.check_toy_circuits(gpu_ctx)
    qc.check_random_polynomials(gpu_ctx, log_sizes=(0, 1, 2, 3, 5, 8, 10))


def test_r1cs_evaluations(gpu_ctx):
    import r1cs_cases as rc
    rc.check_evals_small(gpu_ctx)
    rc.check_evals_long_lines(gpu_ctx, m=700)


def test_r1cs_setup_and_prove(gpu_ctx):
    import r1cs_cases as rc
    golden = os.path.join(os.path.dirname(GOLDEN), "r1cs_proofs.json")
    rc.check_setup_and_prove(gpu_ctx, golden=json.load(open(golden)))
    rc.check_errors(gpu_ctx)


def test_r1cs_large_polynomial_identity(gpu_ctx, oracle):
    import r1cs_cases as rc
    rc.check_large_properties(gpu_ctx, oracle, log_m=12)


def test_prove_multi_device_random_key(gpu_ctx, oracle, gens):
    """Device-chained prove schedule over index-range shards (all GPUs of the box, or two shards on one GPU) against
    the one-device schedule on a random ProvingKey: lengths that do not divide evenly, arrays of different lengths,
    zero scalars, with and without H."""
    import groth16_cuda
    ndev = gpu_ctx.lib.g16_device_count()
    # five shards: the fold of the partial sums takes the warp-cooperative path (k >= 4) in G1 and G2
    for devs in ([0, 0], [0, 0, 0, 0, 0], list(range(min(ndev, 8))) if ndev > 1 else [0, 0, 0]):
        ctx = groth16_cuda.Context(devices=devs)
        try:
            prove_cases.check_random_key(ctx, gpu_ctx, oracle, gens, n=3001, seed=len(devs))
        finally:
            ctx.close()
