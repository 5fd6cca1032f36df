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
    rc.check_duplicate_entries(gpu_ctx)
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


def test_prove_matches_cpu_five_msms_and_exponent_2_13(gpu_ctx, oracle, gens):
    """The prove schedule against BOTH checkers at a size the CPU model finishes in seconds."""
    prove_cases.check_prove_in_exponent(gpu_ctx, oracle, gens, 13, 0x13000, precompute=False, cpu_msms=True)
    prove_cases.check_prove_in_exponent(gpu_ctx, oracle, gens, 13, 0x13100, precompute=True, bits_list=(255,), cpu_msms=True)


def test_prove_config3_2_20_in_exponent(gpu_ctx, oracle, gens):
    """BASELINE config 3 at full size (2^20 variables / H coefficients, resident precomputed key): g16_prove checked
    exactly in the exponent -- every base is k_i * G with known k_i (crates/groth16-core/src/lib.rs:164-271)."""
    prove_cases.check_prove_in_exponent(gpu_ctx, oracle, gens, 20, 0x20000, precompute=True)


def test_prove_config3_2_20_multi_shard_in_exponent(oracle, gens):
    """The same on a sharded context (all GPUs of the box, or three index-range shards on one GPU)."""
    import groth16_cuda
    ndev = groth16_cuda.load_library().g16_device_count()
    ctx = groth16_cuda.Context(devices=list(range(min(ndev, 8))) if ndev > 1 else [0, 0, 0])
    try:
        prove_cases.check_prove_in_exponent(ctx, oracle, gens, 20, 0x20300, precompute=False, bits_list=(255,))
    finally:
        ctx.close()
