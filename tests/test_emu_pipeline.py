"""CPU-side (no GPU) check of the engine's pipeline logic: the kernel bodies of
zero-knowledge-proofs_b200/csrc are compiled for the host by tests/emu/build_emu.py and driven through
the same C ABI.  This validates digit decomposition, the counting sort, bucket accumulation with all
exceptional group cases, the multi-level bucket reduction, the window fold and the fixed-base tables
before any GPU time is spent.  The emulation library is test infrastructure; the product never loads it."""
import parity_cases as pc


def test_emu_golden_msm(emu_ctx):
    pc.check_golden_msm(emu_ctx)
    pc.check_empty(emu_ctx)
    pc.check_length_mismatch(emu_ctx)


def test_emu_golden_fixed_base(emu_ctx, gens):
    pc.check_golden_fixed_base(emu_ctx, gens)


def test_emu_field_and_group_hooks(emu_ctx, oracle, gens):
    pc.check_debug_field(emu_ctx, oracle, n=300)
    pc.check_debug_group_add(emu_ctx, oracle, gens)


def test_emu_random_and_window_sweep(emu_ctx, oracle, gens):
    pc.check_random_msm(emu_ctx, oracle, gens, "g1", 257, 1, windows=(0, 2, 3, 7, 11, 16), pre=(8, 13, 0))
    pc.check_random_msm(emu_ctx, oracle, gens, "g2", 40, 2, windows=(0, 5), pre=(9,))


def test_emu_adversarial_and_skewed(emu_ctx, oracle, gens):
    pc.check_adversarial(emu_ctx, oracle, gens, "g1", 300, 5)
    pc.check_adversarial(emu_ctx, oracle, gens, "g2", 60, 6)
    pc.check_skewed_scalars(emu_ctx, oracle, gens, 200, 7)
    pc.check_skewed_scalars(emu_ctx, oracle, gens, 1500, 8)   # > ITEM_MAX entries per bucket: split + merge path


def test_emu_chunked_host_path(emu_ctx, oracle, gens):
    pc.check_chunked_host_path(emu_ctx, oracle, gens, 400, 21)


def test_emu_fixed_base_groups(emu_ctx, oracle, gens):
    # n not a multiple of the inversion group, zero scalars inside a group, another base than the generator
    pc.check_fixed_base_random(emu_ctx, oracle, gens, 203, 9)


def test_emu_quotient_polynomial(emu_ctx):
    import quotient_cases as qc
    qc.check_toy_circuits(emu_ctx)
    qc.check_random_polynomials(emu_ctx, log_sizes=(0, 1, 2, 3, 5))


def test_emu_r1cs_evaluations(emu_ctx):
    import r1cs_cases as rc
    rc.check_evals_small(emu_ctx)
    rc.check_duplicate_entries(emu_ctx)
    rc.check_evals_long_lines(emu_ctx, m=300)


def test_emu_r1cs_setup_and_prove(emu_ctx):
    import r1cs_cases as rc
    rc.check_setup_and_prove(emu_ctx, circuits=rc.CIRCUITS[:3])
    rc.check_errors(emu_ctx)


def test_emu_wire_format(emu_ctx):
    import wire_cases as wc
    wc.check_known_answers(emu_ctx)
    wc.check_golden(emu_ctx)
    wc.check_roundtrip(emu_ctx, "g1")
    wc.check_roundtrip(emu_ctx, "g2", ks=wc.KS[:6])
    wc.check_rejects(emu_ctx, "g1")
    wc.check_rejects(emu_ctx, "g2")
    wc.check_proof(emu_ctx)


def test_emu_prove_multi_device_schedule(emu_ctx, oracle, gens):
    """g16_prove on a multi-device context (index-range shards of every ProvingKey array, per-device slices of the
    assignment, partial sums folded on device 0): config-1 proofs, and a random key against the one-device schedule."""
    import json
    import os
    import groth16_cuda
    import prove_cases
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1_proofs.json")))
    for devs in ([0, 1], [0, 1, 2, 3, 4]):
        ctx = groth16_cuda.Context(devices=devs, lib_path=emu_ctx.lib._name)
        try:
            prove_cases.check_config1(ctx, golden)
            prove_cases.check_random_key(ctx, emu_ctx, oracle, gens, n=23, seed=len(devs))
        finally:
            ctx.close()


def test_emu_prove_with_scalar_width_hint(emu_ctx, oracle, gens):
    """g16_pk_precompute_bits: tables chosen for 64-bit scalars (what the reference's truncation yields), proofs checked in
    the exponent -- including one full-width assignment entry and the full-width r, s that break the promise on purpose."""
    import prove_cases
    prove_cases.check_prove_in_exponent(emu_ctx, oracle, gens, log_n=9, seed=0x51, precompute=True, bits_list=(64, 255))
