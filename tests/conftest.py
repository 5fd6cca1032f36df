import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "zero-knowledge-proofs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (C port of ark's msm_bigint_wnaf + Python big-int model)."""
    import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle


@pytest.fixture(scope="session")
def bls():
    import bls12_381
    return bls12_381


@pytest.fixture(scope="session")
def gens(bls):
    g1 = np.array(bls.g1_to_mont(bls.G1_GEN)[0], dtype=np.uint64)
    g2 = np.array(bls.g2_to_mont(bls.G2_GEN)[0], dtype=np.uint64)
    return g1, g2


@pytest.fixture(scope="session")
def gpu_ctx():
    """Context on the real CUDA library.  Fails loudly (no fallback) if the library or GPU is missing."""
    import groth16_cuda
    ctx = groth16_cuda.Context()
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def emu_ctx():
    """Context on the host-emulation build (tests/emu) -- CPU-side check of the pipeline logic."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    import groth16_cuda
    path = build_emu.build()
    ctx = groth16_cuda.Context(lib_path=path)
    yield ctx
    ctx.close()
