"""CPU-only: the C-ABI shared library loads and exports every symbol include/g16_cuda.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    sys.path.insert(0, ROOT)
    import __graft_entry__
    path = __graft_entry__.build_library()
    assert os.path.exists(path)
    return path


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "g16_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g16_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/g16_cuda.h but not exported"


def test_python_binding_lists_the_same_entry_points():
    sys.path.insert(0, os.path.join(ROOT, "zero-knowledge-proofs_b200"))
    import groth16_cuda
    syms = set(declared_symbols())
    assert set(groth16_cuda.EXPORTS) <= syms


def test_sass_is_sm_100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys.path.insert(0, os.path.join(ROOT, "zero-knowledge-proofs_b200"))
    import groth16_cuda
    with pytest.raises(groth16_cuda.MSMError) as e:
        groth16_cuda.Context()
    assert e.value.code == groth16_cuda.G16_ERR_NO_DEVICE
