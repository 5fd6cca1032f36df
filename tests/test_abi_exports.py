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


def test_rust_crate_binds_every_entry_point():
    """rust/groth16-cuda/src/sys.rs (not compilable here: no Rust toolchain) declares exactly the header's functions,
    and the safe layer uses every one of them except the raw handles' destructors it wraps in Drop."""
    crate = os.path.join(ROOT, "zero-knowledge-proofs_b200", "rust", "groth16-cuda", "src")
    sys_rs = open(os.path.join(crate, "sys.rs")).read()
    lib_rs = open(os.path.join(crate, "lib.rs")).read()
    bound = sorted(set(re.findall(r"\bfn\s+(g16_[a-z0-9_]+)\s*\(", sys_rs)))
    assert bound == declared_symbols()
    for name in bound:
        assert re.search(r"\b" + name + r"\s*\(", lib_rs), f"{name} is declared in sys.rs but never called from lib.rs"
    assert "impl rand_core::RngCore for FixedLimbsRng" in lib_rs


def build_c_harness(lib_path, tmp_path):
    """gcc -std=c99 -pedantic -Werror: the header is valid C and every declared function links."""
    exe = str(tmp_path / "abi_harness")
    src = os.path.join(ROOT, "tests", "c_harness", "abi_harness.c")
    libdir = os.path.dirname(lib_path)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src,
                    "-L", libdir, "-lg16cuda", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_header_compiles_as_c_and_links(lib_path, tmp_path):
    exe = build_c_harness(lib_path, tmp_path)
    out = subprocess.run([exe, "symbols"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"entry points: {len(declared_symbols())}" in out.stdout and "sm_100a" in out.stdout


def test_c_caller_fails_loudly_without_gpu(lib_path, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([build_c_harness(lib_path, tmp_path), "nodevice"], capture_output=True, text=True)
    assert out.returncode == 0 and "rc=3" in out.stdout, out.stdout + out.stderr


def test_version_carries_the_source_digest(lib_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("g16_build", os.path.join(ROOT, "zero-knowledge-proofs_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(lib_path)
    lib.g16_version.restype = ctypes.c_char_p
    assert ("src:" + mod.source_hash()).encode() in lib.g16_version()


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
