"""CPU-only: the wide product / separate reduction path of csrc/fp.cuh (`Fp::mul_wide`, `Fp::redc_wide`) and the lazy
Fq2 multiplication built on it (`Fq2::mul_lazy`, compiled into the kernels only with -DG16_FQ2_LAZY: measured slower)
agree with the fused CIOS multiplication and the Karatsuba form on random and edge operands."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fq2_lazy_matches_karatsuba(tmp_path):
    exe = str(tmp_path / "fq2_lazy_check")
    src = os.path.join(ROOT, "tests", "emu", "fq2_lazy_check.cpp")
    inc = os.path.join(ROOT, "zero-knowledge-proofs_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "mismatches: 0" in out.stdout, out.stdout + out.stderr


def test_dpf_multiplication_prototype(tmp_path):
    """tools/dpf_mul_prototype.cpp (next-round scaffold): Montgomery multiplication on 8 x 48-bit double limbs, limb products
    by fma round-toward-zero pairs, equals Fq::mul on 200 000 operand pairs."""
    exe = str(tmp_path / "dpf_mul_prototype")
    src = os.path.join(ROOT, "zero-knowledge-proofs_b200", "tools", "dpf_mul_prototype.cpp")
    inc = os.path.join(ROOT, "zero-knowledge-proofs_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-frounding-math", "-I", inc, "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "dpf mismatches: 0" in out.stdout, out.stdout + out.stderr
