"""CPU-only: host check of the double-precision-limb multiplication scaffold (tools/dpf_mul_prototype.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dpf_multiplication_prototype(tmp_path):
    """tools/dpf_mul_prototype.cpp (next-round scaffold): Montgomery multiplication on 8 x 48-bit double limbs, limb products
    by fma round-toward-zero pairs, equals Fq::mul on 200 000 operand pairs."""
    exe = str(tmp_path / "dpf_mul_prototype")
    src = os.path.join(ROOT, "zero-knowledge-proofs_b200", "tools", "dpf_mul_prototype.cpp")
    inc = os.path.join(ROOT, "zero-knowledge-proofs_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-frounding-math", "-I", inc, "-o", exe, src], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "dpf mismatches: 0" in out.stdout, out.stdout + out.stderr
