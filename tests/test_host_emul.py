"""CPU: the per-thread work functions of the two-sided Jacobi kernel (csrc/jacobi2_core.h) compiled for the host and
scheduled exactly as the CUDA block schedules them (tests/host_emul/jacobi2_host.cpp): relabelling covers every pair
once per sweep; residual / orthogonality at fp64 and fp32 level; odd, tiny and rank-deficient sizes."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_jacobi2_host_emulation(tmp_path):
    exe = str(tmp_path / "jacobi2_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(HERE, "host_emul", "jacobi2_host.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "all ok" in out.stdout, out.stdout + out.stderr
