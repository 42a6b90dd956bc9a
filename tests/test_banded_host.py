"""The pivoted banded LU the Newton kernels fall back to (csrc/sb_banded.cuh) is __host__ __device__: exercise it on the
CPU against dense Gaussian elimination with partial pivoting on random block-tridiagonal systems that are NOT diagonally
dominant (the case the fast, pivot-free path cannot handle; reference: pivoted LU, src/utils.jl:317-319)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_banded_lu_with_row_interchanges(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "banded_host")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "native", "banded_host.cu")], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == 45
    for line in out:
        P, N, flavour, info, res, err = line.split()
        assert int(info) == 0, line
        assert float(res) < 1e-12, line                  # backward stable: tiny relative residual whatever the conditioning
