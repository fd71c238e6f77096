"""The C ABI used from a plain C++/CUDA host program (no torch, no Python buffers): cudaMalloc'ed memory, a
user stream, checked against the C oracle.  Compiled on the box with nvcc and run as a subprocess."""
import os
import subprocess

import pytest

from oracle import c_oracle
from relation_detr_b200 import build

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_program_against_oracle(tmp_path):
    lib = build.build()
    oracle_so = c_oracle.build()
    exe = str(tmp_path / "c_abi_smoke")
    cmd = [build.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
           "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi", "c_abi_smoke.cu"),
           "-o", exe, lib, oracle_so,
           "-Xlinker", f"-rpath={os.path.dirname(lib)}", "-Xlinker", f"-rpath={os.path.dirname(oracle_so)}", "-Xcompiler", "-fopenmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "C ABI OK" in run.stdout, run.stdout
