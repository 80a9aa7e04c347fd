"""The padded position layout of csrc/common.cuh (Geo) checked on the host: nvcc compiles a small host-only program
against the header the kernels use and runs it on the CPU.  Property: for every pixel and 3x3 tap, position + dy*Wp + dx is
the neighbouring pixel's position when it exists and a zero-halo position otherwise (this is what lets the implicit-GEMM
kernels skip every bounds test, /root/reference/src/train_brain.jl:111-140 `pad=1`), pos/decode are inverse, and the halo'ed
slab of the last tile stays inside the guard zone."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_geo_layout_invariants(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "layout_check")
    src = os.path.join(ROOT, "tests", "host_src", "layout_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-O1", "-o", exe, src], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    lines = res.stdout.strip().splitlines()
    assert len(lines) == 4 and all(l.endswith("bad=0") for l in lines), res.stdout
    # one zero column per row: 32x32 -> 33x33 positions per image (+ one closing separator row)
    assert "Wp=33 Hs=33" in lines[0] and "Wp=17 Hs=17" in lines[2], res.stdout
