"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI export check, host-compiled codelet
checks (tests/hostcheck) and the world_size-2 gloo sharding test.  `-m gpu`: parity tests proper,
calling libavfe.so through avsl_b200 on a B200.
"""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def libavfe_path():
    from avsl_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def hostcheck():
    """TEST-ONLY: the library's __host__ __device__ codelets compiled for the CPU."""
    src = ROOT / "tests" / "hostcheck" / "hostcheck.cu"
    out_dir = ROOT / "tests" / "hostcheck" / "_build"
    out_dir.mkdir(exist_ok=True)
    so = out_dir / "libhostcheck.so"
    deps = [src] + list((ROOT / "avsl_b200" / "csrc").glob("*.cuh"))
    if not so.exists() or so.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        cmd = ["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr",
               "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", str(ROOT / "include"),
               "-I", str(ROOT / "avsl_b200" / "csrc"), "-o", str(so), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            pytest.fail("hostcheck build failed:\n" + res.stderr[-3000:])
    return ctypes.CDLL(os.fspath(so))


def vp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)
