import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are skipped loudly, never silently passed, without a GPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


_HOST_SRC = os.path.join(ROOT, "tests", "host", "host_samplers.cpp")
_HOST_OUT = os.path.join(ROOT, "tests", "host", "_build", "libhost_samplers.so")
_HOST_HDRS = [os.path.join(ROOT, "ursm_b200", "csrc", h) for h in ("samplers.cuh", "rng.cuh")]


@pytest.fixture(scope="session")
def hostlib():
    """ursm_b200/csrc/samplers.cuh compiled for the HOST with g++ (test infrastructure)."""
    os.makedirs(os.path.dirname(_HOST_OUT), exist_ok=True)
    newest = max(os.path.getmtime(p) for p in [_HOST_SRC] + _HOST_HDRS)
    if not os.path.exists(_HOST_OUT) or os.path.getmtime(_HOST_OUT) < newest:
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", _HOST_OUT, _HOST_SRC])
    L = ctypes.CDLL(_HOST_OUT)
    dbl, i32, u32, u64, vp = (ctypes.c_double, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64,
                              ctypes.c_void_p)
    L.h_pg_mass.restype = dbl
    L.h_pg_mass.argtypes = [dbl]
    L.h_ndtri_sq.restype = dbl
    L.h_ndtri_sq.argtypes = [dbl]
    L.h_s_threshold.restype = dbl
    L.h_s_threshold.argtypes = [dbl] * 4
    L.h_pg_draws.argtypes = [dbl, i32, u64, vp]
    L.h_binomial.argtypes = [i32, dbl, i32, u64, vp]
    L.h_binomial_f32.argtypes = [i32, dbl, i32, u64, vp]
    L.h_gamma.argtypes = [dbl, i32, u64, vp]
    L.h_gene_uniforms.argtypes = [u64, u32, u32, i32, vp]
    L.h_cell_normals.argtypes = [u64, u32, u32, vp]
    L.h_init_bits.argtypes = [u64, u32, i32, vp]
    L.h_sc_key.restype = u64
    L.h_sc_key.argtypes = [u64, u64]
    return L
