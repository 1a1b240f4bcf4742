"""ctypes binding of the C-ABI CUDA library (include/ursm_b200.h).

There is no CPU fallback: if the library is missing or a call fails, this module
raises.  The library is built in-tree by `python -m ursm_b200.build`
(`__graft_entry__.build()`), never JIT-compiled at import time.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libursm_b200.so")

c_int, c_i32, c_i64, c_u64 = ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
c_dbl, c_vp = ctypes.c_double, ctypes.c_void_p
PP = ctypes.POINTER(c_vp)

# name -> (restype, argtypes).  Kept in one table so tests can check that every
# symbol the header declares is exported and bound.
SIGNATURES = {
    "ursm_last_error": (ctypes.c_char_p, []),
    "ursm_version": (c_int, []),
    "ursm_launch_count": (c_i64, []),
    "ursm_launch_count_reset": (None, []),
    "ursm_host_alloc": (c_int, [PP, c_i64]),
    "ursm_host_free": (c_int, [c_vp]),
    "ursm_host_trim": (c_int, []),
    "ursm_sc_create": (c_int, [PP, c_vp, c_vp, c_i64, c_i32, c_i32, c_i64]),
    "ursm_sc_create_dev": (c_int, [PP, c_vp, c_vp, c_i64, c_i32, c_i32, c_i64]),
    "ursm_sc_destroy": (c_int, [c_vp]),
    "ursm_sc_stats_len": (c_i64, [c_vp]),
    "ursm_sc_set_params": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ursm_sc_gibbs": (c_int, [c_vp, c_i32, c_i32, c_i32, c_u64, c_u64, c_vp, c_vp]),
    "ursm_sc_get_cell_moments": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ursm_sc_get_exp_S": (c_int, [c_vp, c_vp, c_i64, c_i64]),
    "ursm_sc_inject": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "ursm_sc_sweep_debug": (c_int, [c_vp, c_vp, c_vp, c_vp, c_u64, c_u64, c_i32, c_vp, c_vp,
                                    c_vp, c_vp, c_vp]),
    "ursm_sc_set_counts": (c_int, [c_vp, c_vp, c_i32, c_vp]),
    "ursm_sc_chain_debug": (c_int, [c_vp, c_vp, c_vp, c_vp, c_u64, c_u64, c_i32, c_i32, c_i32,
                                    c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ursm_sc_kernel_ms": (c_int, [c_vp, c_vp]),
    "ursm_bk_kernel_ms": (c_int, [c_vp, c_vp]),
    "ursm_sc_geometry": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ursm_bk_create": (c_int, [PP, c_vp, c_i64, c_i32, c_i32, c_i64]),
    "ursm_bk_create_dev": (c_int, [PP, c_vp, c_i64, c_i32, c_i32, c_i64]),
    "ursm_bk_destroy": (c_int, [c_vp]),
    "ursm_bk_stats_len": (c_i64, [c_vp]),
    "ursm_bk_set_params": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_bk_set_state": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_bk_mean_step": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_bk_gibbs": (c_int, [c_vp, c_i32, c_i32, c_i32, c_u64, c_u64, c_i32, c_vp, c_vp]),
    "ursm_bk_get_sample_stats": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ursm_bk_get_W": (c_int, [c_vp, c_vp]),
    "ursm_sc_type_sums": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_bk_profile_sum": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_mle_create": (c_int, [PP, c_vp, c_vp, c_i32, c_i32, c_dbl]),
    "ursm_mle_destroy": (c_int, [c_vp]),
    "ursm_mle_set_A": (c_int, [c_vp, c_vp]),
    "ursm_mle_get_A": (c_int, [c_vp, c_vp]),
    "ursm_mle_begin": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_mle_pass_u": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ursm_mle_set_coeffs": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ursm_mle_opt_begin": (c_int, [c_vp, c_vp, c_dbl, c_dbl, c_i32, c_vp]),
    "ursm_mle_opt_step": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_mle_opt_pass": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_mle_opt_merge": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ursm_mle_opt_poll": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ursm_mle_opt_run": (c_int, [c_vp, c_vp, c_dbl, c_dbl, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ursm_mle_p2p_alloc": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_mle_p2p_state": (c_int, [c_vp, c_vp]),
    "ursm_mle_p2p_reset": (c_int, [c_vp]),
    "ursm_mle_p2p_disable": (c_int, [c_vp]),
    "ursm_mle_p2p_open": (c_int, [c_vp, c_vp, c_i32, c_i32]),
    "ursm_mle_opt_pass_p2p": (c_int, [c_vp, c_vp]),
    "ursm_mle_opt_merge_p2p": (c_int, [c_vp, c_vp, c_vp]),
    "ursm_mle_p2p_error": (c_int, [c_vp, c_vp]),
    "ursm_mle_closed_form": (c_int, [c_vp, c_i32, c_vp, c_vp]),
    "ursm_mle_elbo_terms": (c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ursm_mle_get_u": (c_int, [c_vp, c_vp]),
    "ursm_simplex_proj": (c_int, [c_vp, c_i32, c_dbl, c_vp]),
    "ursm_simplex_proj_rows": (c_int, [c_vp, c_i32, c_i32, c_dbl, c_vp]),
    "ursm_synth_bulk": (c_int, [c_vp, c_i32, c_i32, c_i64, c_vp, c_dbl, c_u64, c_i64, c_vp, c_vp]),
    "ursm_synth_cells": (c_int, [c_vp, c_i32, c_i32, c_vp, c_i64, c_u64, c_i64, c_dbl, c_dbl, c_dbl,
                                 c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp]),
    "ursm_csv_write_f64": (c_int, [ctypes.c_char_p, c_vp, c_i64, c_i64, c_i32]),
    "ursm_csv_shape": (c_int, [ctypes.c_char_p, c_vp, c_vp]),
    "ursm_csv_read_f64": (c_int, [ctypes.c_char_p, c_vp, c_i64, c_i64, c_i32]),
}

_lib = None


class UrsmError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes library; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UrsmError(
            "CUDA library %s not found: run `python -m ursm_b200.build` (needs nvcc). "
            "ursm_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ursm_last_error()
        raise UrsmError("%s failed (%d): %s" % (what or "ursm call", rc,
                                               msg.decode() if msg else "?"))


def call(name, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)


def f64(a):
    """C-contiguous float64 view/copy + its pointer holder."""
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a):
    """void* of a numpy array (None -> NULL)."""
    return None if a is None else c_vp(a.__array_interface__["data"][0])


class PinnedBlock(object):
    """A page-locked host block from the library's pool, exposed to numpy through the
    array interface: `np.asarray(block)` is a view that keeps the block alive, and the
    block goes back to the pool when the last view is gone."""

    def __init__(self, shape, dtype=np.float64):
        self.shape = tuple(int(x) for x in shape)
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = c_vp()
        call("ursm_host_alloc", ctypes.byref(p), max(nbytes, 1))
        self._p = p.value
        self.__array_interface__ = {"data": (self._p, False), "shape": self.shape,
                                    "typestr": self.dtype.str, "version": 3}

    def __del__(self):
        p, self._p = getattr(self, "_p", None), None
        if p:
            try:
                load().ursm_host_free(c_vp(p))
            except Exception:
                pass


# read-backs up to this size land in pooled page-locked memory; larger ones use
# ordinary numpy memory (chunked copies)
PINNED_READBACK_MAX = 8 << 30


def host_result(shape, dtype=np.float64):
    """Uninitialised host array for a device read-back."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if 0 < nbytes <= PINNED_READBACK_MAX:
        try:
            return np.asarray(PinnedBlock(shape, dtype))
        except UrsmError:
            pass
    return np.empty(shape, dtype=dtype)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise UrsmError("ursm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return c_vp(torch.cuda.current_stream().cuda_stream)
