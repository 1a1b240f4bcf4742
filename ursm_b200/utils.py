"""Helper functions with the reference's names (utils.py:8-38)."""
import numpy as np

from . import _lib


def simplex_proj(y, min_y):
    """Project `y` onto {y >= min_y, sum(y) = 1}  (utils.py:8-31).

    Runs the same sort-free device routine the M-step kernel uses
    (csrc/mle.cu::proj_lambda); the projection is unique, so the result equals
    the reference's sort-based one up to summation-order rounding.
    """
    _lib.require_cuda()
    y = _lib.f64(np.asarray(y).ravel())
    out = np.empty_like(y)
    _lib.call("ursm_simplex_proj", _lib.ptr(y), int(y.size), float(min_y), _lib.ptr(out))
    return out


def simplex_proj_columns(B, min_y):
    """`simplex_proj` of every column of `B` (N x K) in one launch (gem.py:252-254)."""
    _lib.require_cuda()
    Bt = _lib.f64(np.asarray(B, dtype=float).T)            # K x N, rows contiguous
    out = np.empty_like(Bt)
    _lib.call("ursm_simplex_proj_rows", _lib.ptr(Bt), int(Bt.shape[0]), int(Bt.shape[1]),
              float(min_y), _lib.ptr(out))
    return np.ascontiguousarray(out.T)


def std_row(B):
    """Rows scaled to sum to one (utils.py:34-38)."""
    return np.true_divide(B, B.sum(axis=1)[:, np.newaxis])
