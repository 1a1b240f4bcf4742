"""Native, multi-threaded replacements for the two numpy calls of scUnif.py's file
contract (scUnif.py:85-94): `np.savetxt(handle, a, delimiter=',')` and
`np.loadtxt(filename, delimiter=',')` for float matrices.  Byte-exact with numpy
(tests/test_csvio.py); SURVEY 8f row 1.  Host code from the C-ABI library
(csrc/csv_io.cpp) -- it does not need a GPU."""
import ctypes

import numpy as np

from . import _lib


def savetxt_csv(filename, array, nthreads=0):
    """np.savetxt(filename, array, delimiter=',') for real-valued 0/1/2-D arrays."""
    a = np.asarray(array, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:      # np.savetxt writes a 1-D array as one value per line
        a = a.reshape(-1, 1)
    elif a.ndim != 2:
        raise ValueError("Expected 1D or 2D array, got %dD array instead" % a.ndim)
    a = np.ascontiguousarray(a)
    _lib.call("ursm_csv_write_f64", str(filename).encode(), _lib.ptr(a), int(a.shape[0]),
              int(a.shape[1]), int(nthreads))


def loadtxt_csv(filename, nthreads=0):
    """np.loadtxt(filename, dtype=float, delimiter=',') (result squeezed as numpy does)."""
    rows, cols = ctypes.c_int64(), ctypes.c_int64()
    _lib.call("ursm_csv_shape", str(filename).encode(), ctypes.byref(rows), ctypes.byref(cols))
    out = np.empty((rows.value, cols.value), dtype=np.float64)
    _lib.call("ursm_csv_read_f64", str(filename).encode(), _lib.ptr(out), rows.value, cols.value,
              int(nthreads))
    if out.size == 0:
        return np.empty(0)
    return np.squeeze(out) if (rows.value == 1 or cols.value == 1) else out
