"""Synthetic inputs on the device, following the reference's demo generator
(demo/demo_simulate_data.py:11-114) so that BASELINE.json's configs
c2-c5 never have to exist as CSV text or on the host.

Same distributions as the reference generator, not the same stream:
  A      exp(N(0,1)) with `anchor` anchor and `anti` anti-anchor genes zeroed per
         type, columns normalised                              (:11-42; host, K x N numbers)
  bulk   W_j ~ Dir(alpha), depth ~ Poisson(50 N), X_j ~ Mult(depth, A W_j)   (:45-69)
  cells  kappa_l ~ N(-1, .5), tau_l ~ N(1.5 N, .15 N), S ~ Bern(expit(kappa + tau A[:,G])),
         depth ~ Poisson(2 N), Y_l ~ Mult(depth, normalise(A[:,G_l] S_l))    (:72-114)
The counts are drawn by the library's own kernels (csrc/synth.cu: ursm_synth_bulk /
ursm_synth_cells; counter-based Philox keyed on the GLOBAL row, Poisson by inversion / PTRS,
Dirichlet by Marsaglia-Tsang gammas); torch only allocates the output tensors.  This module is
not on the timed path of bench.py.
"""
import numpy as np

from . import _lib


def profile_matrix(N, K, seed=0, anchor=2, anti=2):
    """A (N x K, float64, host) -- identical on every rank for a given seed."""
    rs = np.random.RandomState(seed)
    A = np.exp(rs.normal(size=(N, K)))
    for k in range(K):
        rows = np.arange(anchor * k, anchor * (k + 1))
        for other in range(K):
            if other != k:
                A[rows, other] = 0
    for k in range(K):
        A[np.arange(anchor * K + anti * k, anchor * K + anti * (k + 1)), k] = 0
    return A / A.sum(axis=0)


def cells_on_device(A, G, seed, row_offset=0, device="cuda", sc_depth=2.0, kappa=-1.0,
                    kappa_sd=0.5, tau_scale=1.5, tau_sd_frac=0.1):
    """fp32 L x N count matrix in HBM for the cells whose types are `G` (local rows;
    `row_offset` = global index of the first one); returns (Y, S_true, kappa_true, tau_true).
    Every cell gets at least one read (the reference divides by the depth, m_step.py:125)."""
    torch = _lib.require_cuda()
    A = _lib.f64(A)
    N, K = A.shape
    G = np.ascontiguousarray(np.asarray(G), dtype=np.int64)
    L = len(G)
    Y = torch.empty(L, N, dtype=torch.float32, device=device)
    S = torch.empty(L, N, dtype=torch.uint8, device=device)
    kap, tau = np.empty(L), np.empty(L)
    _lib.call("ursm_synth_cells", _lib.ptr(A), int(N), int(K), _lib.ptr(G), int(L), int(seed),
              int(row_offset), float(sc_depth), float(kappa), float(kappa_sd), float(tau_scale),
              float(tau_sd_frac), Y.data_ptr(), S.data_ptr(), _lib.ptr(kap), _lib.ptr(tau))
    return Y, S, kap, tau


def bulk_on_device(A, M, seed, row_offset=0, alpha=None, device="cuda", bk_depth=50.0):
    """fp32 M x N bulk count matrix in HBM; returns (X, W_true[M x K] on host)."""
    torch = _lib.require_cuda()
    A = _lib.f64(A)
    N, K = A.shape
    alpha = _lib.f64(np.full(K, 2.0) if alpha is None else np.asarray(alpha, dtype=float))
    X = torch.empty(int(M), N, dtype=torch.float32, device=device)
    W = np.empty((int(M), K))
    _lib.call("ursm_synth_bulk", _lib.ptr(A), int(N), int(K), int(M), _lib.ptr(alpha),
              float(bk_depth), int(seed), int(row_offset), X.data_ptr(), _lib.ptr(W))
    return X, W
