"""Synthetic inputs on the device, following the reference's demo generator
(demo/demo_simulate_data.py:11-114) so that BASELINE.json's configs
c2-c5 never have to exist as CSV text or on the host.

Same distributions as the reference generator, not the same stream:
  A      exp(N(0,1)) with `anchor` anchor and `anti` anti-anchor genes zeroed per
         type, columns normalised                              (:11-42)
  bulk   W_j ~ Dir(alpha), depth ~ Poisson(50 N), X_j ~ Mult(depth, A W_j)   (:45-69)
  cells  kappa_l ~ N(-1, .5), tau_l ~ N(1.5 N, .15 N), S ~ Bern(expit(kappa + tau A[:,G])),
         depth ~ Poisson(2 N), Y_l ~ Mult(depth, normalise(A[:,G_l] S_l))    (:72-114)
A multinomial with a Poisson total is a vector of independent Poissons, which
is how the counts are drawn here (exactly the same joint law, one pass).

torch is used as plumbing (device RNG + elementwise ops); this module is not
on the timed path of bench.py.
"""
import numpy as np


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
                    kappa_sd=0.5, tau_scale=1.5, tau_sd_frac=0.1, chunk_elems=1 << 26):
    """fp32 L x N count matrix in HBM for the cells whose types are `G` (local rows);
    returns (Y, S_true, kappa_true, tau_true).  Every cell gets at least one read
    (the reference divides by the depth, m_step.py:125)."""
    import torch
    N, K = A.shape
    L = len(G)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1000003 + int(row_offset) + 17)
    At = torch.as_tensor(A.T.copy(), dtype=torch.float32, device=device)  # K x N
    Gd = torch.as_tensor(np.asarray(G), dtype=torch.long, device=device)
    Y = torch.empty(L, N, dtype=torch.float32, device=device)
    S = torch.empty(L, N, dtype=torch.uint8, device=device)
    kap = kappa + kappa_sd * torch.randn(L, generator=gen, device=device)
    tau = tau_scale * N * (1 + tau_sd_frac * torch.randn(L, generator=gen, device=device))
    step = max(1, chunk_elems // N)
    for r0 in range(0, L, step):
        r1 = min(L, r0 + step)
        a = At[Gd[r0:r1]]                                            # rows x N
        prob = torch.sigmoid(kap[r0:r1, None] + tau[r0:r1, None] * a)
        s = torch.rand(r1 - r0, N, generator=gen, device=device) < prob
        p = a * s
        tot = p.sum(dim=1, keepdim=True)
        dead = (tot <= 0).squeeze(1)
        if dead.any():  # no expressed gene at all: fall back to the profile itself
            p[dead] = a[dead]
            tot = p.sum(dim=1, keepdim=True)
        y = torch.poisson(p / tot * (sc_depth * N), generator=gen)
        empty = y.sum(dim=1) == 0
        if empty.any():
            idx = torch.argmax(p[empty], dim=1)
            y[empty.nonzero().squeeze(1), idx] = 1.0
        Y[r0:r1] = y
        S[r0:r1] = s
    return Y, S, kap.double().cpu().numpy(), tau.double().cpu().numpy()


def bulk_on_device(A, M, seed, row_offset=0, alpha=None, device="cuda", bk_depth=50.0):
    """fp32 M x N bulk count matrix in HBM; returns (X, W_true[M x K] on host)."""
    import torch
    N, K = A.shape
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 1000003 + int(row_offset) + 7919)
    alpha = np.full(K, 2.0) if alpha is None else np.asarray(alpha, dtype=float)
    rs = np.random.RandomState((int(seed) * 7 + int(row_offset)) % (2 ** 31))
    W = rs.dirichlet(alpha, M)
    At = torch.as_tensor(A.T.copy(), dtype=torch.float32, device=device)
    Wd = torch.as_tensor(W, dtype=torch.float32, device=device)
    X = torch.empty(M, N, dtype=torch.float32, device=device)
    step = max(1, (1 << 26) // N)
    for r0 in range(0, M, step):
        r1 = min(M, r0 + step)
        X[r0:r1] = torch.poisson(Wd[r0:r1] @ At * (bk_depth * N), generator=gen)
    return X, W
