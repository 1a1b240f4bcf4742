#!/usr/bin/env python
"""Repeatability diagnostic (run on a GPU box): three identical E-step + M-step passes on
the demo data with the same seed; prints short hashes of every output so that one sees
which quantities repeat bit for bit (the chains: exp_S, E kappa) and which only to
rounding (the atomically summed K x N statistics and what the M-step derives from them;
DESIGN.md sec. 7).

    python tools/det_check.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ursm_b200.e_step_gibbs as es  # noqa: E402
import ursm_b200.m_step as ms  # noqa: E402
from ursm_b200.gem import LogitNormalGEM  # noqa: E402


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "demo_data.npz"))
    Y, X, G = d["Y"], d["X"], d["G"]
    K, N = 3, Y.shape[1]
    itype = [np.where(G == k)[0] for k in range(K)]
    A = LogitNormalGEM(BKexpr=X, SCexpr=Y, K=K, G=G, burnin=1, sample=1).init_A  # gem.py:225-255
    pk, pt, al = np.array([-1., .01]), np.array([N, .01 * N], dtype=float), np.ones(K)
    h = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]
    for rep in range(3):
        sc = es.LogitNormalGibbs_SC(A, pk, pt, Y, G, itype, seed=5)
        sc.gibbs(10, 20, 1)
        st = sc.suff_stats
        bk = es.LogitNormalGibbs_BK(A, al, X, seed=5)
        bk.init_gibbs()
        bk.gibbs(mean_approx=True)
        allst = dict(st)
        allst.update(bk.suff_stats)
        mle = ms.LogitNormalMLE(X, Y, G, K, itype, True, True, A, al, pk, pt, 1e-6, 1, 1e-6, 500)
        mle.update_suff_stats(allst)
        e1 = mle.compute_elbo()
        mle.opt_kappa_tau()
        mle.opt_alpha()
        mle.opt_A_u()
        print(rep, "expS", h(np.asarray(st["exp_S"])), "kap", h(st["exp_kappa"]), "cA", h(st["coeffA"]),
              "cQ", h(st["coeffAsq"]), "elboc", repr(st["exp_elbo_const"]),
              "| W", h(bk.suff_stats["exp_W"]), "Zik", h(bk.suff_stats["exp_Zik"]),
              "| u", h(mle.u), "elbo", repr(e1), "A", h(mle.A), "niter", mle.niter_A,
              repr(mle.compute_elbo()))


if __name__ == "__main__":
    main()
