#!/usr/bin/env python
"""Moment fixtures for the synthetic-data generators, produced by RUNNING the
reference's own generator functions (/root/reference/demo/demo_simulate_data.py:45-114,
`simulate_bulk` / `simulate_sc`; imported, nothing copied) on a fixed profile matrix:

    python tests/golden/make_synth_golden.py        # build container only (needs /root/reference)

tests/golden/stat_synth.npz then holds, per gene, the mean / variance of the bulk counts and, per
gene and cell type, the dropout rate and the mean single-cell count over many reference draws --
what `ursm_b200.synth` (device generator) and `oracle.simulate` are checked against."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    spec = importlib.util.spec_from_file_location(
        "ref_sim", "/root/reference/demo/demo_simulate_data.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from ursm_b200 import synth
    N, K, M, L = 200, 3, 3000, 3000
    A = synth.profile_matrix(N, K, seed=0, anchor=2, anti=2)   # host numpy, deterministic
    np.random.seed(20240601)
    alpha = np.full(K, 2.0)
    depths = np.random.poisson(50 * N, M)
    X, W = ref.simulate_bulk(N, M, K, alpha, A, depths)
    G = np.arange(L) % K
    tau = 1.5 * N
    sc_depths = np.random.poisson(2 * N, L)
    Y, S = ref.simulate_sc(N, L, K, G, A, sc_depths, tau, -1.0, tau * 0.1, 0.5)
    out = dict(A=A, N=N, K=K, M=M, L=L,
               bk_depth_mean=X.sum(axis=1).mean(), bk_depth_var=X.sum(axis=1).var(),
               bk_mean=X.mean(axis=0), bk_var=X.var(axis=0),
               sc_depth_mean=Y.sum(axis=1).mean(), sc_depth_var=Y.sum(axis=1).var(),
               sc_S_rate=np.stack([S[G == k].mean(axis=0) for k in range(K)]),
               sc_zero_rate=np.stack([(Y[G == k] == 0).mean(axis=0) for k in range(K)]),
               sc_mean=np.stack([Y[G == k].mean(axis=0) for k in range(K)]),
               sc_var=np.stack([Y[G == k].var(axis=0) for k in range(K)]),
               cells_per_type=np.array([(G == k).sum() for k in range(K)]))
    np.savez_compressed(os.path.join(HERE, "stat_synth.npz"), **out)
    print("wrote stat_synth.npz:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
