#!/usr/bin/env python
"""`make simulate`: the demo inputs of URSM (demo/demo_simulate_data.py:117-150 of the
reference: N = 100 genes, K = 3 types, 50 single cells of types 5/5/40, 50 bulk samples,
alpha = (1,2,3), 3 anchor and 2 anti-anchor genes per type), drawn on the device by
ursm_b200.synth -- the same distributions, not the same stream -- and written under
demo/demo_data/ with the reference's file names and CSV layout.

    python demo/demo_simulate_data.py [--seed 0] [--N 100 --L 50 --M 50]
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--N", type=int, default=100)
    ap.add_argument("--L", type=int, default=50)
    ap.add_argument("--M", type=int, default=50)
    ap.add_argument("--outdir", default=os.path.join(HERE, "demo_data"))
    args = ap.parse_args()
    from ursm_b200 import _lib, csvio, synth
    _lib.require_cuda()
    K, N, L, M = 3, args.N, args.L, args.M
    A = synth.profile_matrix(N, K, seed=args.seed, anchor=3, anti=2)
    n_small = max(1, L // 10)
    G = np.array([0] * n_small + [1] * n_small + [2] * (L - 2 * n_small))
    Y, S, _, _ = synth.cells_on_device(A, G, args.seed)
    X, W = synth.bulk_on_device(A, M, args.seed, alpha=np.arange(1, K + 1))
    os.makedirs(args.outdir, exist_ok=True)
    out = lambda name: os.path.join(args.outdir, name)
    csvio.savetxt_csv(out("demo_bulk_rnaseq_counts.csv"), X.double().cpu().numpy())
    csvio.savetxt_csv(out("demo_bulk_mix_proportions.csv"), W)
    csvio.savetxt_csv(out("demo_single_cell_rnaseq_counts.csv"), Y.double().cpu().numpy())
    csvio.savetxt_csv(out("demo_single_cell_dropout_status.csv"), S.double().cpu().numpy())
    csvio.savetxt_csv(out("demo_profile_matrix.csv"), A)
    np.savetxt(out("demo_single_cell_types.csv"), G, fmt="%d")
    print("wrote %d cells, %d bulk samples, %d genes to %s" % (L, M, N, args.outdir))


if __name__ == "__main__":
    main()
