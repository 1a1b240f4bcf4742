#!/usr/bin/env python
"""`make check`: what demo/demo_plots.py of the reference shows as two scatter plots
(estimated against true profile matrix A and mixing proportions W), as numbers: Pearson
correlation and mean absolute error of both, with a non-zero exit status when the
simulation truth is not recovered.

    python demo/demo_check.py [--data demo/demo_data] [--out demo/demo_out]
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=os.path.join(HERE, "demo_data"))
    ap.add_argument("--out", default=os.path.join(HERE, "demo_out"))
    ap.add_argument("--min-corr-A", type=float, default=0.9)
    ap.add_argument("--min-corr-W", type=float, default=0.9)
    args = ap.parse_args()
    load = lambda d, f: np.loadtxt(os.path.join(d, f), dtype=float, delimiter=",")
    true_A = load(args.data, "demo_profile_matrix.csv")
    true_W = load(args.data, "demo_bulk_mix_proportions.csv")
    est_A = load(args.out, "gemout_est_A.csv")
    est_W = load(args.out, "gemout_exp_W.csv").transpose()  # written K x M (scUnif.py:80)
    ok = True
    for name, t, e, lim in (("A", true_A, est_A, args.min_corr_A), ("W", true_W, est_W, args.min_corr_W)):
        c = float(np.corrcoef(t.ravel(), e.ravel())[0, 1])
        mae = float(np.abs(t - e).mean())
        print("%s: corr(true, estimated) = %.4f   mean |error| = %.3g   (shape %s)"
              % (name, c, mae, "x".join(map(str, t.shape))))
        ok = ok and c >= lim
    if not ok:
        print("ERROR: the simulation truth was not recovered")
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
