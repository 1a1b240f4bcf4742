#!/usr/bin/env python
"""Small driver for ncu captures of one kernel family at a BASELINE shape (run on the GPU box):

    python tools/prof_kernels.py bulk  --workload c2 [--sweeps 12]
    python tools/prof_kernels.py sc    --workload c2
    python tools/prof_kernels.py mstep --workload c3
    python tools/prof_kernels.py mean  --workload c4

Prints the device time of the part it ran (CUDA events inside the library)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ursm_b200 import synth
    from ursm_b200.gem import LogitNormalGEM
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["bulk", "sc", "mstep", "mean"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--sweeps", type=int, default=12)
    ap.add_argument("--rows", type=int, default=0, help="override the row count (cells or samples)")
    args = ap.parse_args()
    cfg = dict(bench.WORKLOADS[args.workload])
    K, N = cfg["K"], cfg["N"]
    L = cfg["L"] if args.what in ("sc", "mstep") else 0
    M = cfg["M"] if args.what in ("bulk", "mean") else 0
    if args.rows:
        L, M = (args.rows if L else 0), (args.rows if M else 0)
    torch.cuda.set_device(0)
    A_true = synth.profile_matrix(N, K, seed=0)
    G = (np.arange(L) % K).astype(np.int64) if L else None
    dY = synth.cells_on_device(A_true, G, 0)[0] if L else None
    dX = synth.bulk_on_device(A_true, M, 0)[0] if M else None
    burn = args.sweeps // 2
    gem = LogitNormalGEM(K=K, G=G, burnin=burn, sample=args.sweeps - burn, thin=1,
                         bk_mean_approx=(args.what == "mean"), MLE_CONV=1e-6, MLE_maxiter=500,
                         EM_CONV=1e-6, EM_maxiter=10 ** 6, seed=1, device_SC=dY, device_BK=dX,
                         L_total=L, M_total=M)
    gem.init_gibbs()
    gem.init_mle()
    for it in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        gem.estep_gibbs()
        ev[1].record()
        if args.what == "mstep":
            gem.mle.update_suff_stats(gem.suff_stats)
            gem.mle.opt_kappa_tau()
            gem.mle.opt_A_u()
            gem.A = gem.mle.A
        ev[2].record()
        torch.cuda.synchronize()
        msg = "iter %d: E-step %.3f ms" % (it, ev[0].elapsed_time(ev[1]))
        if gem.hasSC:
            msg += " (sc kernel %.3f ms)" % gem.Gibbs_SC.kernel_ms()
        if gem.hasBK:
            msg += " (bulk kernels %.3f ms)" % gem.Gibbs_BK.kernel_ms()
        if args.what == "mstep":
            msg += ", M-step %.3f ms, outer iterations %s" % (ev[1].elapsed_time(ev[2]),
                                                              list(gem.mle.niter_A))
        print(msg)


if __name__ == "__main__":
    main()
