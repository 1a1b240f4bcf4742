#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time (host buffers -> public API ->
host results): phase timers with a device synchronize on both sides, then a cProfile
of one more step.  Run on the GPU box:  python tools/e2e_breakdown.py [--workload c2]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ursm_b200 import synth
    from ursm_b200.gem import LogitNormalGEM
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--bulk-mode", dest="bulk_mode", default="gibbs")
    ap.add_argument("--weak", action="store_true",
                    help="under torchrun: every rank holds the workload's rows (bench.py's e2e arm)")
    args = ap.parse_args()
    cfg = bench.WORKLOADS[args.workload]
    K, N, L, M = cfg["K"], cfg["N"], cfg["L"], cfg["M"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:  # torchrun: every rank keeps its block of the rows (sharded fit)
        import torch.distributed as td
        td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    A_true = synth.profile_matrix(N, K, seed=0)
    G = (np.arange(L) % K).astype(np.int64) if L else None
    hY = hX = None
    if L:
        dY = synth.cells_on_device(A_true, G, 0)[0]
        hY = torch.empty(L, N, dtype=torch.float64).pin_memory()
        hY.copy_(dY)
        del dY
    if M:
        dX = synth.bulk_on_device(A_true, M, 0)[0]
        hX = torch.empty(M, N, dtype=torch.float64).pin_memory()
        hX.copy_(dX)
        del dX
    nY = hY.numpy() if hY is not None else None
    nX = hX.numpy() if hX is not None else None
    init_A = [None]
    rank = int(os.environ.get("RANK", "0"))
    shard_kw = {}
    if args.weak and world > 1:
        shard_kw = dict(local_rows=True, L_total=L * world, M_total=M * world,
                        row_offset_SC=rank * L, row_offset_BK=rank * M)

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    def step(times):
        t = [sync()]
        g = LogitNormalGEM(BKexpr=nX, SCexpr=nY, K=K, G=G, burnin=cfg["burnin"],
                           sample=cfg["sample"], thin=1,
                           bk_mean_approx=(args.bulk_mode == "mean"), MLE_CONV=1e-6,
                           MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=10 ** 6, seed=1,
                           init_A=init_A[0], **shard_kw)
        t.append(sync())
        g.init_gibbs()
        t.append(sync())
        g.init_mle()
        t.append(sync())
        g.estep_gibbs()
        t.append(sync())
        g.mle.update_suff_stats(g.suff_stats)
        g.mle.compute_elbo()
        if g.hasSC:
            g.mle.opt_kappa_tau()
            g.pkappa, g.ptau = g.mle.pkappa, g.mle.ptau
        if g.hasBK:
            g.mle.opt_alpha()
            g.alpha = g.mle.alpha
        t.append(sync())
        g.mle.opt_A_u()
        g.A = g.mle.A
        g.mle.compute_elbo()
        t.append(sync())
        st = g.gathered_stats(local=bool(shard_kw))
        t.append(sync())
        if init_A[0] is None:
            init_A[0] = np.copy(g.init_A)
        del g, st
        t.append(sync())
        if times is not None:
            times.append(np.diff(t))

    step(None)
    step(None)
    times = []
    for _ in range(args.reps):
        step(times)
    if local != 0:
        sys.stdout = open(os.devnull, "w")
    names = ["ctor(init_para)", "init_gibbs(upload)", "init_mle", "estep", "suffstats+elbo+alpha",
             "opt_A_u+elbo", "gathered_stats(d2h)", "destroy"]
    tm = np.median(np.array(times), axis=0) * 1e3
    for n, v in zip(names, tm):
        print("%-24s %8.2f ms" % (n, v))
    print("%-24s %8.2f ms" % ("total", tm.sum()))
    # ---- steady state: one model, repeated EM iterations (bench.py's device-timed arm)
    g = LogitNormalGEM(BKexpr=nX, SCexpr=nY, K=K, G=G, burnin=cfg["burnin"], sample=cfg["sample"],
                       thin=1, bk_mean_approx=(args.bulk_mode == "mean"), MLE_CONV=1e-6,
                       MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=10 ** 6, seed=1, **shard_kw)
    g.init_gibbs()
    g.init_mle()
    rows = []
    for it in range(2 + args.reps):
        t = [sync()]
        g.estep_gibbs()
        t.append(sync())
        g.mle.update_suff_stats(g.suff_stats)
        t.append(sync())
        g.mle.compute_elbo()
        t.append(sync())
        if g.hasSC:
            g.mle.opt_kappa_tau()
            g.pkappa, g.ptau = g.mle.pkappa, g.mle.ptau
        if g.hasBK:
            g.mle.opt_alpha()
            g.alpha = g.mle.alpha
        t.append(sync())
        g.mle.opt_A_u()
        g.A = g.mle.A
        t.append(sync())
        g.mle.compute_elbo()
        t.append(sync())
        if it >= 2:
            rows.append(np.diff(t))
    tm = np.median(np.array(rows), axis=0) * 1e3
    print("steady state (same model, EM iterations 3..):")
    for n, v in zip(["estep", "update_suff_stats", "elbo", "kappa_tau+alpha (host)", "opt_A_u", "elbo"], tm):
        print("  %-24s %8.2f ms" % (n, v))
    print("  outer iterations %s, graph launches %s" % (list(g.mle.niter_A),
                                                        getattr(g.mle, "graph_launches", None)))
    del g
    pr = cProfile.Profile()
    pr.enable()
    step(None)
    pr.disable()
    ps = pstats.Stats(pr, stream=sys.stdout)
    ps.sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
