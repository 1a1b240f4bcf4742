#!/usr/bin/env python
"""Benchmark of the URSM Monte-Carlo EM hot path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]

A *step* is one EM iteration (gem.py:163-192): the single-cell Gibbs E-step
(burnin + sample sweeps of draw_w / draw_S / draw_kappa_tau), the bulk Gibbs E-step
(draw_W / draw_Z, `-no_mean_approx`), update_suff_stats, the ELBO twice and the
M-step (opt_kappa_tau, opt_alpha, opt_A_u).  Work is counted in Gibbs cell x gene
updates: (L*N + M*N) * (burnin + sample*thin) per step.

  value     updates of all ranks / device time of the K timed steps, inputs resident in
            HBM (weak scaling: every rank holds one copy of the workload's rows)
  e2e       the same metric through the public API from HOST buffers, at any rank count:
            every step builds LogitNormalGEM from the rank's pinned host count matrices
            (upload), runs the EM iteration and reads A / alpha / E[kappa] / E[tau] /
            E[W] / E[Z] / E[S] back to host arrays; wall clock, max over ranks
  roofline  the single-cell Gibbs kernel against the measured HBM peak, using the
            algorithmic bytes per update of SURVEY 8d / DESIGN.md
  cpu_baseline   the CPU oracle (numpy restatement of the reference + compiled PG
            sampler) on a bounded row-subsample of the same workload, on this box's cores

`--impl reference` prints the CPU arm alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# per-GPU row counts of BASELINE.json's configs (weak scaling: each rank holds these)
WORKLOADS = {
    # configs[1]: the configuration the metric is quoted on for one GPU
    "c2": dict(K=5, L=2000, M=200, N=5000, burnin=50, sample=50, desc="synthetic joint fit K=5, "
               "2,000 cells, 200 bulk samples, 5,000 genes per GPU"),
    # configs[2] / 8: the per-GPU share of the 8-GPU target problem
    "c3": dict(K=10, L=12500, M=125, N=20000, burnin=50, sample=100, desc="1/8 of the K=10, "
               "100,000-cell x 1,000-bulk x 20,000-gene joint fit per GPU"),
    "c4": dict(K=10, L=0, M=2500, N=20000, burnin=50, sample=100, desc="bulk-only K=10, 2,500 "
               "bulk samples x 20,000 genes per GPU (1/8 of configs[3])"),
    "c5": dict(K=20, L=125000, M=0, N=2000, burnin=50, sample=100, desc="single-cell-only K=20, "
               "125,000 cells x 2,000 genes per GPU (1/8 of configs[4])"),
    "tiny": dict(K=3, L=64, M=16, N=400, burnin=5, sample=5, desc="tiny self-test"),
}
# warp instructions the single-cell kernel executes per (cell, gene, sweep) update: from the
# ncu capture summarised in profiles/ (smsp__inst_executed.sum / updates of the launch)
SC_WARP_INSTR_PER_UPDATE = 11.1
METRIC = "gibbs_cell_gene_updates_per_s"
UNIT = "updates/s"


def load_traffic(workload):
    """Measured DRAM bytes per launch of the dominant kernel (one ncu --set full capture,
    committed under profiles/); None when no capture exists for the workload."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        with open(path) as fh:
            ent = json.load(fh).get(workload)
        return None if ent is None else int(ent["bytes"])
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region: NVML polled every 10 ms from
    a thread of this process (nvidia_ml_py), `nvidia-smi -lms` as the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index
        self.nvml, self.thread, self.stop_flag = None, None, False
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.index])
            except Exception:
                return self.index
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                if mx:
                    self.mx.append(mx)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1e3)
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                    "sm_max_mhz": float(max(self.mx)) if self.mx else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons),
                    "power_w_max": float(max(self.power)) if self.power else None,
                    "source": "nvml, 10 ms polling during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 20"}


def updates_per_step(cfg, bulk_mode="gibbs"):
    sweeps = cfg["burnin"] + cfg["sample"]
    # the default bulk mode of the CLI is one deterministic mean update per EM iteration
    bulk = cfg["M"] * cfg["N"] * (sweeps if bulk_mode == "gibbs" else 1)
    return cfg["L"] * cfg["N"] * sweeps + bulk


# ---------------------------------------------------------------------------
# CPU arm: the oracle on a bounded row-subsample (the only place bench.py runs oracle/)
# ---------------------------------------------------------------------------
def cpu_sample_sizes(cfg, mode="subsample"):
    K = cfg["K"]
    if mode == "full":  # every row of the workload, burn-in 1 / sample 2
        return cfg["L"], cfg["M"], 1, 2
    Ls = min(cfg["L"], max(2 * K, 8)) if cfg["L"] else 0
    Ms = min(cfg["M"], 4) if cfg["M"] else 0
    # keep one step to a few seconds: ~3 us per single-cell entry, ~5 us per bulk entry
    per_sweep = 3e-6 * Ls * cfg["N"] + 5e-6 * Ms * cfg["N"]
    sweeps = int(max(2, min(cfg["burnin"] + cfg["sample"], 6.0 / max(per_sweep, 1e-9))))
    burn = sweeps // 2
    return Ls, Ms, burn, sweeps - burn


def run_cpu_arm(cfg, steps, warmup, seed=0, verbose=False, mode="subsample", mle_maxiter=500):
    """Times the oracle (kind "port": the reference is pure Python and cannot travel to
    the GPU box; its one native dependency, pypolyagamma, is restated in C/OpenMP).

    `value` has the GPU arm's definition: Gibbs updates of a step / seconds of the WHOLE step
    (E-step + update_suff_stats + 2 x ELBO + M-step).  Default: a bounded row-subsample at the
    full gene count with fewer sweeps, so that the driver's --steps/--warmup finish in minutes;
    `mode="full"` is BASELINE.md sec. 3's plan for c2 (every row, burn-in 1 / sample 2), which
    takes minutes per step on one host."""
    from oracle import ursm_oracle as orc
    K, N = cfg["K"], cfg["N"]
    Ls, Ms, burn, samp = cpu_sample_sizes(cfg, mode)
    cores = max(1, min(os.cpu_count() or 1, orc.pg_c_max_threads()))
    sim = orc.simulate(N=N, K=K, L=Ls, M=Ms, seed=seed)
    Y, X, G = sim.get("Y"), sim.get("X"), sim.get("G")
    rng = np.random.RandomState(seed + 1)
    gem = orc.OracleGEM(BKexpr=X, SCexpr=Y, K=K, G=G, burnin=burn, sample=samp, thin=1,
                        bk_mean_approx=False, MLE_CONV=1e-6, MLE_maxiter=mle_maxiter, EM_CONV=1e-6,
                        EM_maxiter=1, rng=rng, pg_factory=orc.CPGSampler, pg_draw=orc.pgdrawv_c)
    # build the samplers exactly as OracleGEM.gem() does, with `cores` PG generators
    if gem.hasSC:
        gem.sc = orc.OracleGibbsSC(gem.init_A, gem.init_pkappa, gem.init_ptau, Y, G, gem.itype,
                                   rng=rng, num_threads=cores, pg_factory=orc.CPGSampler,
                                   pg_draw=orc.pgdrawv_c)
        gem.sc.init_gibbs()
    if gem.hasBK:
        gem.bk = orc.OracleGibbsBK(gem.init_A, gem.init_alpha, X, None, rng=rng)
        gem.bk.init_gibbs()
    gem.mle = orc.OracleMLE(X, Y, G, K, gem.itype, gem.hasBK, gem.hasSC, gem.init_A,
                            gem.init_alpha, gem.init_pkappa, gem.init_ptau, 1e-6, 1, 1e-6,
                            mle_maxiter)
    upd = (Ls * N + Ms * N) * (burn + samp)
    e_times, m_times = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        gem.estep()
        t1 = time.perf_counter()
        gem.mle.update_suff_stats(gem.suff_stats)
        gem.mle.compute_elbo()
        if gem.hasSC:
            gem.mle.opt_kappa_tau()
            gem.pkappa, gem.ptau = gem.mle.pkappa, gem.mle.ptau
        if gem.hasBK:
            gem.mle.opt_alpha()
            gem.alpha = gem.mle.alpha
        gem.mle.opt_A_u()
        gem.A = gem.mle.A
        gem.mle.compute_elbo()
        t2 = time.perf_counter()
        if it >= warmup:
            e_times.append(t1 - t0)
            m_times.append(t2 - t1)
        if verbose:
            print("cpu step %d: E %.2fs M %.2fs" % (it, t1 - t0, t2 - t1), file=sys.stderr)
    e_mean, m_mean = float(np.mean(e_times)), float(np.mean(m_times))
    what = ("every row of the workload" if mode == "full" else
            "row-subsample of the workload at the full gene count")
    sample = ("%s: %d cells + %d bulk samples x %d genes, K=%d, burnin %d / sample %d sweeps per "
              "step, bulk Gibbs; value = Gibbs updates of the step / seconds of the whole step "
              "(E-step %.2f s + M-step %.2f s), the GPU arm's definition; same_config=false: "
              "fewer rows and sweeps than the GPU arm's step (rates per update are scale-invariant, "
              "the M-step's share is not)"
              % (what, Ls, Ms, N, K, burn, samp, e_mean, m_mean))
    return dict(value=upd / (e_mean + m_mean), unit=UNIT, cores=cores, kind="port", sample=sample,
                estep_s=e_mean, mstep_s=m_mean, updates_per_step=upd,
                estep_updates_per_s=upd / e_mean, same_config=False)


def reference_main(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    res = run_cpu_arm(cfg, args.steps, args.warmup, verbose=args.verbose, mode=args.cpu_mode,
                      mle_maxiter=args.mle_maxiter)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (res["estep_s"] + res["mstep_s"]), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + cfg["desc"], "K": cfg["K"], "N": cfg["N"],
                   "same_config": False},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "estep_s",
                                             "mstep_s", "estep_updates_per_s", "same_config")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
class StdoutToStderr(object):
    """Native libraries (NCCL's version banner, for one) write to file descriptor 1; the
    contract is ONE JSON line on stdout.  Everything written to fd 1 while this is active
    goes to stderr instead."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def gpu_main(args, cfg):
    with StdoutToStderr():
        line = gpu_run(args, cfg)
    if line is not None:
        print(json.dumps(line))
        sys.stdout.flush()
    return 0


def make_inputs(cfg, rank):
    """Synthetic counts for this rank's rows, generated on the device."""
    from ursm_b200 import synth
    K, N, Ll, Ml = cfg["K"], cfg["N"], cfg["L"], cfg["M"]
    A_true = synth.profile_matrix(N, K, seed=0)
    dev_SC = dev_BK = G = None
    if Ll:
        G = (np.arange(rank * Ll, (rank + 1) * Ll) % K).astype(np.int64)
        dev_SC = synth.cells_on_device(A_true, G, 0, row_offset=rank * Ll)[0]
    if Ml:
        dev_BK = synth.bulk_on_device(A_true, Ml, 0, row_offset=rank * Ml)[0]
    return dev_SC, dev_BK, G


def run_workload(name, cfg, args, steps, warmup, world, rank, flush, want_e2e, e2e_steps,
                 sampler=None):
    """Times `steps` EM iterations of workload `cfg` on this rank's rows (every rank holds
    cfg's row counts), device-timed with a barrier + synchronize on both sides, max over
    ranks; optionally the end-to-end arm from pinned host buffers."""
    import torch
    import torch.distributed as td
    from ursm_b200 import _lib
    from ursm_b200.gem import LogitNormalGEM
    lib = _lib.load()
    K, N, Ll, Ml = cfg["K"], cfg["N"], cfg["L"], cfg["M"]
    burnin, sample = cfg["burnin"], cfg["sample"]
    dev_SC, dev_BK, G = make_inputs(cfg, rank)

    def make_gem(**kw):
        return LogitNormalGEM(K=K, G=G, burnin=burnin, sample=sample, thin=1,
                              bk_mean_approx=(args.bulk_mode == "mean"), MLE_CONV=1e-6,
                              MLE_maxiter=args.mle_maxiter,
                              EM_CONV=1e-6, EM_maxiter=10 ** 6, seed=1, **kw)

    gem = make_gem(device_SC=dev_SC, device_BK=dev_BK, L_total=Ll * world, M_total=Ml * world,
                   row_offset_SC=rank * Ll, row_offset_BK=rank * Ml)
    gem.init_gibbs()
    gem.init_mle()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def one_step():
        flush.zero_()  # L2 flush between iterations (inputs at c2 are smaller than L2)
        return gem.em_iteration()

    for _ in range(warmup):
        one_step()
    barrier()
    if sampler is not None:
        sampler.start()
    lib.ursm_launch_count_reset()
    gem.record_events, gem.events = True, []
    sc_ms, bk_ms, outer, halv0 = [], [], [], gem.mle.nhalvings
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    t_wall = time.perf_counter()
    ev[0].record()
    elbo = None
    for _ in range(steps):
        _, elbo = one_step()
        if gem.hasSC:
            sc_ms.append(gem.Gibbs_SC.kernel_ms())
            outer.append([int(v) for v in gem.mle.niter_A])
        if gem.hasBK:
            bk_ms.append(gem.Gibbs_BK.kernel_ms())
    ev[1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(lib.ursm_launch_count())
    clocks = sampler.stop() if sampler is not None else None
    gem.record_events = False
    dev_ms = ev[0].elapsed_time(ev[1])
    e_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in gem.events]))
    m_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in gem.events]))
    t = torch.tensor([dev_ms, t_wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    upd_rank = updates_per_step(cfg, args.bulk_mode)
    res = {
        "workload": name, "desc": cfg["desc"], "K": K, "N": N, "cells_per_gpu": Ll,
        "bulk_per_gpu": Ml, "cells_total": Ll * world, "bulk_total": Ml * world,
        "burnin": burnin, "sample": sample, "steps": steps, "warmup": warmup,
        "value": upd_rank * world * steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / steps,
        "s_per_em_iter": dev_ms / steps / 1e3, "wall_ms_per_step": wall_ms / steps,
        "estep_ms": e_ms, "mstep_ms": m_ms,
        "kernel_ms": {"sc_gibbs": float(np.mean(sc_ms)) if sc_ms else None,
                      "bk_estep": float(np.mean(bk_ms)) if bk_ms else None},
        "mle_outer_iters_per_step": (float(np.mean([max(o) for o in outer])) if outer else 0.0),
        # the M-step's streaming pass reads one E[S] counter (1 B at sample <= 255, else
        # 2 B) per cell x gene of the still-active columns per outer iteration.  Lower bound on its
        # rate: counters of every column's own iterations / the WHOLE M-step time (trial steps,
        # gradient and host parts included)
        "mstep_stream": ({
            "bytes_per_counter": 1 if sample <= 255 else 2,
            "counter_bytes_read_per_step": float(np.mean([sum(o) for o in outer])) * (Ll / max(K, 1)) * N
            * (1 if sample <= 255 else 2),
            "gbs_over_whole_mstep": float(np.mean([sum(o) for o in outer])) * (Ll / max(K, 1)) * N
            * (1 if sample <= 255 else 2) / (m_ms * 1e-3) / 1e9,
            "peak_gbs": load_peaks()[0]} if outer and m_ms > 0 else None),
        "mle_outer_iters_by_column_last_step": outer[-1] if outer else None,
        "mle_halvings_per_step": (gem.mle.nhalvings - halv0) / float(steps),
        "sc_geometry": gem.Gibbs_SC.geometry() if gem.hasSC else None,
        "elbo_last": elbo, "gpu_launches": launches, "clocks": clocks,
        "updates_per_step_per_gpu": upd_rank,
    }

    # ---- e2e: host buffers -> public API -> host results, every step -------------
    if want_e2e:
        hY = hX = None
        if Ll:
            hY = torch.empty(Ll, N, dtype=torch.float64).pin_memory()
            hY.copy_(dev_SC)
        if Ml:
            hX = torch.empty(Ml, N, dtype=torch.float64).pin_memory()
            hX.copy_(dev_BK)
        nY = hY.numpy() if hY is not None else None
        nX = hX.numpy() if hX is not None else None
        h2d = (nY.nbytes if nY is not None else 0) + (nX.nbytes if nX is not None else 0)
        d2h_box = [0]
        del gem
        dev_SC = dev_BK = None

        def e2e_step():
            flush.zero_()
            # what a user of the reference does per fit: build the model from the host
            # count matrices (this rank's rows, float64 as np.loadtxt returns them; the
            # parameters are initialised from them, gem.py:208-280), run the EM iteration,
            # read every output gem2csv writes back to the host
            g2 = make_gem(SCexpr=nY, BKexpr=nX, local_rows=True,
                          L_total=Ll * world, M_total=Ml * world,
                          row_offset_SC=rank * Ll, row_offset_BK=rank * Ml)
            g2.init_gibbs()
            g2.init_mle()
            g2.em_iteration()
            out = [g2.A, g2.alpha if g2.hasBK else None]
            st = g2.gathered_stats(local=True)
            d2h_box[0] = sum(np.asarray(v).nbytes for v in st.values()) + g2.A.nbytes
            return out, st

        for _ in range(min(warmup, 3)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            td.all_reduce(te, op=td.ReduceOp.MAX)
        e2e_s = float(te[0])
        res["e2e"] = {
            "value": upd_rank * world / e2e_s, "unit": UNIT,
            "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h_box[0]) * world,
            "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
            "what": "per step and rank: LogitNormalGEM built from pinned host float64 count "
                    "matrices (upload), one EM iteration, A / alpha / E[S] / E[kappa] / E[tau] "
                    "/ E[W] / E[Z] read back to host arrays; wall clock, max over ranks"}
    else:
        res["e2e"] = None
    return res


def run_c1(verbose=False):
    """configs[0], the reference's `make run` (makefile:22-31): the demo arrays (committed
    fixture tests/golden/demo_data.npz = demo/demo_data/*.csv of the reference) at burn-in 10 /
    sample 20 / EM_maxiter 10 through the command-line path -- CSV in, fit, CSV out -- and the
    CPU port of the same fit on this box.  The reference's own published wall time for this
    run is 24.59 s (demo/demo_out/demo_logging.log:64, author's machine)."""
    import logging
    import tempfile
    from ursm_b200 import scUnif
    d = np.load(os.path.join(ROOT, "tests", "golden", "demo_data.npz"))
    tmp = tempfile.mkdtemp(prefix="ursm_c1_")
    np.savetxt(os.path.join(tmp, "sc.csv"), d["Y"], delimiter=",")
    np.savetxt(os.path.join(tmp, "bk.csv"), d["X"], delimiter=",")
    np.savetxt(os.path.join(tmp, "ct.csv"), d["G"], fmt="%d", delimiter=",")
    argv = ["-K", "3", "-sc", tmp + "/sc.csv", "-ctype", tmp + "/ct.csv", "-bk", tmp + "/bk.csv",
            "-outdir", tmp + "/out", "-log", tmp + "/out/log.log", "-burnin", "10", "-sample", "20",
            "-EM_maxiter", "10", "-verbose", "0", "-seed", "1"]
    times = []
    for rep in range(3):  # first call pays one-off CUDA context / module load
        t0 = time.perf_counter()
        gem = scUnif.main(argv)
        times.append(time.perf_counter() - t0)
    # the README's usage example / CLI defaults (BASELINE.json configs[0]): 50 / 50 / 50
    argv50 = list(argv)
    for flag, val in (("-burnin", "50"), ("-sample", "50"), ("-EM_maxiter", "50")):
        argv50[argv50.index(flag) + 1] = val
    t0 = time.perf_counter()
    gem50 = scUnif.main(argv50)
    t50 = time.perf_counter() - t0
    for h in list(logging.getLogger("").handlers):
        logging.getLogger("").removeHandler(h)
    out = {"cli_s_burnin50_sample50_EM50": t50, "em_iterations_50_50_50": int(len(gem50.path_elbo)),
           "config": "make run: K=3, 50 cells + 50 bulk x 100 genes, burnin 10 / sample 20 / "
                     "EM_maxiter 10, bulk mean update (CLI default)",
           "cli_s_first_call": times[0], "cli_s": float(min(times[1:])),
           "elbo_last": float(gem.path_elbo[-1]), "em_iterations": int(len(gem.path_elbo)),
           "reference_published_s": 24.59,
           "reference_published_source": "demo/demo_out/demo_logging.log:64 (author's machine)"}
    from oracle import ursm_oracle as orc
    t0 = time.perf_counter()
    og = orc.OracleGEM(BKexpr=d["X"], SCexpr=d["Y"], K=3, G=d["G"], burnin=10, sample=20, thin=1,
                       bk_mean_approx=True, MLE_CONV=1e-6, MLE_maxiter=500, EM_CONV=1e-6,
                       EM_maxiter=10, rng=np.random.RandomState(1), pg_factory=orc.CPGSampler,
                       pg_draw=orc.pgdrawv_c)
    og.gem()
    out["cpu_port_s"] = time.perf_counter() - t0
    out["cpu_port_elbo_last"] = float(og.path_elbo[-1])
    out["cpu_port_cores"] = max(1, min(os.cpu_count() or 1, orc.pg_c_max_threads()))
    return out


def p2p_check(world, rank):
    """Driver-visible correctness of the M-step's NVLink peer exchange (N > 1): the small
    joint fit of tests/test_gpu_multi.py through the peer-memory path and through the NCCL
    all-reduce path must give the same ELBO trajectory."""
    from oracle import ursm_oracle as orc  # data generator of the test (checker side)
    from ursm_b200.gem import LogitNormalGEM
    sim = orc.simulate(N=400, K=3, L=90, M=12, seed=5)
    res = {}
    for mode in ("1", "0"):
        os.environ["URSM_MLE_P2P"] = mode
        gem = LogitNormalGEM(BKexpr=sim["X"], SCexpr=sim["Y"], K=3, G=sim["G"], burnin=10,
                             sample=20, thin=1, bk_mean_approx=True, MLE_CONV=1e-6, MLE_maxiter=60,
                             EM_CONV=1e-9, EM_maxiter=3, seed=9)
        path = gem.gem()[3]
        res[mode] = (np.asarray(path), bool(getattr(gem.mle, "_p2p", False)))
    os.environ.pop("URSM_MLE_P2P", None)
    ok = bool(res["1"][1]) and (not res["0"][1]) and \
        bool(np.allclose(res["1"][0], res["0"][0], rtol=2e-5))
    return {"status": "ok" if ok else "MISMATCH", "peer_path_used": res["1"][1],
            "elbo_peer": [float(x) for x in res["1"][0]],
            "elbo_nccl": [float(x) for x in res["0"][0]], "rtol": 2e-5}


def gpu_run(args, cfg):
    import torch
    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB of L2
    sampler = ClockSampler(local) if rank == 0 else None
    main = run_workload(args.workload, cfg, args, args.steps, args.warmup, world, rank, flush,
                        not args.no_e2e, max(1, min(args.steps, 5)), sampler)

    # ---- the north-star configurations (BASELINE.json configs[2..4]): every rank holds 1/8 of
    # the full problem, so at --gpus 8 these ARE c3 / c4 / c5, strong-sharded over the box
    north = None
    if not args.no_north_star and args.workload == "c2":
        north = {"note": "per-GPU rows are 1/8 of BASELINE.json's configs[2..4]; with n_gpus = 8 "
                         "the figures are the full configurations sharded over the box, with "
                         "fewer GPUs they are that fraction of the rows (weak scaling)",
                 "n_gpus": world}
        for wl, st, e2e_st in (("c3", 5, 2), ("c5", 3, 0), ("c4", 3, 0)):
            c = WORKLOADS[wl]
            r = run_workload(wl, c, args, st, 1, world, rank, flush, e2e_st > 0, max(e2e_st, 1))
            keep = ("desc", "cells_total", "bulk_total", "K", "N", "burnin", "sample", "steps",
                    "s_per_em_iter", "value", "estep_ms", "mstep_ms", "kernel_ms",
                    "mle_outer_iters_per_step", "mstep_stream", "sc_geometry")
            ent = {k: r[k] for k in keep}
            ent["unit"] = UNIT
            if r["e2e"]:
                ent["e2e_s_per_em_iter"] = r["e2e"]["ms_per_step"] / 1e3
                ent["e2e_h2d_bytes_per_step"] = r["e2e"]["h2d_bytes_per_step"]
                ent["e2e_d2h_bytes_per_step"] = r["e2e"]["d2h_bytes_per_step"]
            north[wl] = ent
        # the command line's DEFAULT bulk mode (one deterministic mean update per EM
        # iteration instead of the Gibbs sweeps) on the headline workload, for comparison
        import copy
        margs = copy.copy(args)
        margs.bulk_mode = "mean"
        r = run_workload("c2", WORKLOADS["c2"], margs, 5, 2, world, rank, flush, False, 1)
        north["c2_bulk_mean_approx"] = {
            "desc": WORKLOADS["c2"]["desc"] + "; bulk E-step = mean update (CLI default)",
            "s_per_em_iter": r["s_per_em_iter"], "value": r["value"], "unit": UNIT,
            "estep_ms": r["estep_ms"], "mstep_ms": r["mstep_ms"], "kernel_ms": r["kernel_ms"],
            "updates_per_step_per_gpu": r["updates_per_step_per_gpu"]}
        if world == 8:
            north["target"] = {"what": "c3 on 8 x B200: one EM iteration (100 Gibbs samples) < 1 s",
                               "s_per_em_iter": north["c3"]["s_per_em_iter"],
                               "met": bool(north["c3"]["s_per_em_iter"] < 1.0)}
    check = p2p_check(world, rank) if world > 1 and not args.no_north_star else None

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return None

    K, N, Ll, Ml = cfg["K"], cfg["N"], cfg["L"], cfg["M"]
    burnin, sample = cfg["burnin"], cfg["sample"]
    peak, peak_src = load_peaks()
    roofline = None
    sweeps = burnin + sample
    sc_ms, bk_ms = main["kernel_ms"]["sc_gibbs"], main["kernel_ms"]["bk_estep"]
    if sc_ms:
        b_per_update = (2.0 * burnin + 6.0 * sample) / sweeps  # SURVEY 8d / DESIGN.md
        alg_bytes = b_per_update * Ll * N * sweeps
        achieved = alg_bytes / (sc_ms * 1e-3) / 1e9
        upd_s = Ll * N * sweeps / (sc_ms * 1e-3)
        issue_peak = 148 * 4 * 1.965e9  # warp instructions per second the chip can issue
        roofline = {"bound": "hbm", "kernel": "sc_gibbs_kernel", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": load_traffic(args.workload),
                    "algorithmic_bytes": alg_bytes, "peak_source": peak_src, "kernel_ms": sc_ms,
                    "bytes_per_update": b_per_update,
                    "note": "chain-resident kernel: issue-bound, not HBM-bound (DESIGN.md 4.1)",
                    "updates_per_s_kernel": upd_s,
                    "issue": {"warp_instr_per_update": SC_WARP_INSTR_PER_UPDATE,
                              "achieved_warp_instr_per_s": upd_s * SC_WARP_INSTR_PER_UPDATE,
                              "peak_warp_instr_per_s": issue_peak,
                              "frac": upd_s * SC_WARP_INSTR_PER_UPDATE / issue_peak,
                              "source": "instructions per update from the ncu capture under "
                                        "profiles/ (smsp__inst_executed.sum / updates)"}}
    elif bk_ms:
        passes = sweeps if args.bulk_mode == "gibbs" else 3  # mean update: three passes over X
        alg_bytes = 4.0 * Ml * N * passes
        achieved = alg_bytes / (bk_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "bk_split_kernel (all sweeps)" if args.bulk_mode == "gibbs"
                    else "bk_mean_rows_kernel x2 + bk_mean_cols_kernel (mean update, 3 passes)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (load_traffic(args.workload) * passes
                                if args.bulk_mode == "gibbs" and load_traffic(args.workload) else None),
                    "peak_source": peak_src, "kernel_ms": bk_ms, "bytes_per_update": 4.0,
                    "note": ("issue-bound lockstep binomial chains (DESIGN.md 4.2): 79 % of the issue "
                             "slots busy, X read once per sweep" if args.bulk_mode == "gibbs" else
                             "bound by the fp64 pipe (DESIGN.md 4.2): fp64 statistics are pinned to 1e-8")}
    cpu = c1 = None
    if world == 1 and not args.no_cpu:
        r = run_cpu_arm(cfg, steps=1, warmup=0, verbose=args.verbose)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "estep_s", "mstep_s",
                                 "estep_updates_per_s", "same_config")}
        if not args.no_north_star:
            c1 = run_c1(args.verbose)
    line = {
        "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 sampler / f64 reductions",
        "data": "synthetic (demo_simulate_data.py distributions, generated on device)",
        "config": {"workload": args.workload + ": " + cfg["desc"], "K": K, "N": N,
                   "cells_per_gpu": Ll, "bulk_per_gpu": Ml, "burnin": burnin, "sample": sample,
                   "thin": 1, "bulk_mode": "gibbs (-no_mean_approx)" if args.bulk_mode == "gibbs"
                   else "mean update (CLI default)",
                   "MLE_maxiter": args.mle_maxiter, "l2": "flushed between steps (256 MiB write)",
                   "parallelism": "cells/bulk samples sharded over %d rank(s)" % world},
        "s_per_em_iter": main["s_per_em_iter"], "wall_ms_per_step": main["wall_ms_per_step"],
        "estep_ms": main["estep_ms"], "mstep_ms": main["mstep_ms"],
        "mle_outer_iters_per_step": main["mle_outer_iters_per_step"],
        "mstep_stream": main["mstep_stream"],
        "mle_outer_iters_by_column_last_step": main["mle_outer_iters_by_column_last_step"],
        "mle_halvings_per_step": main["mle_halvings_per_step"],
        "kernel_ms": main["kernel_ms"], "sc_geometry": main["sc_geometry"],
        "elbo_last": main["elbo_last"], "roofline": roofline, "cpu_baseline": cpu,
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
        "north_star": north, "c1": c1, "p2p_check": check,
    }
    if world > 1:
        td.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ursm_b200", choices=["ursm_b200", "reference"])
    ap.add_argument("--workload", type=str, default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--mle-maxiter", dest="mle_maxiter", type=int, default=500)
    ap.add_argument("--bulk-mode", dest="bulk_mode", type=str, default="gibbs",
                    choices=["gibbs", "mean"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-north-star", dest="no_north_star", action="store_true",
                    help="skip the c3/c4/c5 block, the c1 CLI timing and the peer-exchange check")
    ap.add_argument("--cpu-mode", dest="cpu_mode", default="subsample", choices=["subsample", "full"],
                    help="--impl reference: bounded row-subsample (default) or every row at "
                         "burn-in 1 / sample 2 (BASELINE.md sec. 3; minutes per step)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    cfg = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_main(args, cfg)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # not launched by torchrun: re-launch ourselves one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                   "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)]
            cmd += sys.argv[1:]
            return subprocess.call(cmd)
    return gpu_main(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
