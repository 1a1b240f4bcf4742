"""Gibbs-EM driver with the reference's `LogitNormalGEM` contract (gem.py:69-295).

Same constructor arguments, attributes (`A`, `alpha`, `pkappa`, `ptau`,
`suff_stats`, `path_elbo`, `init_*`) and methods (`gem`, `estep_gibbs`,
`init_gibbs`, `init_mle`, `init_para*`) as the reference, so `scUnif.py` and any
user script drive it unchanged.  The E-step samplers and the M-step it builds
are the sm_100a-backed classes of this package.

Extras (all optional keyword arguments, absent from the reference):
  * `seed`      -- seed of the counter-based RNG (the reference is unseeded);
  * `device_SC` / `device_BK` -- fp32 count matrices already resident in HBM
    (torch CUDA tensors holding THIS rank's rows), for inputs too large to
    stage through the host (configs c3-c5 of BASELINE.json);
  * multi-GPU: when `torch.distributed` is initialised, host matrices are taken
    to be the FULL problem and each rank keeps the contiguous row block
    `dist.shard_bounds` assigns it; device matrices are taken to be
    the rank's own block, with `L_total` / `M_total` / `row_offset_*` describing
    the global problem.  `local_rows=True` says the same of HOST matrices (each rank
    passes only its own rows; bench.py's end-to-end arm at N > 1).
"""
import logging

import numpy as np

from . import dist
from .e_step_gibbs import LogitNormalGibbs_BK, LogitNormalGibbs_SC
from .m_step import LogitNormalMLE
from .utils import simplex_proj, simplex_proj_columns  # noqa: F401


class LogitNormalGEM(object):
    """Gibbs-EM for the joint bulk + single-cell model (gem.py:69-295)."""

    def __init__(self, BKexpr=None, SCexpr=None, K=3, G=None, iMarkers=None,
                 init_A=None, min_A=1e-6, init_alpha=None, est_alpha=True,
                 init_pkappa=None, init_ptau=None,
                 burnin=200, sample=200, thin=1, bk_mean_approx=True,
                 MLE_CONV=1e-6, MLE_maxiter=100, EM_CONV=1e-3, EM_maxiter=100,
                 seed=None, device_SC=None, device_BK=None, L_total=None, M_total=None,
                 row_offset_SC=None, row_offset_BK=None, local_rows=False):
        self.hasBK = BKexpr is not None or device_BK is not None
        self.hasSC = SCexpr is not None or device_SC is not None
        self.K, self.min_A, self.est_alpha = K, min_A, est_alpha
        self.iMarkers = iMarkers
        self.EM_CONV, self.EM_maxiter = EM_CONV, EM_maxiter
        self.MLE_CONV, self.MLE_maxiter = MLE_CONV, MLE_maxiter
        self.burnin, self.sample, self.thin = burnin, sample, thin
        self.bk_mean_approx = bk_mean_approx
        self.seed = seed
        self._dev_SC, self._dev_BK = device_SC, device_BK

        # ---- row blocks owned by this rank --------------------------------
        self.SCexpr_full, self.BKexpr_full, self.G_full = SCexpr, BKexpr, G
        if self.hasSC:
            G = np.asarray(G).ravel().astype(np.int64)
            if device_SC is not None:
                self.L_local, self.N = int(device_SC.shape[0]), int(device_SC.shape[1])
                self.L = int(L_total) if L_total is not None else self.L_local
                self._off_SC = int(row_offset_SC or 0)
                self.SCexpr, self.G = None, G
            elif local_rows:
                self.L_local, self.N = SCexpr.shape
                self.L = int(L_total) if L_total is not None else self.L_local
                self._off_SC = int(row_offset_SC or 0)
                self.SCexpr, self.G = SCexpr, G
            else:
                self.L, self.N = SCexpr.shape
                lo, hi = dist.shard_bounds(self.L)
                self._off_SC, self.L_local = lo, hi - lo
                self.SCexpr, self.G = SCexpr[lo:hi], G[lo:hi]
            if self.G.shape[0] != self.L_local:
                raise ValueError("G must have one entry per (local) single cell")
            self.itype = [np.where(self.G == k)[0] for k in range(self.K)]
            self.type_counts = dist.all_reduce_np(
                np.array([len(t) for t in self.itype], dtype=float)).astype(np.int64)
        else:
            self.SCexpr, self.G, self.itype = None, None, None
        if self.hasBK:
            if device_BK is not None:
                self.M_local, self.N = int(device_BK.shape[0]), int(device_BK.shape[1])
                self.M = int(M_total) if M_total is not None else self.M_local
                self._off_BK = int(row_offset_BK or 0)
                self.BKexpr = None
            elif local_rows:
                self.M_local, self.N = BKexpr.shape
                self.M = int(M_total) if M_total is not None else self.M_local
                self._off_BK = int(row_offset_BK or 0)
                self.BKexpr = BKexpr
            else:
                self.M, self.N = BKexpr.shape
                lo, hi = dist.shard_bounds(self.M)
                self._off_BK, self.M_local = lo, hi - lo
                self.BKexpr = BKexpr[lo:hi]
        else:
            self.BKexpr = None

        if self.hasSC:
            dist.require_rows_everywhere(self.L_local, "single cells")
        if self.hasBK:
            dist.require_rows_everywhere(self.M_local, "bulk samples")
        # the device shards exist from here on: the counts are uploaded once, parameter
        # initialisation (below) and the samplers (init_gibbs) work on the same copy
        self._shard_SC = self._shard_BK = None
        if self.hasSC and self.L_local > 0:
            self._shard_SC = LogitNormalGibbs_SC(
                A=None, pkappa=None, ptau=None, SCexpr=self.SCexpr, G=self.G, itype=self.itype,
                row_offset=self._off_SC, total_rows=self.L, seed=self.seed,
                device_counts=self._dev_SC, K=self.K)
        if self.hasBK and self.M_local > 0:
            self._shard_BK = LogitNormalGibbs_BK(
                A=None, alpha=None, BKexpr=self.BKexpr, iMarkers=self.iMarkers,
                row_offset=self._off_BK, total_rows=self.M, seed=self.seed,
                device_counts=self._dev_BK, K=self.K)

        self.init_para(init_A, init_pkappa, init_ptau, init_alpha)

    # ------------------------------------------------------------------
    # parameter initialisation (gem.py:208-280)
    # ------------------------------------------------------------------
    def init_para(self, init_A, init_pkappa, init_ptau, init_alpha):
        if self.hasSC:
            self.init_para_SC(init_pkappa, init_ptau)
            self.init_alpha = None
        if self.hasBK:
            self.init_para_BK(init_alpha)
            # gem.py:219-220: these defaults overwrite the user's values for the
            # objects built in init_gibbs/init_mle; self.pkappa/self.ptau keep them
            # and reach the sampler at the first E-step
            self.init_pkappa = np.array([-1., 0.01], dtype=float)
            self.init_ptau = np.array([self.N, 0.01 * self.N], dtype=float)
        self.init_para_A(init_A)

    def _type_means_SC(self):
        """K x N per-type SUMS of row-normalised single-cell counts, over all ranks
        (csrc: normalised_colsum_kernel, one pass over the device-resident counts)."""
        sums = np.zeros((self.K, self.N))
        if self._shard_SC is not None:
            sums = self._shard_SC.type_sums()
        return dist.all_reduce_np(sums)

    def _mean_BK(self):
        """N-vector mean over ALL bulk samples of the row-normalised counts."""
        s = np.zeros(self.N)
        if self._shard_BK is not None:
            s = self._shard_BK.profile_sum()
        return dist.all_reduce_np(s) / float(self.M)

    def init_para_A(self, init_A):
        """gem.py:225-255."""
        K, N = self.K, self.N
        if init_A is not None:
            self.init_A = np.array(init_A, dtype=float)
        elif self.hasSC:
            self.init_A = np.ones([N, K], dtype=float) / N
            sums = self._type_means_SC()
            bkmean = self._mean_BK() if (self.hasBK and (self.type_counts == 0).any()) else None
            for k in range(K):
                if self.type_counts[k] > 0:
                    self.init_A[:, k] = sums[k] / float(self.type_counts[k])
                elif self.hasBK:
                    self.init_A[:, k] = bkmean
        else:
            bkmean = self._mean_BK()
            self.init_A = np.zeros([N, K])
            for k in range(K):
                self.init_A[:, k] = bkmean
        self.init_A = simplex_proj_columns(self.init_A, self.min_A)  # gem.py:252-254, one launch
        self.A = np.copy(self.init_A)

    def init_para_SC(self, init_pkappa, init_ptau):
        """gem.py:258-271."""
        if (init_pkappa is not None) and (len(init_pkappa) == 2):
            self.init_pkappa = np.array(init_pkappa, dtype=float)
        else:
            self.init_pkappa = np.array([-1., 0.01], dtype=float)
        if (init_ptau is not None) and (len(init_ptau) == 2):
            self.init_ptau = np.array(init_ptau, dtype=float)
        else:
            self.init_ptau = np.array([self.N, 0.01 * self.N], dtype=float)
        self.pkappa = np.copy(self.init_pkappa)
        self.ptau = np.copy(self.init_ptau)

    def init_para_BK(self, init_alpha):
        """gem.py:274-280."""
        if (init_alpha is not None) and (len(init_alpha) == self.K):
            self.init_alpha = np.array(init_alpha, dtype=float)
        else:
            self.init_alpha = np.ones([self.K])
        self.alpha = np.copy(self.init_alpha)

    # ------------------------------------------------------------------
    # construction of the E-step / M-step objects (gem.py:198-205,283-295)
    # ------------------------------------------------------------------
    def init_gibbs(self):
        if self.hasSC:
            if self._shard_SC is None:
                raise ValueError("this rank holds no single cells")
            self.Gibbs_SC = self._shard_SC
            self.Gibbs_SC.update_parameters(self.init_A, self.init_pkappa, self.init_ptau)
            self.Gibbs_SC.init_gibbs()
        if self.hasBK:
            if self._shard_BK is None:
                raise ValueError("this rank holds no bulk samples")
            self.Gibbs_BK = self._shard_BK
            self.Gibbs_BK.update_parameters(self.init_A, self.init_alpha)
            self.Gibbs_BK.init_gibbs()

    def init_mle(self):
        self.mle = LogitNormalMLE(
            BKexpr=self.BKexpr, SCexpr=self.SCexpr, G=self.G, K=self.K, itype=self.itype,
            hasBK=self.hasBK, hasSC=self.hasSC, init_A=self.init_A,
            init_alpha=self.init_alpha, init_pkappa=self.init_pkappa,
            init_ptau=self.init_ptau, min_A=self.min_A, MLE_CONV=self.MLE_CONV,
            MLE_maxiter=self.MLE_maxiter,
            L_total=self.L if self.hasSC else None, M_total=self.M if self.hasBK else None,
            type_counts=self.type_counts if self.hasSC else None, N=self.N)

    # ------------------------------------------------------------------
    # EM (gem.py:135-195)
    # ------------------------------------------------------------------
    def estep_gibbs(self):
        """gem.py:135-149.  Both samplers only enqueue their kernels (each on its own stream:
        the bulk sweeps fill the SMs the persistent single-cell kernel leaves idle in its
        tail); the statistics are collected afterwards."""
        self.suff_stats = {}
        if self.hasSC:
            self.Gibbs_SC.update_parameters(self.A, self.pkappa, self.ptau)
        if self.hasBK:
            self.Gibbs_BK.update_parameters(self.A, self.alpha)
        if self.hasSC:
            self.Gibbs_SC.gibbs(burnin=self.burnin, sample=self.sample, thin=self.thin)
        if self.hasBK:
            self.Gibbs_BK.gibbs(burnin=self.burnin, sample=self.sample, thin=self.thin,
                                mean_approx=self.bk_mean_approx)
        if self.hasSC:
            self.suff_stats.update(self.Gibbs_SC.suff_stats)
        if self.hasBK:
            self.suff_stats.update(self.Gibbs_BK.suff_stats)

    def em_iteration(self):
        """One body of the reference's while loop (gem.py:163-192); returns
        (elbo after the E-step, elbo after the M-step)."""
        ev = self._events_begin()
        self.estep_gibbs()
        self._events_mark(ev)
        self.mle.update_suff_stats(self.suff_stats)
        elbo_e = self.mle.compute_elbo()
        logging.info("\tE-step finished: elbo=%.6f", elbo_e)
        if self.hasSC:
            self.mle.opt_kappa_tau()
            self.pkappa = self.mle.pkappa
            self.ptau = self.mle.ptau
        if self.hasBK and self.est_alpha:
            self.mle.opt_alpha()
            self.alpha = self.mle.alpha
        self.mle.opt_A_u()
        self.A = self.mle.A
        elbo = self.mle.compute_elbo()
        self._events_mark(ev)
        logging.info("\tM-step finished: elbo=%.6f", elbo)
        return elbo_e, elbo

    # optional device-side timing of an EM iteration (bench.py): `record_events = True` makes
    # em_iteration append [start, after the E-step, end] CUDA events to `self.events`
    record_events = False

    def _events_begin(self):
        if not self.record_events:
            return None
        import torch
        ev = [torch.cuda.Event(enable_timing=True)]
        ev[0].record()
        if not hasattr(self, "events"):
            self.events = []
        self.events.append(ev)
        return ev

    def _events_mark(self, ev):
        if ev is not None:
            import torch
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append(e)

    def gem(self):
        (converged, niter) = (self.EM_CONV + 1, 0)
        old_elbo = -10 ** 6
        path_elbo = []
        self.path_elbo_estep = []
        self.init_gibbs()
        self.init_mle()
        elbo = None
        while (abs(converged) > self.EM_CONV) and (niter < self.EM_maxiter):
            logging.info("%d-th EM iteration started...", niter + 1)
            elbo_e, elbo = self.em_iteration()
            self.path_elbo_estep.append(elbo_e)
            converged = abs(elbo - old_elbo)
            old_elbo = elbo
            niter += 1
            path_elbo.append(elbo)
        self.path_elbo = np.array(path_elbo)
        return (niter, elbo, converged, self.path_elbo)

    # ------------------------------------------------------------------
    # outputs of a sharded run, gathered on every rank (for gem2csv)
    # ------------------------------------------------------------------
    def gathered_stats(self, local=False):
        """`suff_stats` entries that `scUnif.gem2csv` writes, with the per-cell /
        per-sample ones concatenated over ranks in global row order (`local=True`: this
        rank's rows only, no collective)."""
        out = {}
        st = self.suff_stats
        gather = (lambda a: np.ascontiguousarray(a)) if local else dist.gather_rows
        if self.hasSC:
            out["exp_S"] = gather(np.asarray(st["exp_S"]))
            for key in ("exp_kappa", "exp_tau", "exp_kappasq", "exp_tausq"):
                out[key] = gather(st[key].reshape(-1, 1)).reshape(1, -1)
        if self.hasBK:
            for key in ("exp_W", "exp_logW"):
                out[key] = np.ascontiguousarray(gather(st[key].T.copy()).T)
            out["exp_Zjk"] = gather(st["exp_Zjk"])
            out["exp_Zik"] = st["exp_Zik"]
        return out
