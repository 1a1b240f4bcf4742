"""M-step with the reference's class contract (m_step.py:13-320) on the device.

`LogitNormalMLE` keeps the reference constructor and the methods `gem.py` calls
(`update_suff_stats`, `compute_elbo`, `opt_kappa_tau`, `opt_alpha`, `opt_A_u`)
and attributes (`A`, `alpha`, `pkappa`, `ptau`).  The O(L*N) work (opt_u,
update_gd_coeffConst, the E[S]-weighted column sums) and the projected-gradient
loop for A run in csrc/mle.cu; the K-vector problems (`opt_alpha`,
`opt_kappa_tau`) stay on the host exactly as SURVEY 8a/a19-a20 scopes them.

The statistics must come from this package's samplers (their `suff_stats` carry
hidden handles to the device-resident counters); that is what `gem.py` passes.
"""
import ctypes
import logging

import numpy as np
from scipy.special import gammaln, psi

from . import _lib, dist


class LogitNormalMLE(object):

    def __init__(self, BKexpr=None, SCexpr=None, G=None, K=3, itype=None,
                 hasBK=True, hasSC=True, init_A=None, init_alpha=None,
                 init_pkappa=None, init_ptau=None, min_A=1e-6, min_alpha=1,
                 MLE_CONV=1e-4, MLE_maxiter=100,
                 L_total=None, M_total=None, type_counts=None, N=None):
        _lib.require_cuda()
        (self.K, self.min_A, self.min_alpha) = (K, min_A, min_alpha)
        (self.MLE_CONV, self.MLE_maxiter) = (MLE_CONV, MLE_maxiter)
        (self.SCexpr, self.G, self.BKexpr) = (SCexpr, G, BKexpr)
        (self.hasBK, self.hasSC, self.itype) = (hasBK, hasSC, itype)
        # global sizes (a rank may hold only a shard of the cells / samples)
        if self.hasSC:
            self.L = int(L_total if L_total is not None else SCexpr.shape[0])
            self.N = int(N if N is not None else SCexpr.shape[1])
            if type_counts is None:
                type_counts = dist.all_reduce_np(
                    np.array([len(itype[k]) for k in range(K)], dtype=float))
            self.type_counts = np.asarray(type_counts).astype(np.int64)
        if self.hasBK:
            self.M = int(M_total if M_total is not None else BKexpr.shape[0])
            self.N = int(N if N is not None else BKexpr.shape[1])
        self.A = np.copy(init_A)
        self.alpha = np.copy(init_alpha)
        self.pkappa = np.copy(init_pkappa)
        self.ptau = np.copy(init_ptau)
        self._h = None
        self.niter_A = np.zeros(K, dtype=int)
        self.nhalvings = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().ursm_mle_destroy(h)
            except Exception:
                pass

    # -- device workspace -------------------------------------------------
    def _ensure_workspace(self, sc, bk):
        import torch
        if self._h is not None and self._owners == (sc, bk):
            return
        if self._h is not None:
            _lib.call("ursm_mle_destroy", self._h)
        h = ctypes.c_void_p()
        _lib.call("ursm_mle_create", ctypes.byref(h), sc._h if sc is not None else None,
                  bk._h if bk is not None else None, self.N, self.K, float(self.min_A))
        self._h, self._owners = h, (sc, bk)
        KN = self.K * self.N
        f = dict(dtype=torch.float64, device="cuda")
        self._part_local = torch.zeros(KN + self.K, **f)   # per-pass local partials
        self._part = torch.zeros(KN + self.K, **f)         # all-reduced, per-column current
        self._cAinv = torch.zeros(KN, **f)
        self._cA = torch.zeros(KN, **f)
        self._coeffA = torch.zeros(KN, **f)
        self._zik = torch.zeros(KN, **f)
        _lib.call("ursm_mle_set_coeffs", self._h, self._p(self._cAinv), self._p(self._cA),
                  self._p(self._coeffA))
        self._p2p = self._setup_p2p() if sc is not None else False

    def _setup_p2p(self):
        """Peer-memory exchange for the per-iteration sum of opt_Ak (one NVLink domain, NCCL
        backend): every rank shares one IPC block, the handles travel once through
        torch.distributed.  `URSM_MLE_P2P=0` keeps the NCCL all-reduce (A/B measurements).

        Blocks are cached per process (the reference's API builds a new M-step object per
        fit).  A cached block is reused only when ONE small all-reduce shows that every rank
        holds a usable mapping of the same block at the same epoch; anything else (first
        use, a rank that picked another block, a fit that was aborted half way, a barrier
        that timed out) goes through the full handshake: handles exchanged, flag words /
        error word / epoch reset between two barriers -- or, if the ranks' cached mappings
        cannot be trusted, the NCCL path for this fit."""
        import os
        import torch
        import torch.distributed as td
        world = dist.world_size()
        if world < 2 or world > 16 or os.environ.get("URSM_MLE_P2P", "1") == "0":
            return False
        if td.get_backend() != "nccl":
            return False
        handle = (ctypes.c_char * 64)()
        cached = ctypes.c_int32(0)
        state = np.array([-1, -1, -1, 0], dtype=np.int64)  # usable, block id, epoch, opened
        ok = 1
        try:
            _lib.call("ursm_mle_p2p_alloc", self._h, ctypes.cast(handle, ctypes.c_void_p),
                      ctypes.byref(cached))
            _lib.call("ursm_mle_p2p_state", self._h, _lib.ptr(state))
        except _lib.UrsmError as e:  # pragma: no cover
            logging.warning("peer-memory exchange unavailable (%s); using NCCL", e)
            ok = 0
            state[:] = (-1, -1 - dist.rank(), -1, 0)  # cannot agree with anybody
        # ---- fast path: one collective tells whether all ranks agree on (usable, id, epoch)
        probe = torch.tensor([state[0], -state[0], state[1], -state[1], state[2], -state[2]],
                             dtype=torch.int64, device="cuda")
        td.all_reduce(probe, op=td.ReduceOp.MAX)
        v = [int(x) for x in probe.tolist()]
        agree = v[0] == -v[1] and v[2] == -v[3] and v[4] == -v[5]
        if agree and v[0] == 1:
            _lib.call("ursm_mle_p2p_open", self._h, None, world, dist.rank())
            return True
        # ---- slow path: first use, or the ranks' cached state differs
        gathered = [None] * world
        td.all_gather_object(gathered, (ok, int(state[3]), int(state[1]), bytes(handle.raw)))
        n_opened = sum(g[1] for g in gathered)
        if not all(g[0] for g in gathered) or len(set(g[2] for g in gathered)) != 1 \
                or n_opened not in (0, world):
            # a rank cannot take part, or the ranks hold different blocks / only some of them
            # hold mappings: do not guess, this fit sums through NCCL
            if ok:
                _lib.call("ursm_mle_p2p_disable", self._h)
            return False
        blob = b"".join(g[3] for g in gathered)
        try:
            _lib.call("ursm_mle_p2p_open", self._h, None if n_opened else ctypes.c_char_p(blob),
                      world, dist.rank())
            ok = 1
        except _lib.UrsmError as e:  # pragma: no cover
            logging.warning("peer-memory exchange unavailable (%s); using NCCL", e)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        td.all_reduce(flag, op=td.ReduceOp.MIN)  # all ranks or none; also barrier #1
        if not int(flag.item()):
            _lib.call("ursm_mle_p2p_disable", self._h)
            return False
        _lib.call("ursm_mle_p2p_reset", self._h)  # zero my flag words / error word / epoch
        td.all_reduce(flag, op=td.ReduceOp.MIN)   # barrier #2: nobody signals before all reset
        flag.item()
        return True

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(t.data_ptr())

    def _push_A(self):
        A = _lib.f64(self.A)
        _lib.call("ursm_mle_set_A", self._h, _lib.ptr(A))

    def _pull_A(self):
        A = np.empty((self.N, self.K))
        _lib.call("ursm_mle_get_A", self._h, _lib.ptr(A))
        self.A = A

    def _pass_u(self, active):
        """opt_u (:108-115) + update_gd_coeffConst (:118-125) for the active columns."""
        act = np.ascontiguousarray(active, dtype=np.int32)
        st = _lib.stream_ptr()
        _lib.call("ursm_mle_pass_u", self._h, _lib.ptr(act), self._p(self._part_local), st)
        dist.all_reduce_(self._part_local)  # N*K + K values
        # installs the rows of the active columns and the per-type sum_l R_l log u_l
        _lib.call("ursm_mle_opt_merge", self._h, self._p(self._part_local), self._p(self._part), st)

    # -- reference API ------------------------------------------------------
    def update_suff_stats(self, suff_stats):
        """m_step.py:50-89."""
        import torch
        self.suff_stats = suff_stats
        deref = lambda r: r() if r is not None else None  # the samplers hand out weak references
        sc = deref(suff_stats.get("_ursm_sc")) if self.hasSC else None
        bk = deref(suff_stats.get("_ursm_bk")) if self.hasBK else None
        if self.hasSC and sc is None:
            sc = self._adopt_foreign_sc(suff_stats)
        if self.hasBK and bk is None:
            bk = self._adopt_foreign_bk(suff_stats)
        self._ensure_workspace(sc, bk)
        K, N, KN = self.K, self.N, self.K * self.N
        self._push_A()
        self.elbo_const = 0
        self._cAinv.zero_()
        self._cA.zero_()
        self._coeffA.zero_()
        if self.hasBK:
            tail = suff_stats["_ursm_bk_tail"]
            self.elbo_const += float(tail[K])
            self._zik.copy_(bk._stats[:KN])
            self._cAinv += self._zik
            self.suff_stats["exp_mean_logW"] = np.asarray(tail[:K]) / float(self.M)
        if self.hasSC:
            self.elbo_const += suff_stats["exp_elbo_const"]
            self._sums = np.asarray(suff_stats["_ursm_sc_sums"], dtype=float)
            self._coeffA.copy_(sc._stats[:KN])
            self._cA.copy_(sc._stats[KN:2 * KN])
            self._cA *= 2.0
            # sum_{l in k} Y_li E[S_li]
            _lib.call("ursm_mle_begin", self._h, self._p(self._part_local), _lib.stream_ptr())
            dist.all_reduce_(self._part_local)
            self._cAinv += self._part_local[:KN]
            self._has_cells = self.type_counts > 0
            self._pass_u(self._has_cells)
        self._refresh_host_coeffs()

    # -- statistics that did not come from this package's samplers -----------------------
    # The reference's M-step takes any dict with the keys of e_step_gibbs.py:59-66,228-242
    # (m_step.py:50-89).  Such a dict holds host arrays; they are uploaded into a device shard
    # built from the count matrices the constructor was given, and the device path runs as
    # usual.  (One process only: a foreign dict carries no sharding information.)
    def _adopt_foreign_sc(self, st):
        import torch
        from .e_step_gibbs import LogitNormalGibbs_SC
        if dist.world_size() > 1:
            raise _lib.UrsmError("suff_stats from another sampler are supported in one process only")
        if self.SCexpr is None or self.G is None:
            raise _lib.UrsmError("LogitNormalMLE needs SCexpr and G to adopt statistics that were "
                                 "not produced by ursm_b200.e_step_gibbs")
        for key in ("exp_S", "exp_kappa", "exp_tau", "exp_kappasq", "exp_tausq", "coeffA",
                    "coeffAsq", "exp_elbo_const"):
            if key not in st:
                raise _lib.UrsmError("suff_stats lacks %r (e_step_gibbs.py:228-242)" % key)
        eS = np.asarray(st["exp_S"], dtype=float)
        # exp_S accumulates S / sample (e_step_gibbs.py:247): recover the integer counts
        vals = np.unique(eS)
        pos = vals[vals > 0]
        gaps = np.diff(np.concatenate([[0.0], pos]))
        step = gaps[gaps > 1e-9].min() if pos.size else 1.0
        sample = int(round(1.0 / step))
        cnt = eS * sample
        if not (1 <= sample <= 65535) or np.abs(cnt - np.round(cnt)).max() > 1e-6 * max(sample, 1):
            raise _lib.UrsmError("exp_S is not a multiple of 1/sample with sample <= 65535: not "
                                 "the statistic of a Gibbs run (e_step_gibbs.py:247)")
        sc = getattr(self, "_foreign_sc", None)
        if sc is None:
            itype = self.itype if self.itype is not None else \
                [np.where(np.asarray(self.G) == k)[0] for k in range(self.K)]
            sc = LogitNormalGibbs_SC(self.A, self.pkappa, self.ptau, self.SCexpr, self.G, itype)
            self._foreign_sc = sc
        L, N, K = sc.L, self.N, self.K
        mom = np.ascontiguousarray(np.stack([np.ravel(st[k]) for k in
                                             ("exp_kappa", "exp_tau", "exp_kappasq", "exp_tausq")]),
                                   dtype=np.float64)
        cnt16 = np.ascontiguousarray(np.round(cnt), dtype=np.uint16)
        _lib.call("ursm_sc_set_counts", sc._h, _lib.ptr(cnt16), sample, _lib.ptr(mom))
        host = np.zeros(2 * N * K + 6)
        host[:N * K] = np.asarray(st["coeffA"], dtype=float).T.ravel()
        host[N * K:2 * N * K] = np.asarray(st["coeffAsq"], dtype=float).T.ravel()
        host[2 * N * K] = float(st["exp_elbo_const"])
        host[2 * N * K + 1:2 * N * K + 5] = [mom[0].sum(), mom[2].sum(), mom[1].sum(), mom[3].sum()]
        sc._stats.copy_(torch.from_numpy(host))
        st["_ursm_sc_sums"] = host[2 * N * K + 1:2 * N * K + 5].copy()
        return sc

    def _adopt_foreign_bk(self, st):
        import torch
        if dist.world_size() > 1:
            raise _lib.UrsmError("suff_stats from another sampler are supported in one process only")
        for key in ("exp_Zik", "exp_Zjk", "exp_logW"):
            if key not in st:
                raise _lib.UrsmError("suff_stats lacks %r (e_step_gibbs.py:59-66)" % key)
        N, K = self.N, self.K
        eLogW = np.asarray(st["exp_logW"], dtype=float)           # K x M
        host = np.zeros(N * K + K + 2)
        host[:N * K] = np.asarray(st["exp_Zik"], dtype=float).T.ravel()
        host[N * K:N * K + K] = eLogW.sum(axis=1)
        host[N * K + K] = float((np.asarray(st["exp_Zjk"], dtype=float).T * eLogW).sum())

        class _HostBulkStats(object):   # what update_suff_stats reads from a bulk sampler
            _h = None
        bk = _HostBulkStats()
        bk._stats = torch.from_numpy(host).to("cuda")
        st["_ursm_bk_tail"] = host[N * K:N * K + K + 1].copy()
        self._foreign_bk = bk
        return bk

    def _refresh_host_coeffs(self):
        """The reference's attributes `gd_coeffAinv`, `gd_coeffA`, `gd_coeffConst` (N x K) and
        `u` (L) mirror device buffers; they are pulled when somebody reads them (tests,
        inspection), not on every update -- the EM loop itself never needs them."""
        self._host_cache = {}

    def _host(self, name):
        c = self.__dict__.setdefault("_host_cache", {})
        if name not in c:
            K, N, KN = self.K, self.N, self.K * self.N
            t = lambda x: np.ascontiguousarray(x.cpu().numpy().reshape(K, N).T)
            if name == "gd_coeffAinv":
                c[name] = t(self._cAinv)
            elif name == "gd_coeffA":
                c[name] = t(self._cA)
            elif name == "gd_coeffConst":
                c[name] = t(self._coeffA - self._part[:KN]) if self.hasSC else np.zeros((N, K))
            elif name == "u":
                u = np.empty(self._owners[0].L)
                _lib.call("ursm_mle_get_u", self._h, _lib.ptr(u))
                c[name] = u
        return c[name]

    gd_coeffAinv = property(lambda self: self._host("gd_coeffAinv"))
    gd_coeffA = property(lambda self: self._host("gd_coeffA"))
    gd_coeffConst = property(lambda self: self._host("gd_coeffConst"))
    u = property(lambda self: self._host("u"))

    def opt_kappa_tau(self):
        """m_step.py:92-103 (closed form; global sums came through the all-reduce)."""
        sk, skk, st, stt = self._sums
        km, tm = sk / self.L, st / self.L
        self.pkappa = np.array([km, skk / self.L - km ** 2])
        self.ptau = np.array([tm, stt / self.L - tm ** 2])

    def opt_u(self):
        """m_step.py:108-115 (all types)."""
        self._push_A()
        self._pass_u(self._has_cells)
        self._refresh_host_coeffs()

    def opt_A_u(self):
        """m_step.py:128-154: every column with cells runs opt_Ak (:157-182)
        concurrently (columns are independent); the others use the closed form.
        The loop state lives on the device; the host only enqueues iterations and
        polls the per-column flags every `poll_every` iterations."""
        if self.hasBK and self.hasSC:
            init_step = 0.01 / (self.L + self.M)
        elif self.hasBK:
            init_step = 0.01 / self.M
        else:
            init_step = 0.01 / self.L
        K = self.K
        self._push_A()
        active = np.zeros(K, dtype=np.int32)
        for k in range(K):
            if self.hasSC and self.type_counts[k] > 0:
                active[k] = 1
            else:
                _lib.call("ursm_mle_closed_form", self._h, k, self._p(self._zik),
                          _lib.stream_ptr())
                logging.debug("\t\tOptimized A%d with closed form solution", k)
        self.niter_A[:] = 0
        self.graph_launches = 0
        if active.any():
            st = _lib.stream_ptr()
            multi = dist.world_size() > 1
            niter = np.zeros(K, dtype=np.int32)
            halv = np.zeros(K, dtype=np.int32)
            if self._p2p or not multi:
                # the whole loop on the device: a captured graph of 16 outer iterations per
                # launch (step + decision + installation, then the opt_u / coeffConst pass; with
                # peers, the N*K + K values per iteration of SURVEY 8e are summed over NVLink
                # peer memory inside the same graph -- no collective call)
                nl = ctypes.c_int32(0)
                _lib.call("ursm_mle_opt_run", self._h, _lib.ptr(active), float(init_step),
                          float(self.MLE_CONV), int(self.MLE_maxiter), self._p(self._part),
                          _lib.ptr(niter), _lib.ptr(halv), ctypes.byref(nl), st)
                self.graph_launches = int(nl.value)
                if self._p2p:
                    err = ctypes.c_int32(0)
                    _lib.call("ursm_mle_p2p_error", self._h, ctypes.byref(err))
                    if err.value:
                        raise _lib.UrsmError("M-step peer exchange: a rank did not reach the "
                                             "iteration barrier (timeout)")
            else:
                # several ranks without the peer exchange (URSM_MLE_P2P=0, gloo, > 16 ranks): one
                # NCCL all-reduce per outer iteration, flags polled every `poll_every` iterations
                _lib.call("ursm_mle_opt_begin", self._h, _lib.ptr(active), float(init_step),
                          float(self.MLE_CONV), int(self.MLE_maxiter), st)
                it, poll_every = 0, 8
                while True:
                    _lib.call("ursm_mle_opt_step", self._h, self._p(self._part), st)
                    _lib.call("ursm_mle_opt_pass", self._h, self._p(self._part_local), st)
                    dist.all_reduce_(self._part_local)
                    _lib.call("ursm_mle_opt_merge", self._h, self._p(self._part_local),
                              self._p(self._part), st)
                    it += 1
                    if it % poll_every == 0:
                        _lib.call("ursm_mle_opt_poll", self._h, _lib.ptr(active), _lib.ptr(niter),
                                  _lib.ptr(halv), None, st)
                        if not active.any():
                            break
            self.niter_A[:] = niter
            self.nhalvings += int(halv.sum())
            for k in range(K):
                if niter[k] > 0:
                    logging.debug("\t\tOptimized A%d in %d iterations", k, niter[k])
            # u, gd_coeffConst and sum_l R_l log u_l of the final columns (the loop leaves the
            # partials of columns that stopped early behind)
            self._pass_u(self._has_cells)
        self._pull_A()
        self._refresh_host_coeffs()

    def opt_alpha(self):
        """m_step.py:185-206: backtracking gradient descent on the K-vector (host)."""
        (converged, niter) = (self.MLE_CONV + 1, 0)
        while (converged > self.MLE_CONV) and (niter < self.MLE_maxiter):
            old_alpha = np.copy(self.alpha)
            self.alpha = self.backtracking(
                old_val=old_alpha, grad_func=lambda a: (-self.get_grad_alpha(a)),
                obj_func=lambda a: (-self.get_obj_alpha(a)), init_step=10)[0]
            self.alpha = np.maximum(self.min_alpha, self.alpha)
            niter += 1
            converged = np.linalg.norm(self.alpha - old_alpha)
        logging.debug("\t\tOptimized alpha in %s iterations", niter)
        return niter

    def backtracking(self, old_val, grad_func, obj_func, init_step=0.1, proj_func=None):
        """m_step.py:209-233 (host version, used for alpha only)."""
        grad_old, obj_old = grad_func(old_val), obj_func(old_val)
        stepsize = init_step

        def trial(s):
            v = old_val - s * grad_old
            return proj_func(v) if proj_func is not None else v

        new_val = trial(stepsize)
        Gt = (old_val - new_val) / stepsize
        obj_new = obj_func(new_val)
        while obj_new > (obj_old + (stepsize * 0.5) * (Gt ** 2).sum()
                         - stepsize * np.dot(grad_old, Gt)):
            stepsize = stepsize * 0.5
            new_val = trial(stepsize)
            Gt = (old_val - new_val) / stepsize
            obj_new = obj_func(new_val)
        return (new_val, obj_new, stepsize)

    def get_grad_alpha(self, alpha_val):
        """m_step.py:236-241."""
        return self.suff_stats["exp_mean_logW"] + psi(sum(alpha_val)) - psi(alpha_val)

    def get_obj_alpha(self, alpha_val):
        """m_step.py:244-251 (scaled by 1/M)."""
        obj = np.dot(self.suff_stats["exp_mean_logW"], alpha_val)
        obj += gammaln(sum(alpha_val)) - sum(gammaln(alpha_val))
        return obj

    def compute_elbo(self):
        """m_step.py:290-320 (scaled by 1/N)."""
        KN = self.K * self.N
        terms = np.zeros(3)
        self._push_A()
        _lib.call("ursm_mle_elbo_terms", self._h,
                  self._p(self._part) if self.hasSC else None, _lib.ptr(terms),
                  _lib.stream_ptr())
        elbo = self.elbo_const
        elbo += terms[0]
        if self.hasBK:
            elbo += self.get_obj_alpha(self.alpha) * self.M
        if self.hasSC:
            elbo += terms[1]
            elbo += terms[2]
            elbo -= float(self._part[KN:].sum().item())
            sk, skk, st, stt = self._sums
            elbo -= np.log(self.pkappa[1] * self.ptau[1]) * self.L / 2.0
            elbo -= (skk - 2 * self.pkappa[0] * sk + self.L * (self.pkappa[0] ** 2)) \
                / (2.0 * self.pkappa[1])
            elbo -= (stt - 2 * self.ptau[0] * st + self.L * (self.ptau[0] ** 2)) \
                / (2.0 * self.ptau[1])
        return elbo / self.N
