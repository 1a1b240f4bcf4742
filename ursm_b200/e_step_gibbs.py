"""Gibbs E-step samplers with the reference's class contract (e_step_gibbs.py),
backed by the sm_100a kernels behind include/ursm_b200.h.

    LogitNormalGibbs_SC(A, pkappa, ptau, SCexpr, G, itype)      e_step_gibbs.py:171-377
    LogitNormalGibbs_BK(A, alpha, BKexpr, iMarkers)             e_step_gibbs.py:19-164

Same constructor arguments, methods (`init_gibbs`, `update_parameters`, `gibbs`)
and `suff_stats` keys as the reference, so `gem.py` drives them unchanged.  Extra
keyword arguments (`row_offset`, `total_rows`, `seed`) are optional and only used
for multi-GPU sharding and reproducibility.

Differences that a caller can observe (all deliberate):
  * the chain state (psi, w, S, Z, AW) never exists on the host; `suff_stats`
    holds the reference's keys, with `exp_S` materialised lazily from the
    device-resident 8- or 16-bit counters when somebody reads it;
  * random numbers come from counter-based Philox keyed on (seed, EM iteration,
    global cell/sample index, gene, sweep) instead of the unseeded global numpy
    state, so the chains of an E-step depend only on (seed, parameters, data) and not on
    the GPU count (the K x N statistics are summed with floating-point atomics, so a whole
    fit repeats to rounding, not bit for bit).
"""
import ctypes
import logging
import weakref

import numpy as np

from . import _lib, dist

_DEFAULT_SEED = 20171106


class _on_side_stream(object):
    """Context: enqueue on the sampler's own (non-blocking) stream, ordered after everything
    already enqueued on the caller's stream."""

    def __init__(self, sampler):
        self.sampler = sampler

    def __enter__(self):
        torch = self.sampler._torch
        if self.sampler._stream is None:
            self.sampler._stream = torch.cuda.Stream()
        self.sampler._stream.wait_stream(torch.cuda.current_stream())
        self.ctx = torch.cuda.stream(self.sampler._stream)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self.ctx.__exit__(*exc)


def _join_side_stream(sampler):
    """The caller's stream waits for what the sampler has enqueued on its own."""
    if sampler._stream is not None:
        sampler._torch.cuda.current_stream().wait_stream(sampler._stream)


class _Shard(object):
    """Owner of a device shard handle (ursm_sc / ursm_bk): destroyed when the last Python
    object that needs the device data is gone.  The sampler and the lazy `exp_S` view both hold
    one, so a `suff_stats` dict stays readable after its sampler has been dropped (the
    reference's dicts are plain host arrays) without a sampler <-> dict reference cycle."""

    def __init__(self, h, destroy):
        self.h, self._destroy = h, destroy

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h:
            try:
                getattr(_lib.load(), self._destroy)(h)
            except Exception:
                pass


class DeviceExpS(object):
    """`suff_stats["exp_S"]` (L x N, e_step_gibbs.py:232,247): a lazy view of the
    8- or 16-bit counters that stay on the device (local shard rows)."""

    def __init__(self, sampler):
        # the shard handle, not the sampler: the sampler owns the dict this object sits in, and
        # a reference back to it would leave whole device shards to the garbage collector
        self._shard = sampler._shard
        self.shape = (sampler.L, sampler.N)
        self.dtype = np.dtype(np.float64)
        self.ndim = 2

    def rows(self, lo, hi):
        out = _lib.host_result((hi - lo, self.shape[1]))
        _lib.call("ursm_sc_get_exp_S", self._shard.h, _lib.ptr(out), int(lo), int(hi - lo))
        return out

    def __array__(self, dtype=None, copy=None):
        a = self.rows(0, self.shape[0])
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            return self.rows(int(key), int(key) + 1)[0]
        return np.asarray(self)[key]

    def __len__(self):
        return self.shape[0]

    def sum(self, *a, **k):
        return np.asarray(self).sum(*a, **k)


class LogitNormalGibbs_SC(object):
    """Single-cell Gibbs sampler (e_step_gibbs.py:171-377)."""

    def __init__(self, A, pkappa, ptau, SCexpr, G, itype, row_offset=0, total_rows=None,
                 seed=None, device_counts=None, K=None):
        """`A=None` (with `K`) builds the device shard before any parameters exist -- the
        driver does that to initialise the parameters from the device-resident counts --
        and `update_parameters` installs them later."""
        torch = _lib.require_cuda()
        self._torch = torch
        (self.SCexpr, self.G) = (SCexpr, G)
        data = SCexpr if SCexpr is not None else device_counts
        self.L = int(data.shape[0])
        (self.N, self.K) = A.shape if A is not None else (int(data.shape[1]), int(K))
        self.itype = itype
        self.row_offset = int(row_offset)
        self.total_rows = int(total_rows) if total_rows is not None else self.L
        self.A = np.array(A, dtype=float, copy=True) if A is not None else None
        self.pkappa = np.array(pkappa, dtype=float, copy=True) if pkappa is not None else None
        self.ptau = np.array(ptau, dtype=float, copy=True) if ptau is not None else None
        self.seed = _DEFAULT_SEED if seed is None else int(seed)
        self.em_iter = 0
        Gi = _lib.i64(np.asarray(G).ravel())
        h = ctypes.c_void_p()
        if device_counts is not None:  # fp32 L x N torch tensor already in HBM
            self._keep = device_counts
            _lib.call("ursm_sc_create_dev", ctypes.byref(h), ctypes.c_void_p(device_counts.data_ptr()),
                      _lib.ptr(Gi), self.L, self.N, self.K, self.row_offset)
        else:
            Y = _lib.f64(SCexpr)
            _lib.call("ursm_sc_create", ctypes.byref(h), _lib.ptr(Y), _lib.ptr(Gi), self.L,
                      self.N, self.K, self.row_offset)
        self._shard = _Shard(h, "ursm_sc_destroy")
        n = _lib.load().ursm_sc_stats_len(h)
        self._stats = torch.zeros(n, dtype=torch.float64, device="cuda")
        if self.A is not None:
            self._push_params()
        self._suff, self._pending, self._stream = {}, False, None

    # `gibbs()` only ENQUEUES the E-step (on this sampler's own stream, so that the single-cell
    # and the bulk E-step of an EM iteration can share the machine); the statistics are
    # all-reduced and read back when `suff_stats` is first looked at.
    @property
    def suff_stats(self):
        if self._pending:
            self._publish()
        return self._suff

    @suff_stats.setter
    def suff_stats(self, value):
        self._suff, self._pending = value, False

    def type_sums(self):
        """K x N: per type, the sum over this shard's cells of the row-normalised counts
        (`std_row`, utils.py:34-38; the numerator of gem.py:234-238), one pass on the device."""
        out = self._torch.empty(self.K * self.N, dtype=self._torch.float64, device="cuda")
        _lib.call("ursm_sc_type_sums", self._h, ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr())
        return out.cpu().numpy().reshape(self.K, self.N)

    @property
    def _h(self):
        return self._shard.h

    # -- reference API ---------------------------------------------------
    def init_gibbs(self):
        """e_step_gibbs.py:212-225.  The chain is re-initialised on the device at
        the start of every `gibbs()` (as the reference does, :293); nothing to do."""

    def update_parameters(self, A, pkappa, ptau):
        """e_step_gibbs.py:273-280."""
        self.A = np.array(A, dtype=float, copy=True)
        self.pkappa = np.array(pkappa, dtype=float, copy=True)
        self.ptau = np.array(ptau, dtype=float, copy=True)
        self._push_params()

    def gibbs(self, burnin=100, sample=100, thin=1):
        """e_step_gibbs.py:286-304: restart the chain, `burnin` + `sample*thin`
        cycles, statistics from every `thin`-th retained cycle."""
        logging.debug("\tE-step for single cells started.")
        with _on_side_stream(self):
            _lib.call("ursm_sc_gibbs", self._h, int(burnin), int(sample), int(thin),
                      ctypes.c_uint64(self.seed), ctypes.c_uint64(self.em_iter),
                      ctypes.c_void_p(self._stats.data_ptr()), _lib.stream_ptr())
        self.em_iter += 1
        self._pending = True

    # -- parity hooks ----------------------------------------------------
    def inject(self, w, S, kappa, tau, sample, reset=False):
        """One `update_suffStats(sample)` (e_step_gibbs.py:245-270) on a supplied state."""
        w, S = _lib.f64(w), _lib.i64(S)
        kappa, tau = _lib.f64(np.ravel(kappa)), _lib.f64(np.ravel(tau))
        _lib.call("ursm_sc_inject", self._h, _lib.ptr(w), _lib.ptr(S), _lib.ptr(kappa),
                  _lib.ptr(tau), int(sample), int(bool(reset)),
                  ctypes.c_void_p(self._stats.data_ptr()), _lib.stream_ptr())
        self._publish()

    def sweep_debug(self, S, kappa, tau, sweep=0, seed=None, em_iter=0):
        """One production `gibbs_cycle` from a supplied (S, kappa, tau); returns
        (w, S_new, kappa_new, tau_new, scan_rounds)."""
        S = _lib.i64(S)
        kappa, tau = _lib.f64(np.ravel(kappa)), _lib.f64(np.ravel(tau))
        w = np.empty((self.L, self.N), dtype=np.float32)
        S_out = np.empty((self.L, self.N), dtype=np.uint8)
        k_out, t_out = np.empty(self.L), np.empty(self.L)
        rounds = np.empty(self.L, dtype=np.int32)
        _lib.call("ursm_sc_sweep_debug", self._h, _lib.ptr(S), _lib.ptr(kappa), _lib.ptr(tau),
                  ctypes.c_uint64(self.seed if seed is None else seed), ctypes.c_uint64(em_iter),
                  int(sweep), _lib.ptr(w), _lib.ptr(S_out), _lib.ptr(k_out), _lib.ptr(t_out),
                  _lib.ptr(rounds))
        return w, S_out, k_out, t_out, rounds

    def chain_debug(self, S, kappa, tau, burnin, sample, thin=1, seed=None, em_iter=0):
        """`burnin + sample*thin` production cycles from a supplied (S, kappa, tau).  Returns a
        dict: per-sweep `w` [T, L, N], `S` [T, L, N], `kappa` / `tau` [T, L], and what the
        production fold accumulated: `stats` (packed vector), `cnt` [L, N], `cellmom` [4, L]."""
        S = _lib.i64(S)
        kappa, tau = _lib.f64(np.ravel(kappa)), _lib.f64(np.ravel(tau))
        T = int(burnin) + int(sample) * int(thin)
        w = np.empty((T, self.L, self.N), dtype=np.float32)
        S_out = np.empty((T, self.L, self.N), dtype=np.uint8)
        kt = np.empty((T, 2, self.L))
        stats = np.empty(2 * self.N * self.K + 6)
        cnt = np.empty((self.L, self.N), dtype=np.uint16)
        mom = np.empty((4, self.L))
        _lib.call("ursm_sc_chain_debug", self._h, _lib.ptr(S), _lib.ptr(kappa), _lib.ptr(tau),
                  ctypes.c_uint64(self.seed if seed is None else seed), ctypes.c_uint64(em_iter),
                  int(burnin), int(sample), int(thin), _lib.ptr(w), _lib.ptr(S_out), _lib.ptr(kt),
                  _lib.ptr(stats), _lib.ptr(cnt), _lib.ptr(mom))
        return dict(w=w, S=S_out, kappa=kt[:, 0], tau=kt[:, 1], stats=stats, cnt=cnt, cellmom=mom)

    def kernel_ms(self):
        """Device time of the last Gibbs kernel (CUDA events on its launch stream)."""
        ms = ctypes.c_float()
        _lib.call("ursm_sc_kernel_ms", self._h, ctypes.byref(ms))
        return ms.value

    def geometry(self):
        v = [ctypes.c_int32() for _ in range(5)]
        _lib.load().ursm_sc_geometry(self._h, *[ctypes.byref(x) for x in v])
        return dict(zip(("threads", "genes_per_thread", "grid", "smem_bytes", "acc_in_smem"),
                        [x.value for x in v]))

    # -- internals -------------------------------------------------------
    def _push_params(self):
        A, pk, pt = _lib.f64(self.A), _lib.f64(self.pkappa), _lib.f64(self.ptau)
        _lib.call("ursm_sc_set_params", self._h, _lib.ptr(A), _lib.ptr(pk), _lib.ptr(pt))

    def _publish(self):
        """All-reduce the packed statistics and expose them under the reference's keys."""
        N, K, L = self.N, self.K, self.L
        self._pending = False
        _join_side_stream(self)
        dist.all_reduce_(self._stats)  # the ONE collective of the single-cell E-step
        host = self._stats.cpu().numpy()
        mom = [np.empty(L) for _ in range(4)]
        _lib.call("ursm_sc_get_cell_moments", self._h, *[_lib.ptr(m) for m in mom])
        st = {}
        st["exp_S"] = DeviceExpS(self)
        st["exp_kappa"] = mom[0].reshape(1, L)
        st["exp_tau"] = mom[1].reshape(1, L)
        st["exp_kappasq"] = mom[2].reshape(1, L)
        st["exp_tausq"] = mom[3].reshape(1, L)
        st["coeffA"] = np.ascontiguousarray(host[:N * K].reshape(K, N).T)
        st["coeffAsq"] = np.ascontiguousarray(host[N * K:2 * N * K].reshape(K, N).T)
        st["exp_elbo_const"] = float(host[2 * N * K])
        if host[2 * N * K + 5] != 0:
            raise _lib.UrsmError("single-cell Gibbs kernel: the draw_S scan did not reach its "
                                 "fixed point for %d (cell, sweep) pairs" % int(host[2 * N * K + 5]))
        # hidden hand-over to LogitNormalMLE: device buffers + global scalar sums
        st["_ursm_sc"] = weakref.ref(self)
        st["_ursm_sc_sums"] = host[2 * N * K + 1:2 * N * K + 5].copy()  # sum k, k^2, t, t^2
        self._suff = st


class LogitNormalGibbs_BK(object):
    """Bulk-sample Gibbs sampler / mean-update (e_step_gibbs.py:19-164)."""

    def __init__(self, A, alpha, BKexpr, iMarkers=None, row_offset=0, total_rows=None,
                 seed=None, device_counts=None, K=None):
        """`A=None` (with `K`): see LogitNormalGibbs_SC."""
        torch = _lib.require_cuda()
        self._torch = torch
        self.BKexpr, self.iMarkers = BKexpr, iMarkers
        data = BKexpr if BKexpr is not None else device_counts
        self.M = int(data.shape[0])
        (self.N, self.K) = A.shape if A is not None else (int(data.shape[1]), int(K))
        self.row_offset = int(row_offset)
        self.total_rows = int(total_rows) if total_rows is not None else self.M
        self.A = np.array(A, dtype=float, copy=True) if A is not None else None
        self.alpha = np.array(alpha, dtype=float, copy=True) if alpha is not None else None
        self.seed = _DEFAULT_SEED if seed is None else int(seed)
        self.em_iter = 0
        self.freeze = 0  # parity hook: bit0 keep W, bit1 keep sum_i Z
        h = ctypes.c_void_p()
        if device_counts is not None:
            self._keep = device_counts
            _lib.call("ursm_bk_create_dev", ctypes.byref(h), ctypes.c_void_p(device_counts.data_ptr()),
                      self.M, self.N, self.K, self.row_offset)
        else:
            X = _lib.f64(BKexpr)
            _lib.call("ursm_bk_create", ctypes.byref(h), _lib.ptr(X), self.M, self.N, self.K,
                      self.row_offset)
        self._shard = _Shard(h, "ursm_bk_destroy")
        n = _lib.load().ursm_bk_stats_len(h)
        self._stats = torch.zeros(n, dtype=torch.float64, device="cuda")
        if self.A is not None:
            self._push_params()
        self._suff, self._pending, self._stream = {}, False, None

    @property
    def suff_stats(self):
        """See LogitNormalGibbs_SC.suff_stats: materialised on first access after `gibbs()`."""
        if self._pending:
            self._publish()
        return self._suff

    @suff_stats.setter
    def suff_stats(self, value):
        self._suff, self._pending = value, False

    def profile_sum(self):
        """N-vector: the sum over this shard's bulk samples of the row-normalised counts
        (the numerator of gem.py:228-230 / :246), one pass on the device."""
        out = self._torch.empty(self.N, dtype=self._torch.float64, device="cuda")
        _lib.call("ursm_bk_profile_sum", self._h, ctypes.c_void_p(out.data_ptr()),
                  _lib.stream_ptr())
        return out.cpu().numpy()

    @property
    def _h(self):
        return self._shard.h

    def init_gibbs(self):
        """e_step_gibbs.py:39-55: W = 1/K, or seeded from the marker genes (needs the
        host copy of the marker columns only)."""
        M, K = self.M, self.K
        if self.iMarkers is not None:
            logging.debug("\t\tZ is initialized with Marker info.")
            if self.BKexpr is None:
                raise _lib.UrsmError("marker initialisation needs the host copy of BKexpr "
                                     "(the marker columns); it was built from device_BK")
            Zjk = np.zeros([M, K])
            # the reference ASSIGNS Z[:, i, k] per marker row (:50-51) and then sums over genes:
            # a (gene, type) pair listed twice counts once
            for (i, k) in sorted(set((int(r[0]), int(r[1])) for r in np.asarray(self.iMarkers))):
                Zjk[:, k] += self.BKexpr[:, i] + self.alpha[k]
            W = Zjk.transpose().copy()
            W /= W.sum(axis=0)[np.newaxis, :]
        else:
            Zjk = None
            W = np.full([K, M], 1.0 / K, dtype=float)
        self.set_state(W, Zjk)

    def set_state(self, W, Zjk=None):
        W = _lib.f64(W)
        Zjk = None if Zjk is None else _lib.f64(Zjk)
        _lib.call("ursm_bk_set_state", self._h, _lib.ptr(W), _lib.ptr(Zjk))

    @property
    def W(self):
        W = np.empty((self.K, self.M))
        _lib.call("ursm_bk_get_W", self._h, _lib.ptr(W))
        return W

    def update_parameters(self, A, alpha):
        """e_step_gibbs.py:81-84."""
        self.A = np.array(A, dtype=float, copy=True)
        self.alpha = np.array(alpha, dtype=float, copy=True)
        self._push_params()

    def gibbs(self, burnin=100, sample=100, thin=1, mean_approx=True):
        """e_step_gibbs.py:90-117."""
        sp = ctypes.c_void_p(self._stats.data_ptr())
        with _on_side_stream(self):
            if mean_approx:
                logging.debug("\tE-step (one-step mean-update) for bulk samples started.")
                _lib.call("ursm_bk_mean_step", self._h, sp, _lib.stream_ptr())
            else:
                logging.debug("\tE-step for bulk samples started.")
                _lib.call("ursm_bk_gibbs", self._h, int(burnin), int(sample), int(thin),
                          ctypes.c_uint64(self.seed), ctypes.c_uint64(self.em_iter),
                          int(self.freeze), sp, _lib.stream_ptr())
        self.em_iter += 1
        self._pending = True

    def kernel_ms(self):
        """Device time of the last bulk E-step's kernels."""
        ms = ctypes.c_float()
        _lib.call("ursm_bk_kernel_ms", self._h, ctypes.byref(ms))
        return ms.value

    def _push_params(self):
        A, al = _lib.f64(self.A), _lib.f64(self.alpha)
        _lib.call("ursm_bk_set_params", self._h, _lib.ptr(A), _lib.ptr(al))

    def _publish(self):
        N, K, M = self.N, self.K, self.M
        self._pending = False
        _join_side_stream(self)
        dist.all_reduce_(self._stats)  # the ONE collective of the bulk E-step
        host = self._stats.cpu().numpy()
        eZjk, eLogW, eW = np.empty((M, K)), np.empty((K, M)), np.empty((K, M))
        _lib.call("ursm_bk_get_sample_stats", self._h, _lib.ptr(eZjk), _lib.ptr(eLogW),
                  _lib.ptr(eW))
        st = {}
        st["exp_Zik"] = np.ascontiguousarray(host[:N * K].reshape(K, N).T)
        st["exp_Zjk"] = eZjk
        st["exp_logW"] = eLogW
        st["exp_W"] = eW
        st["_ursm_bk"] = weakref.ref(self)
        st["_ursm_bk_tail"] = host[N * K:N * K + K + 1].copy()  # sum_j E log W_k | sum Zjk*logW
        if host[N * K + K + 1] != 0:
            # (the cold kernel adds 1 << 20 per sweep whose hand-over list did not fit, the
            # Dirichlet kernel 1 per gamma draw that ran out of attempts; neither has been seen)
            bad = int(host[N * K + K + 1])
            raise _lib.UrsmError("bulk Gibbs kernels: %d sweeps overflowed the hand-over list of the "
                                 "split kernel, %d gamma draws ran out of attempts"
                                 % (bad >> 20, bad & ((1 << 20) - 1)))
        self._suff = st
