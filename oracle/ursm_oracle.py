"""CPU oracle for the URSM Monte-Carlo EM hot path.  TEST INFRASTRUCTURE ONLY.

This module is a numpy restatement of the algorithm in the reference
(`/root/reference/e_step_gibbs.py`, `m_step.py`, `utils.py:8-31`,
`gem.py:135-195`).  It exists so that `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` leg have something to check
and time the CUDA path against.  Nothing under `ursm_b200/` may import it: the
product path is CUDA-only and fails loudly without its extension.

Pinning (see `tests/golden/make_golden.py`): every class here is compared with
the *actual* reference modules executed in the build container (Py3 edits
applied in memory, nothing copied) on seeded inputs; the outputs are committed
under `tests/golden/*.npz`.  The deterministic paths (bulk mean-approx E-step,
`update_suffStats`, the whole M-step, `simplex_proj`, ELBO) are pinned to
~1e-12.  The stochastic sweeps consume `numpy.random` in exactly the reference's
call order, so seeded reference runs are reproduced to the last bit as well.

The one piece the reference does not contain is the Polya-Gamma sampler itself
(`pypolyagamma 1.1.1`, an un-vendored C++ dependency, `README.md:22`;
call sites `e_step_gibbs.py:12,195,204,209,323`).  `pg1_devroye` restates its
published algorithm (Polson, Scott & Windle 2013, sec. 4; Devroye 2009) and is
validated against the closed-form PG(1,z) moments and Laplace transform, not
against bit patterns: **parity for the PG draw itself is distributional
("unpinned")**; everything around it is pinned.
"""
import math

import numpy as np
from scipy.special import expit, gammaln, psi as digamma

# --------------------------------------------------------------------------
# utils.py:8-31  --  projection onto {y >= min_y, sum(y) = 1}
# --------------------------------------------------------------------------


def simplex_proj(y, min_y):
    """Sort-based projection (Wang 2013) exactly as `utils.py:8-31` does it.

    The reference walks the descending-sorted vector with a scalar running sum;
    `np.cumsum` performs the same left-to-right additions, so the result is
    bit-identical to the reference loop.
    """
    y = np.asarray(y, dtype=float)
    p = y.shape[0]
    srt = np.sort(y)[::-1]
    csum = np.cumsum(srt)
    idx = np.arange(p)
    rou = srt + (1 - (p - idx - 1) * min_y - csum) / (idx + 1.0)
    ok = np.nonzero(rou > min_y)[0]
    if ok.size == 0:  # reference falls through with iBest=-1, rouBest=-1
        lamb = -1 - srt[-1]
    else:
        best = ok[-1]
        lamb = rou[best] - srt[best]
    return np.maximum(y + lamb, min_y)


def std_row(B):
    """`utils.py:34-38`: rows scaled to sum to one."""
    return np.true_divide(B, B.sum(axis=1)[:, np.newaxis])


# --------------------------------------------------------------------------
# PG(1, z): Devroye alternating-series sampler (pypolyagamma 1.1.1 algorithm)
# --------------------------------------------------------------------------

PG_TRUNC = 0.64
_HALF_PI2 = 0.125 * math.pi * math.pi


def _log_ndtr(x):
    """log Phi(x), stable in both tails."""
    if x > -5.0:
        return math.log(0.5 * math.erfc(-x / math.sqrt(2.0)))
    # Mills-ratio expansion for the far lower tail
    x2 = x * x
    return (-0.5 * x2 - math.log(-x) - 0.5 * math.log(2 * math.pi)
            + math.log1p(-1 / x2 + 3 / x2 ** 2 - 15 / x2 ** 3 + 105 / x2 ** 4))


def pg_mass_texpon(z):
    """P(proposal comes from the exponential tail) for tilt z (= |psi|/2)."""
    t = PG_TRUNC
    fz = _HALF_PI2 + 0.5 * z * z
    b = math.sqrt(1.0 / t) * (t * z - 1)
    a = -math.sqrt(1.0 / t) * (t * z + 1)
    x0 = math.log(fz) + fz * t
    xb = x0 - z + _log_ndtr(b)
    xa = x0 + z + _log_ndtr(a)
    # q/p can overflow for large z -> mass 0
    try:
        qdivp = 4 / math.pi * (math.exp(xb) + math.exp(xa))
    except OverflowError:
        return 0.0
    return 1.0 / (1.0 + qdivp)


def _pg_coef(n, x):
    """a_n(x) of the Jacobi alternating series."""
    k = (n + 0.5) * math.pi
    if x > PG_TRUNC:
        return k * math.exp(-0.5 * k * k * x)
    if x > 0:
        return math.exp(-1.5 * (math.log(0.5 * math.pi) + math.log(x))
                        + math.log(k) - 2.0 * (n + 0.5) ** 2 / x)
    return 0.0


def _rtigauss(z, rs):
    """Inverse-Gaussian(1/z, 1) truncated to (0, PG_TRUNC)."""
    t = PG_TRUNC
    if 1.0 / t > z:
        while True:
            e1, e2 = rs.exponential(), rs.exponential()
            while e1 * e1 > 2 * e2 / t:
                e1, e2 = rs.exponential(), rs.exponential()
            x = 1 + e1 * t
            x = t / (x * x)
            if rs.random_sample() <= math.exp(-0.5 * z * z * x):
                return x
    mu = 1.0 / z
    while True:
        y = rs.normal() ** 2
        mu_y = mu * y
        x = mu + 0.5 * mu * mu_y - 0.5 * mu * math.sqrt(4 * mu_y + mu_y * mu_y)
        if rs.random_sample() > mu / (mu + x):
            x = mu * mu / x
        if x <= t:
            return x


def pg1_devroye(psi, rs):
    """One draw of PG(1, psi) using RandomState-like `rs`."""
    z = abs(psi) * 0.5
    fz = _HALF_PI2 + 0.5 * z * z
    mass = pg_mass_texpon(z)
    while True:
        if rs.random_sample() < mass:
            x = PG_TRUNC + rs.exponential() / fz
        else:
            x = _rtigauss(z, rs)
        s = _pg_coef(0, x)
        y = rs.random_sample() * s
        n = 0
        while True:
            n += 1
            if n % 2 == 1:
                s -= _pg_coef(n, x)
                if y <= s:
                    return 0.25 * x
            else:
                s += _pg_coef(n, x)
                if y > s:
                    break


def pg1_mean(psi):
    """E[PG(1,psi)] = tanh(psi/2)/(2 psi)."""
    psi = np.asarray(psi, dtype=float)
    small = np.abs(psi) < 1e-6
    safe = np.where(small, 1.0, psi)
    return np.where(small, 0.25 - psi * psi / 48.0, np.tanh(safe / 2) / (2 * safe))


def pg1_var(psi):
    """Var[PG(1,psi)] = (sinh psi - psi)/(4 psi^3 cosh^2(psi/2))."""
    psi = np.abs(np.asarray(psi, dtype=float))
    small = psi < 1e-2
    safe = np.where(small, 1.0, psi)
    # overflow-free form: (sinh p - p)/cosh^2(p/2) = 2 (sinh p - p)/(1+cosh p)
    with np.errstate(over="ignore"):
        big = safe > 30
        core = np.where(big, 2.0 * (1 - 2 * safe * np.exp(-np.minimum(safe, 700))),
                        2 * (np.sinh(np.minimum(safe, 30)) - np.minimum(safe, 30))
                        / (1 + np.cosh(np.minimum(safe, 30))))
    return np.where(small, 1 / 24.0 - psi ** 2 / 160.0, core / (4 * safe ** 3))


class PG1Sampler(object):
    """Stand-in for `pypolyagamma.PyPolyaGamma(seed)` (e_step_gibbs.py:209)."""

    def __init__(self, seed=0):
        self.rs = np.random.RandomState(int(seed))

    def draw(self, psi):
        return pg1_devroye(float(psi), self.rs)


def pgdrawv(samplers, ns, zs, out):
    """Stand-in for `pypolyagamma.pgdrawvpar` (e_step_gibbs.py:323): in-place.

    The C++ original splits entries over OpenMP threads; the oracle walks them in
    order with `samplers[0]` (one deterministic stream)."""
    s = samplers[0]
    for i in range(len(zs)):
        out[i] = s.draw(zs[i])


# ---- C build of the same sampler (oracle/pg_devroye.c, OpenMP) -------------
# pypolyagamma is C++/OpenMP in the reference's environment; timing the pure-Python
# restatement above as "the reference's CPU path" would be unfair to it, so the
# cpu_baseline leg of bench.py uses this compiled restatement through the same
# `pgdrawvpar`-shaped hook (one generator per OpenMP thread, e_step_gibbs.py:200-209).

_PG_C = None


def _pg_c():
    global _PG_C
    if _PG_C is None:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build",
                            "libpg_devroye.so")
        if not os.path.exists(path):
            from . import build as _b
            _b.build_pg()
        lib = ctypes.CDLL(path)
        lib.pg_pool_create.restype = ctypes.c_void_p
        lib.pg_pool_create.argtypes = [ctypes.c_int, ctypes.c_void_p]
        lib.pg_pool_destroy.argtypes = [ctypes.c_void_p]
        lib.pg_drawv.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                 ctypes.c_void_p]
        lib.pg_mass_texpon.restype = ctypes.c_double
        lib.pg_mass_texpon.argtypes = [ctypes.c_double]
        lib.pg_max_threads.restype = ctypes.c_int
        _PG_C = lib
    return _PG_C


def pg_c_max_threads():
    return int(_pg_c().pg_max_threads())


class CPGSampler(object):
    """`PyPolyaGamma(seed)` stand-in backed by pg_devroye.c (one per OpenMP thread)."""

    def __init__(self, seed=0):
        self.seed = int(seed)


def pgdrawv_c(samplers, ns, zs, out):
    """`pgdrawvpar(ppgs, ns, zs, out)` stand-in: len(samplers) OpenMP threads, in place."""
    import ctypes
    lib = _pg_c()
    pool = getattr(samplers[0], "_pool", None)
    if pool is None:
        seeds = np.array([s.seed for s in samplers], dtype=np.uint64)
        pool = lib.pg_pool_create(len(samplers), seeds.ctypes.data_as(ctypes.c_void_p))
        samplers[0]._pool = pool
    zs = np.ascontiguousarray(zs, dtype=np.float64)
    if out.flags["C_CONTIGUOUS"] and out.dtype == np.float64:
        lib.pg_drawv(pool, zs.size, zs.ctypes.data_as(ctypes.c_void_p),
                     out.ctypes.data_as(ctypes.c_void_p))
    else:
        tmp = np.empty(zs.size)
        lib.pg_drawv(pool, zs.size, zs.ctypes.data_as(ctypes.c_void_p),
                     tmp.ctypes.data_as(ctypes.c_void_p))
        out[:] = tmp


# --------------------------------------------------------------------------
# e_step_gibbs.py:19-164  --  bulk samples
# --------------------------------------------------------------------------


class OracleGibbsBK(object):
    """Restates `LogitNormalGibbs_BK` (e_step_gibbs.py:19-164)."""

    def __init__(self, A, alpha, BKexpr, iMarkers=None, rng=np.random):
        self.X = BKexpr
        self.iMarkers = iMarkers
        self.M, self.N, self.K = BKexpr.shape[0], A.shape[0], A.shape[1]
        self.A = np.array(A, dtype=float, copy=True)
        self.alpha = np.array(alpha, dtype=float, copy=True)
        self.rng = rng

    def init_gibbs(self):  # :39-55
        M, N, K = self.M, self.N, self.K
        self.Z = np.zeros([M, N, K])
        if self.iMarkers is not None:
            for gene, k in self.iMarkers:
                self.Z[:, gene, k] = self.X[:, gene] + self.alpha[k]
            W = self.Z.sum(axis=1).T
            self.W = W / W.sum(axis=0)[np.newaxis, :]
        else:
            self.W = np.full([K, M], 1.0 / K)

    def update_parameters(self, A, alpha):  # :81-84
        self.A = np.array(A, dtype=float, copy=True)
        self.alpha = np.array(alpha, dtype=float, copy=True)

    def _zero_stats(self):  # :59-66
        self.suff_stats = dict(exp_Zik=np.zeros([self.N, self.K]),
                               exp_Zjk=np.zeros([self.M, self.K]),
                               exp_logW=np.zeros([self.K, self.M]),
                               exp_W=np.zeros([self.K, self.M]))

    def accumulate(self, sample):  # :69-78
        st = self.suff_stats
        st["exp_Zik"] += self.Z.sum(axis=0) / float(sample)
        st["exp_Zjk"] += self.Z.sum(axis=1) / float(sample)
        st["exp_logW"] += np.log(self.W) / float(sample)
        st["exp_W"] += self.W / float(sample)

    def nmf_W(self):  # :157-164  (every k uses the AW computed before the loop)
        AW = np.dot(self.A, self.W)
        self.AW = AW
        for k in range(self.K):
            mult = self.X * self.A[:, k].T / AW.T
            self.W[k, :] = self.W[k, :] * mult.sum(axis=1) + self.alpha[k] - 1
        self.W /= self.W.sum(axis=0)[np.newaxis, :]

    def mean_Z(self):  # :149-155  E[Z | W, A, X]
        self.AW = np.dot(self.A, self.W)
        pval = self.W.T[:, np.newaxis, :] * self.A[np.newaxis, :, :]
        self.Z = pval * self.X[:, :, np.newaxis] / self.AW.T[:, :, np.newaxis]

    def draw_W(self):  # :140-146
        post = self.Z.sum(axis=1)
        for j in range(self.M):
            self.W[:, j] = self.rng.dirichlet(self.alpha + post[j, :])
        self.AW = np.dot(self.A, self.W)

    def draw_Z(self):  # :132-138
        for j in range(self.M):
            for i in range(self.N):
                pval = self.W[:, j] * self.A[i, :]
                self.Z[j, i, :] = self.rng.multinomial(n=self.X[j, i],
                                                       pvals=pval / self.AW[i, j])

    def gibbs(self, burnin=100, sample=100, thin=1, mean_approx=True):  # :90-117
        self._zero_stats()
        if mean_approx:
            self.nmf_W()
            self.mean_Z()
            self.accumulate(1)
            return
        for _ in range(burnin):
            self.draw_W()
            self.draw_Z()
        for it in range(sample * thin):
            self.draw_W()
            self.draw_Z()
            if it % thin == 0:
                self.accumulate(sample)


# --------------------------------------------------------------------------
# e_step_gibbs.py:171-377  --  single cells
# --------------------------------------------------------------------------


def sc_kappa_tau_posterior(w_row, S_row, a_col, sum_AS, N, pkappa, ptau):
    """Mean and covariance of (kappa_l, tau_l) | w, S  (e_step_gibbs.py:366-375).

    Returns (mP[2], sigma[2,2]).  Uses the builtin sequential `sum` like the
    reference so the additions happen in the same order."""
    offdiag = sum(w_row * a_col)
    d0 = sum(w_row) + 1.0 / pkappa[1]
    d1 = sum(w_row * np.square(a_col)) + 1.0 / ptau[1]
    det = d0 * d1 - offdiag * offdiag
    sigma = np.array([[d1, -offdiag], [-offdiag, d0]]) / det
    bP = np.array([sum(S_row) - N / 2.0 + pkappa[0] / pkappa[1],
                   sum_AS - 0.5 + ptau[0] / ptau[1]])
    return np.dot(sigma, bP), sigma


class OracleGibbsSC(object):
    """Restates `LogitNormalGibbs_SC` (e_step_gibbs.py:171-377)."""

    def __init__(self, A, pkappa, ptau, SCexpr, G, itype, rng=np.random,
                 num_threads=1, pg_factory=PG1Sampler, pg_draw=pgdrawv):
        self.Y, self.G, self.L = SCexpr, G, SCexpr.shape[0]
        self.N, self.K = A.shape
        self.R = SCexpr.sum(axis=1)
        self.itype = itype
        self.A = np.array(A, dtype=float, copy=True)
        self.pkappa = np.array(pkappa, dtype=float, copy=True)
        self.ptau = np.array(ptau, dtype=float, copy=True)
        self.izero = np.where(self.Y == 0)
        self.rng = rng
        # the reference burns one randint(2**16, size=num_threads) call (:196)
        # and uses a second one (:208) for the PG seeds
        rng.randint(2 ** 16, size=num_threads)
        seeds = rng.randint(2 ** 16, size=num_threads)
        self.samplers = [pg_factory(s) for s in seeds]
        self.pg_draw = pg_draw

    # -- state ---------------------------------------------------------
    def init_gibbs(self):  # :212-225
        L, N = self.L, self.N
        self.kappa = np.full([1, L], self.pkappa[0], dtype=float)
        self.tau = np.full([1, L], self.ptau[0], dtype=float)
        self.S = np.reshape(self.rng.binomial(1, 0.5, size=L * N), [L, N])
        self.refresh_psi()
        self.w = np.ones([L, N])
        self.S[np.where(self.Y > 0)] = 1
        self.sum_AS = (self.A[:, self.G].T * self.S).sum(axis=1)

    def refresh_psi(self):  # :314-316
        self.psi = np.transpose(self.kappa + self.tau * self.A[:, self.G])

    def update_parameters(self, A, pkappa, ptau):  # :273-280
        self.A = np.array(A, dtype=float, copy=True)
        self.pkappa = np.array(pkappa, dtype=float, copy=True)
        self.ptau = np.array(ptau, dtype=float, copy=True)
        self.refresh_psi()
        self.sum_AS = (self.A[:, self.G].T * self.S).sum(axis=1)

    def _zero_stats(self):  # :228-242
        L, N, K = self.L, self.N, self.K
        self.suff_stats = dict(exp_S=np.zeros([L, N]),
                               exp_kappa=np.zeros([1, L]), exp_tau=np.zeros([1, L]),
                               exp_kappasq=np.zeros([1, L]), exp_tausq=np.zeros([1, L]),
                               coeffA=np.zeros([N, K]), coeffAsq=np.zeros([N, K]),
                               exp_elbo_const=0)

    # -- one retained sample -> statistics -------------------------------
    def accumulate(self, sample):  # :245-270
        st = self.suff_stats
        st["exp_S"] += self.S / float(sample)
        st["exp_kappa"] += self.kappa / sample
        st["exp_tau"] += self.tau / sample
        st["exp_kappasq"] += np.square(self.kappa) / sample
        st["exp_tausq"] += np.square(self.tau) / sample
        st["exp_elbo_const"] += (-self.w * np.transpose(np.square(self.kappa))).sum() \
            / (2.0 * sample)
        st["exp_elbo_const"] += ((self.S - 0.5) * np.transpose(self.kappa)).sum() / sample
        cA = (self.S - 0.5) * np.transpose(self.tau) \
            - self.w * np.transpose(self.tau * self.kappa)
        cAsq = (-self.w * np.transpose(np.square(self.tau))) / 2.0
        for k in range(self.K):
            rows = self.itype[k]
            if len(rows) > 0:
                st["coeffA"][:, k] += cA[rows, :].sum(axis=0) / float(sample)
                st["coeffAsq"][:, k] += cAsq[rows, :].sum(axis=0) / float(sample)

    # -- the three conditional draws -------------------------------------
    def draw_w(self):  # :318-323
        ns = np.ones(self.N, dtype=float)
        for l in range(self.L):
            self.pg_draw(self.samplers, ns, self.psi[l, :], self.w[l, :])

    def S_prob(self, l, i):
        """P(S_li = 1 | rest) used by draw_S (:331-338); also the moment oracle."""
        a = self.A[i, self.G[l]]
        other = self.sum_AS[l] - a * self.S[l, i]
        if other == 0:
            return expit(self.psi[l][i]), other, a
        return expit(self.psi[l][i] - self.R[l] * np.log(1 + a / other)), other, a

    def draw_S(self):  # :326-342, row-major over the zero entries
        rows, cols = self.izero
        for n in range(len(rows)):
            l, i = rows[n], cols[n]
            b, other, a = self.S_prob(l, i)
            self.S[l][i] = self.rng.binomial(1, b)
            self.sum_AS[l] = other + a * self.S[l, i]

    def draw_kappa_tau(self):  # :345-377
        for l in range(self.L):
            mP, sigma = sc_kappa_tau_posterior(self.w[l, :], self.S[l, :],
                                               self.A[:, self.G[l]], self.sum_AS[l],
                                               self.N, self.pkappa, self.ptau)
            self.kappa[0, l], self.tau[0, l] = self.rng.multivariate_normal(mP, sigma)

    def cycle(self):  # :307-312
        self.draw_w()
        self.draw_S()
        self.draw_kappa_tau()
        self.refresh_psi()

    def gibbs(self, burnin=100, sample=100, thin=1):  # :286-304
        self._zero_stats()
        self.init_gibbs()
        for _ in range(burnin):
            self.cycle()
        for it in range(sample * thin):
            self.cycle()
            if it % thin == 0:
                self.accumulate(sample)


# --------------------------------------------------------------------------
# m_step.py:13-320
# --------------------------------------------------------------------------


class OracleMLE(object):
    """Restates `LogitNormalMLE` (m_step.py:13-320)."""

    def __init__(self, BKexpr=None, SCexpr=None, G=None, K=3, itype=None,
                 hasBK=True, hasSC=True, init_A=None, init_alpha=None,
                 init_pkappa=None, init_ptau=None, min_A=1e-6, min_alpha=1,
                 MLE_CONV=1e-4, MLE_maxiter=100):
        self.K, self.min_A, self.min_alpha = K, min_A, min_alpha
        self.MLE_CONV, self.MLE_maxiter = MLE_CONV, MLE_maxiter
        self.Y, self.G, self.X = SCexpr, G, BKexpr
        self.hasBK, self.hasSC, self.itype = hasBK, hasSC, itype
        if hasSC:
            self.L, self.N = SCexpr.shape
            self.R = SCexpr.sum(axis=1)
        if hasBK:
            self.M, self.N = BKexpr.shape
        self.A = np.copy(init_A)
        self.alpha = np.copy(init_alpha)
        self.pkappa = np.copy(init_pkappa)
        self.ptau = np.copy(init_ptau)
        self.niter_A = [0] * K

    # :50-89
    def update_suff_stats(self, st):
        self.suff_stats = st
        shape = self.A.shape
        self.gd_coeffAinv = np.zeros(shape)
        self.gd_coeffA = np.zeros(shape)
        self.gd_coeffConst = np.zeros(shape)
        self.elbo_const = 0
        if self.hasBK:
            self.elbo_const += (st["exp_Zjk"].T * st["exp_logW"]).sum()
            self.gd_coeffAinv += st["exp_Zik"]
            st["exp_mean_logW"] = st["exp_logW"].mean(axis=1)
        if self.hasSC:
            self.elbo_const += st["exp_elbo_const"]
            self.opt_u()
            self.gd_coeffA = st["coeffAsq"] * 2.0
            for k in range(self.K):
                rows = self.itype[k]
                if len(rows) > 0:
                    self.refresh_coeffConst(k)
                    self.gd_coeffAinv[:, k] += (self.Y[rows, :] * st["exp_S"][rows, :]).sum(axis=0)

    def opt_kappa_tau(self):  # :92-103
        st = self.suff_stats
        km, tm = np.mean(st["exp_kappa"]), np.mean(st["exp_tau"])
        self.pkappa = np.array([km, np.mean(st["exp_kappasq"]) - km ** 2])
        self.ptau = np.array([tm, np.mean(st["exp_tausq"]) - tm ** 2])

    def opt_u(self):  # :108-115
        self.u = (self.A[:, self.G].T * self.suff_stats["exp_S"]).sum(axis=1)

    def refresh_coeffConst(self, k):  # :118-125
        rows = self.itype[k]
        st = self.suff_stats
        self.gd_coeffConst[:, k] = st["coeffA"][:, k] - \
            (st["exp_S"][rows, :] * (self.R / self.u)[rows, np.newaxis]).sum(axis=0)

    def opt_A_u(self):  # :128-154
        if self.hasBK and self.hasSC:
            step0 = 0.01 / (self.L + self.M)
        elif self.hasBK:
            step0 = 0.01 / self.M
        else:
            step0 = 0.01 / self.L
        for k in range(self.K):
            if self.hasSC and len(self.itype[k]) > 0:
                self.niter_A[k] = self.opt_Ak(k, step0)
            else:
                col = self.suff_stats["exp_Zik"][:, k]
                self.A[:, k] = col / float(np.sum(col))
                self.A[:, k] = simplex_proj(self.A[:, k], self.min_A)
                self.niter_A[k] = 0

    def opt_Ak(self, k, step0):  # :157-182
        old = np.copy(self.A[:, k])
        conv, niter = self.MLE_CONV + 1, 0
        while conv > self.MLE_CONV and niter < self.MLE_maxiter:
            self.A[:, k], _, _ = self.backtracking(
                old, lambda v: -self.grad_A(v, k), lambda v: -self.obj_A(v, k),
                step0, lambda v: simplex_proj(v, self.min_A))
            self.opt_u()
            self.refresh_coeffConst(k)
            niter += 1
            conv = np.linalg.norm(self.A[:, k] - old, 1)
            old = np.copy(self.A[:, k])
        return niter

    def opt_alpha(self):  # :185-206
        conv, niter = self.MLE_CONV + 1, 0
        while conv > self.MLE_CONV and niter < self.MLE_maxiter:
            old = np.copy(self.alpha)
            self.alpha = self.backtracking(old, lambda a: -self.grad_alpha(a),
                                           lambda a: -self.obj_alpha(a), 10)[0]
            self.alpha = np.maximum(self.min_alpha, self.alpha)
            niter += 1
            conv = np.linalg.norm(self.alpha - old)
        return niter

    @staticmethod
    def backtracking(old, grad_f, obj_f, step0=0.1, proj=None):  # :209-233
        g_old = grad_f(old)
        f_old = obj_f(old)
        step = step0

        def trial(s):
            v = old - s * g_old
            return proj(v) if proj is not None else v

        new = trial(step)
        Gt = (old - new) / step
        f_new = obj_f(new)
        while f_new > f_old + (step * 0.5) * (Gt ** 2).sum() - step * np.dot(g_old, Gt):
            step = step * 0.5
            new = trial(step)
            Gt = (old - new) / step
            f_new = obj_f(new)
        return new, f_new, step

    def grad_alpha(self, a):  # :236-241
        return self.suff_stats["exp_mean_logW"] + digamma(sum(a)) - digamma(a)

    def obj_alpha(self, a):  # :244-251
        f = np.dot(self.suff_stats["exp_mean_logW"], a)
        f += gammaln(sum(a)) - sum(gammaln(a))  # same association as the reference
        return f

    def grad_A(self, Ak, k):  # :262-272
        g = self.gd_coeffAinv[:, k] / Ak + self.gd_coeffA[:, k] * Ak + self.gd_coeffConst[:, k]
        return g / self.N

    def obj_A(self, Ak, k):  # :275-287
        f = (self.gd_coeffAinv[:, k] * np.log(Ak)).sum()
        if self.hasSC:
            f += (self.gd_coeffA[:, k] * np.square(Ak)).sum() / 2.0
            f += (self.gd_coeffConst[:, k] * Ak).sum()
        return f / self.N

    def compute_elbo(self):  # :290-320
        st = self.suff_stats
        elbo = self.elbo_const
        elbo += (self.gd_coeffAinv * np.log(self.A)).sum()
        if self.hasBK:
            elbo += self.obj_alpha(self.alpha) * self.M
        if self.hasSC:
            elbo += (self.gd_coeffA * np.square(self.A)).sum() / 2.0
            elbo += (self.gd_coeffConst * self.A).sum()
            elbo -= np.sum(self.R * np.log(self.u))
            elbo -= np.log(self.pkappa[1] * self.ptau[1]) * self.L / 2.0
            elbo -= (np.sum(st["exp_kappasq"]) - 2 * self.pkappa[0] * np.sum(st["exp_kappa"])
                     + self.L * self.pkappa[0] ** 2) / (2.0 * self.pkappa[1])
            elbo -= (np.sum(st["exp_tausq"]) - 2 * self.ptau[0] * np.sum(st["exp_tau"])
                     + self.L * self.ptau[0] ** 2) / (2.0 * self.ptau[1])
        return elbo / self.N


# --------------------------------------------------------------------------
# gem.py:69-295  --  EM driver (initialisation + loop)
# --------------------------------------------------------------------------


def init_A(SCexpr, BKexpr, K, itype, min_A, user_A=None):
    """`gem.py:225-255`."""
    hasSC, hasBK = SCexpr is not None, BKexpr is not None
    N = (SCexpr if hasSC else BKexpr).shape[1]
    if hasSC:
        stdY = std_row(SCexpr)
    if hasBK:
        bkmean = std_row(BKexpr).mean(axis=0)
    if user_A is not None:
        A = user_A
    elif hasSC:
        A = np.ones([N, K]) / N
        for k in range(K):
            if len(itype[k]) > 0:
                A[:, k] = stdY[itype[k], :].mean(axis=0)
            elif hasBK:
                A[:, k] = bkmean
    else:
        A = np.zeros([N, K])
        for k in range(K):
            A[:, k] = bkmean
    for k in range(K):
        A[:, k] = simplex_proj(A[:, k], min_A)
    return A


class OracleGEM(object):
    """Restates `LogitNormalGEM` (gem.py:69-295) on top of the oracle classes."""

    def __init__(self, BKexpr=None, SCexpr=None, K=3, G=None, iMarkers=None,
                 init_A_=None, min_A=1e-6, init_alpha=None, est_alpha=True,
                 init_pkappa=None, init_ptau=None, burnin=200, sample=200, thin=1,
                 bk_mean_approx=True, MLE_CONV=1e-6, MLE_maxiter=100,
                 EM_CONV=1e-3, EM_maxiter=100, rng=np.random,
                 pg_factory=PG1Sampler, pg_draw=pgdrawv):
        self.hasBK, self.hasSC = BKexpr is not None, SCexpr is not None
        self.Y, self.G, self.X = SCexpr, G, BKexpr
        self.K, self.min_A, self.est_alpha = K, min_A, est_alpha
        self.iMarkers = iMarkers
        self.EM_CONV, self.EM_maxiter = EM_CONV, EM_maxiter
        self.MLE_CONV, self.MLE_maxiter = MLE_CONV, MLE_maxiter
        self.burnin, self.sample, self.thin = burnin, sample, thin
        self.bk_mean_approx = bk_mean_approx
        self.rng, self.pg_factory, self.pg_draw = rng, pg_factory, pg_draw
        self.itype = [np.where(G == k)[0] for k in range(K)] if self.hasSC else None
        # gem.py:208-222
        self.init_alpha = None
        if self.hasSC:
            self.L, self.N = SCexpr.shape
            ok = init_pkappa is not None and len(init_pkappa) == 2
            self.init_pkappa = init_pkappa if ok else np.array([-1., 0.01])
            ok = init_ptau is not None and len(init_ptau) == 2
            self.init_ptau = init_ptau if ok else np.array([self.N, 0.01 * self.N], dtype=float)
            self.pkappa, self.ptau = np.copy(self.init_pkappa), np.copy(self.init_ptau)
        if self.hasBK:
            self.M, self.N = BKexpr.shape
            ok = init_alpha is not None and len(init_alpha) == K
            self.init_alpha = init_alpha if ok else np.ones([K])
            self.alpha = np.copy(self.init_alpha)
            self.init_pkappa = np.array([-1., 0.01])
            self.init_ptau = np.array([self.N, 0.01 * self.N], dtype=float)
        self.init_A = init_A(SCexpr, BKexpr, K, self.itype, min_A, init_A_)
        self.A = np.copy(self.init_A)

    def estep(self):  # gem.py:135-149
        self.suff_stats = {}
        if self.hasSC:
            self.sc.update_parameters(self.A, self.pkappa, self.ptau)
            self.sc.gibbs(self.burnin, self.sample, self.thin)
            self.suff_stats.update(self.sc.suff_stats)
        if self.hasBK:
            self.bk.update_parameters(self.A, self.alpha)
            self.bk.gibbs(self.burnin, self.sample, self.thin, self.bk_mean_approx)
            self.suff_stats.update(self.bk.suff_stats)

    def gem(self):  # gem.py:152-195
        conv, niter, old = self.EM_CONV + 1, 0, -10 ** 6
        path = []
        self.estep_elbo = []
        if self.hasSC:
            self.sc = OracleGibbsSC(self.init_A, self.init_pkappa, self.init_ptau, self.Y,
                                    self.G, self.itype, rng=self.rng,
                                    pg_factory=self.pg_factory, pg_draw=self.pg_draw)
            self.sc.init_gibbs()
        if self.hasBK:
            self.bk = OracleGibbsBK(self.init_A, self.init_alpha, self.X, self.iMarkers,
                                    rng=self.rng)
            self.bk.init_gibbs()
        self.mle = OracleMLE(self.X, self.Y, self.G, self.K, self.itype, self.hasBK,
                             self.hasSC, self.init_A, self.init_alpha, self.init_pkappa,
                             self.init_ptau, self.min_A, 1, self.MLE_CONV, self.MLE_maxiter)
        elbo = None
        while abs(conv) > self.EM_CONV and niter < self.EM_maxiter:
            self.estep()
            self.mle.update_suff_stats(self.suff_stats)
            self.estep_elbo.append(self.mle.compute_elbo())
            if self.hasSC:
                self.mle.opt_kappa_tau()
                self.pkappa, self.ptau = self.mle.pkappa, self.mle.ptau
            if self.hasBK and self.est_alpha:
                self.mle.opt_alpha()
                self.alpha = self.mle.alpha
            self.mle.opt_A_u()
            self.A = self.mle.A
            elbo = self.mle.compute_elbo()
            conv = abs(elbo - old)
            old = elbo
            niter += 1
            path.append(elbo)
        self.path_elbo = np.array(path)
        return niter, elbo, conv, self.path_elbo


# --------------------------------------------------------------------------
# demo/demo_simulate_data.py:11-114  --  synthetic inputs (definition of the
# benchmark's input distribution)
# --------------------------------------------------------------------------


def simulate(N, K, L, M, seed=0, G=None, alpha=None, anchor=2, anti=2,
             kappa=-1.0, kappa_sd=0.5, tau_scale=1.5, tau_sd_frac=0.1,
             sc_depth=2.0, bk_depth=50.0):
    """Synthetic (A, bulk X, W, single-cell Y, S, G) following the demo generator
    (vectorised; same distributions, not the same stream)."""
    rs = np.random.RandomState(seed)
    A = np.exp(rs.normal(size=(N, K)))
    for k in range(K):
        rows = np.arange(anchor * k, anchor * (k + 1))
        for other in range(K):
            if other != k:
                A[rows, other] = 0
    for k in range(K):
        A[np.arange(anchor * K + anti * k, anchor * K + anchor * (k + 1)), k] = 0
    A = A / A.sum(axis=0)
    out = dict(A=A)
    if M > 0:
        alpha = np.full(K, 2.0) if alpha is None else np.asarray(alpha, dtype=float)
        W = rs.dirichlet(alpha, M)
        depths = rs.poisson(bk_depth * N, M)
        mix = W.dot(A.T)
        X = np.empty((M, N))
        for m in range(M):
            X[m] = rs.multinomial(depths[m], mix[m])
        out.update(X=X, W=W)
    if L > 0:
        G = (np.arange(L) % K) if G is None else np.asarray(G)
        tau = tau_scale * N
        kap = rs.normal(kappa, kappa_sd, size=L)
        taus = rs.normal(tau, tau * tau_sd_frac, size=L)
        prob = expit(A[:, G].T * taus[:, None] + kap[:, None])
        S = (rs.random_sample((L, N)) < prob).astype(int)
        depths = rs.poisson(sc_depth * N, L)
        Y = np.empty((L, N))
        for l in range(L):
            pr = A[:, G[l]] * S[l]
            Y[l] = rs.multinomial(max(depths[l], 1), pr / pr.sum())
        out.update(Y=Y, S=S, G=G, kappa=kap, tau=taus)
    return out
