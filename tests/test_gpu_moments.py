"""North-star checks #2 and #3 for the stochastic kernels: per-entry sampler moments
against exact expectations, long chains against long chains of the REFERENCE classes
(tests/golden/stat_chains.npz), and end-to-end estimates against reference reruns
(tests/golden/stat_demo_fits.npz).  Tolerances are Monte-Carlo error bars, written out."""
import numpy as np
import pytest
from scipy.special import psi as digamma

from oracle import ursm_oracle as orc

pytestmark = pytest.mark.gpu


def _itype(G, K):
    return [np.where(G == k)[0] for k in range(K)]


def _zscores(obs, exp, sd, floor=0.0):
    sd = np.maximum(np.asarray(sd, dtype=float), floor)
    ok = sd > 0
    return ((np.asarray(obs) - np.asarray(exp))[ok] / sd[ok]).ravel()


# ---------------------------------------------------------------------------
# E[w | psi] = tanh(psi/2)/(2 psi), Var from the closed form
# ---------------------------------------------------------------------------
def test_pg_draw_moments_on_device(golden):
    import ursm_b200.e_step_gibbs as es
    g = golden("kat_sc")
    Y, G, K = g["Y"], g["G"], 3
    L, N = Y.shape
    A = g["init_A"]
    sc = es.LogitNormalGibbs_SC(A, g["init_pkappa"], g["init_ptau"], Y, G, _itype(G, K), seed=7)
    rs = np.random.RandomState(3)
    S_in = (Y > 0).astype(np.int64)
    kappa, tau = rs.normal(-1, 0.5, L), rs.normal(1.5 * N, 20, L)
    psi = kappa[:, None] + tau[:, None] * A[:, G].T.astype(np.float32)
    T = 300
    acc, acc2 = np.zeros((L, N)), np.zeros((L, N))
    for sweep in range(T):
        w = sc.sweep_debug(S_in, kappa, tau, sweep=sweep, seed=7, em_iter=0)[0].astype(np.float64)
        acc += w
        acc2 += w * w
    m, v = orc.pg1_mean(psi), orc.pg1_var(psi)
    z = _zscores(acc / T, m, np.sqrt(v / T))
    assert np.abs(z).max() < 5.5, np.abs(z).max()      # 5000 entries: P(|z|>5.5) ~ 2e-4
    assert abs(np.mean(z)) < 4.0 / np.sqrt(z.size)       # no systematic bias
    assert abs(np.mean(z * z) - 1.0) < 0.1
    var_hat = acc2 / T - (acc / T) ** 2
    assert abs(np.mean(var_hat / v) - 1.0) < 0.03


# ---------------------------------------------------------------------------
# bulk Gibbs: E[Z | W] = draw_Z_mean (e_step_gibbs.py:149-155) with W frozen,
# E[W | Z], E[log W | Z] of the Dirichlet (:140-146) with Z frozen
# ---------------------------------------------------------------------------
def test_bulk_split_moments_frozen_W(golden):
    import ursm_b200.e_step_gibbs as es
    g = golden("kat_bulk")
    X, A = g["X"], g["init_A"]
    M, N = X.shape
    K = A.shape[1]
    rs = np.random.RandomState(0)
    W = rs.dirichlet(np.ones(K) * 2, M).T  # K x M
    bk = es.LogitNormalGibbs_BK(A, np.ones(K), X, seed=11)
    bk.set_state(W)
    bk.freeze = 1
    S = 400
    bk.gibbs(burnin=0, sample=S, thin=1, mean_approx=False)
    AW = A.dot(W)                                   # N x M
    P = (A[None, :, :] * W.T[:, None, :]) / AW.T[:, :, None]   # M x N x K
    EZ = X[:, :, None] * P
    VZ = X[:, :, None] * P * (1 - P)
    z1 = _zscores(bk.suff_stats["exp_Zik"], EZ.sum(axis=0), np.sqrt(VZ.sum(axis=0) / S))
    z2 = _zscores(bk.suff_stats["exp_Zjk"], EZ.sum(axis=1), np.sqrt(VZ.sum(axis=1) / S))
    for z in (z1, z2):
        assert np.abs(z).max() < 5.0
        # mean of n chi^2_1 variables has sd sqrt(2/n); entries of one sample / gene are
        # negatively correlated through the sum constraint, hence 6 "independent" sd
        assert abs(np.mean(z * z) - 1.0) < 6.0 * np.sqrt(2.0 / z.size)
        assert abs(np.mean(z)) < 5.0 / np.sqrt(z.size)
    # exact invariants of a multinomial split
    np.testing.assert_allclose(bk.suff_stats["exp_Zjk"].sum(axis=1), X.sum(axis=1), rtol=1e-12)
    np.testing.assert_allclose(bk.suff_stats["exp_Zik"].sum(axis=1), X.sum(axis=0), rtol=1e-12)


@pytest.mark.parametrize("split", ["auto", "lockstep"])
def test_bulk_split_moments_large_counts(split, monkeypatch):
    """Deep samples (counts in the thousands): the BTRS branch of the binomial chain, in the
    per-lane kernel (what the library picks for such data) and through the hand-over list of
    the lockstep kernel."""
    import ursm_b200.e_step_gibbs as es
    if split != "auto":
        monkeypatch.setenv("URSM_BK_SPLIT", split)
    rs = np.random.RandomState(4)
    M, N, K = 24, 96, 4
    A = rs.dirichlet(np.ones(N) * 0.7, K).T
    A = np.maximum(A, 1e-6)
    A /= A.sum(axis=0)
    W = rs.dirichlet(np.ones(K) * 1.5, M).T
    X = rs.poisson(3.0e5 * A.dot(W).T).astype(float)      # mean count ~3000
    bk = es.LogitNormalGibbs_BK(A, np.ones(K), X, seed=23)
    bk.set_state(W)
    bk.freeze = 1
    S = 300
    bk.gibbs(burnin=0, sample=S, thin=1, mean_approx=False)
    AW = A.dot(W)
    P = (A[None, :, :] * W.T[:, None, :]) / AW.T[:, :, None]
    EZ, VZ = X[:, :, None] * P, X[:, :, None] * P * (1 - P)
    for got, ax in ((bk.suff_stats["exp_Zik"], 0), (bk.suff_stats["exp_Zjk"], 1)):
        z = _zscores(got, EZ.sum(axis=ax), np.sqrt(VZ.sum(axis=ax) / S))
        assert np.abs(z).max() < 5.0
        assert abs(np.mean(z * z) - 1.0) < 6.0 * np.sqrt(2.0 / z.size)
        assert abs(np.mean(z)) < 5.0 / np.sqrt(z.size)
    np.testing.assert_allclose(bk.suff_stats["exp_Zjk"].sum(axis=1), X.sum(axis=1), rtol=1e-12)


def test_bulk_dirichlet_moments_frozen_Z():
    import ursm_b200.e_step_gibbs as es
    rs = np.random.RandomState(1)
    M, N, K = 64, 40, 4
    X = rs.poisson(30, (M, N)).astype(float)
    A = rs.dirichlet(np.ones(N), K).T
    alpha = np.array([0.3, 1.0, 2.5, 7.0])
    n = rs.poisson(40, (M, K)).astype(float)
    n[:5] = 0  # pure prior draws, including shape < 1
    bk = es.LogitNormalGibbs_BK(A, alpha, X, seed=5)
    bk.set_state(np.full((K, M), 1.0 / K), n)
    bk.freeze = 2
    S = 2000
    bk.gibbs(burnin=0, sample=S, thin=1, mean_approx=False)
    a = alpha[None, :] + n
    a0 = a.sum(axis=1, keepdims=True)
    EW = a / a0
    VW = a * (a0 - a) / (a0 ** 2 * (a0 + 1))
    z = _zscores(bk.suff_stats["exp_W"].T, EW, np.sqrt(VW / S))
    assert np.abs(z).max() < 5.0 and abs(np.mean(z * z) - 1) < 6.0 * np.sqrt(2.0 / z.size)
    from scipy.special import polygamma
    ElogW = digamma(a) - digamma(a0)
    VlogW = polygamma(1, a) - polygamma(1, a0)
    z = _zscores(bk.suff_stats["exp_logW"].T, ElogW, np.sqrt(VlogW / S))
    assert np.abs(z).max() < 5.0 and abs(np.mean(z * z) - 1) < 6.0 * np.sqrt(2.0 / z.size)
    np.testing.assert_allclose(bk.suff_stats["exp_Zjk"], n, rtol=1e-12)


# ---------------------------------------------------------------------------
# long chains vs the reference's own long chains
# ---------------------------------------------------------------------------
def _chain_check(got, g, key, nsd=6.0, floor_rel=2e-3, poisson=False):
    mean, sd = g["mean_" + key], g["sd_" + key]
    reps = int(g["reps"])
    if poisson:  # rare counts: 4 chains cannot estimate their spread; Poisson floor
        sd = np.maximum(sd, np.sqrt((np.abs(mean) + 1.0 / int(g["sample"])) / int(g["sample"])))
    # difference of one chain and the mean of `reps` chains; sd itself is estimated from
    # `reps` chains (noisy), hence the generous multiplier and a small relative floor
    tol = nsd * np.sqrt(1 + 1.0 / reps) * sd + floor_rel * np.maximum(np.abs(mean), 1e-3)
    diff = np.abs(np.asarray(got, dtype=float) - mean)
    assert np.all(diff <= tol), (key, float((diff / tol).max()))
    return diff / np.maximum(sd, 1e-300)


def test_single_cell_chain_matches_reference_chain(golden):
    import ursm_b200.e_step_gibbs as es
    g = golden("stat_chains")
    Y, G = g["Y"], g["G"]
    sc = es.LogitNormalGibbs_SC(g["A"], g["pkappa"], g["ptau"], Y, G, _itype(G, 3), seed=99)
    sc.gibbs(burnin=int(g["burnin"]), sample=int(g["sample"]), thin=1)
    st = sc.suff_stats
    eS = np.asarray(st["exp_S"])
    assert np.all(eS[Y > 0] == 1)
    # E[S] entries: binomial-type noise; floor for the entries every chain agrees on
    mean, sd = g["mean_exp_S"], np.maximum(g["sd_exp_S"], 0.004)
    z = (eS - mean) / (sd * np.sqrt(1.25))
    assert np.abs(z).max() < 7.0 and np.mean(z * z) < 2.5
    for key in ("exp_kappa", "exp_tau", "exp_kappasq", "exp_tausq", "coeffA", "coeffAsq",
                "exp_elbo_const"):
        _chain_check(st[key], g, key)


def test_bulk_chain_matches_reference_chain(golden):
    import ursm_b200.e_step_gibbs as es
    g = golden("stat_chains")
    bk = es.LogitNormalGibbs_BK(g["A"], g["alpha"], g["X"], seed=3)
    bk.init_gibbs()
    bk.gibbs(burnin=int(g["burnin"]), sample=int(g["sample"]), thin=1, mean_approx=False)
    for key in ("exp_Zik", "exp_Zjk", "exp_logW", "exp_W"):
        _chain_check(bk.suff_stats[key], g, key, poisson=key.startswith("exp_Z"))


# ---------------------------------------------------------------------------
# end-to-end: the reference's `make run` configuration on the demo data
# ---------------------------------------------------------------------------
def test_demo_fit_matches_reference_reruns(golden):
    import ursm_b200.gem as gm
    d, f = golden("demo_data"), golden("stat_demo_fits")
    gem = gm.LogitNormalGEM(BKexpr=d["X"], SCexpr=d["Y"], K=3, G=d["G"], burnin=10, sample=20,
                            thin=1, bk_mean_approx=True, MLE_CONV=1e-6, MLE_maxiter=500,
                            EM_CONV=1e-6, EM_maxiter=10, seed=2)
    niter, elbo, conv, path = gem.gem()
    assert niter == 10
    # ELBO path: run-to-run sd of the reference is 0.2 .. 2.3 (of ~ -12300)
    assert np.all(np.abs(path - f["mean_path"]) < 6 * f["sd_path"] + 1.0), path - f["mean_path"]
    # estimates vs reference reruns (SURVEY 4 tolerance bands)
    A, Aref = gem.A, f["mean_A"]
    assert np.linalg.norm(A - Aref) / np.linalg.norm(Aref) < 0.02
    assert np.corrcoef(A.ravel(), Aref.ravel())[0, 1] > 0.9995
    np.testing.assert_allclose(gem.alpha, f["mean_alpha"], rtol=0.02)
    assert abs(gem.pkappa[0] - f["mean_pkappa"][0]) < 0.25
    np.testing.assert_allclose(gem.ptau[0], f["mean_ptau"][0], rtol=0.01)
    assert np.abs(gem.suff_stats["exp_W"] - f["mean_exp_W"]).max() < 0.01
    eS = np.asarray(gem.suff_stats["exp_S"])
    assert np.corrcoef(eS.ravel(), f["mean_exp_S"].ravel())[0, 1] > 0.95
    assert np.abs(eS - f["mean_exp_S"]).mean() < 0.08   # sample=20: MC noise ~ 0.11 per entry
    # and vs ground truth (what demo_plots.py eyeballs): shipped run has corr 0.983 / 0.992
    assert np.corrcoef(A.ravel(), d["A_true"].ravel())[0, 1] > 0.97
    assert np.corrcoef(gem.suff_stats["exp_W"].T.ravel(), d["W_true"].ravel())[0, 1] > 0.98


def test_sharding_invariance_single_process(golden):
    """Philox is keyed on GLOBAL cell / sample ids: two shards (row_offset) reproduce the
    unsharded chain exactly per row; summed statistics agree to rounding."""
    import ursm_b200.e_step_gibbs as es
    g = golden("stat_chains")
    Y, G, X = g["Y"], g["G"], g["X"]
    L, cut = Y.shape[0], 5
    kw = dict(seed=17)
    full = es.LogitNormalGibbs_SC(g["A"], g["pkappa"], g["ptau"], Y, G, _itype(G, 3), **kw)
    full.gibbs(burnin=20, sample=30, thin=1)
    parts = []
    for lo, hi in ((0, cut), (cut, L)):
        p = es.LogitNormalGibbs_SC(g["A"], g["pkappa"], g["ptau"], Y[lo:hi], G[lo:hi],
                                   _itype(G[lo:hi], 3), row_offset=lo, total_rows=L, **kw)
        p.gibbs(burnin=20, sample=30, thin=1)
        parts.append(p.suff_stats)
    eS = np.vstack([np.asarray(p["exp_S"]) for p in parts])
    np.testing.assert_array_equal(eS, np.asarray(full.suff_stats["exp_S"]))
    for key in ("exp_kappa", "exp_tau"):
        np.testing.assert_allclose(np.hstack([p[key] for p in parts]), full.suff_stats[key],
                                   rtol=1e-12)
    for key in ("coeffA", "coeffAsq", "exp_elbo_const"):
        tot = sum(np.asarray(p[key]) for p in parts)
        np.testing.assert_allclose(tot, full.suff_stats[key], rtol=2e-5, atol=1e-5)
    # bulk
    M, cutm = X.shape[0], 2
    fb = es.LogitNormalGibbs_BK(g["A"], g["alpha"], X, **kw)
    fb.init_gibbs()
    fb.gibbs(burnin=5, sample=10, thin=1, mean_approx=False)
    zik, W = 0, []
    for lo, hi in ((0, cutm), (cutm, M)):
        p = es.LogitNormalGibbs_BK(g["A"], g["alpha"], X[lo:hi], row_offset=lo, total_rows=M, **kw)
        p.init_gibbs()
        p.gibbs(burnin=5, sample=10, thin=1, mean_approx=False)
        zik = zik + p.suff_stats["exp_Zik"]
        W.append(p.suff_stats["exp_W"])
    np.testing.assert_allclose(zik, fb.suff_stats["exp_Zik"], rtol=1e-12)
    np.testing.assert_allclose(np.hstack(W), fb.suff_stats["exp_W"], rtol=1e-12)
