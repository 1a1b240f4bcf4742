"""The oracle restatement vs. outputs of the reference itself (tests/golden/*.npz,
produced by tests/golden/make_golden.py running the real reference modules)."""
import numpy as np

from oracle import ursm_oracle as orc

RT = dict(rtol=1e-11, atol=1e-12)


def test_simplex_proj_matches_reference(golden):
    g = golden("kat_simplex")
    for n in range(int(g["ncases"])):
        got = orc.simplex_proj(g["y%d" % n], float(g["min%d" % n]))
        np.testing.assert_array_equal(got, g["proj%d" % n])


def test_simplex_proj_properties():
    rs = np.random.RandomState(0)
    for n in (3, 50, 2000):
        y = rs.normal(0, 0.1, n)
        p = orc.simplex_proj(y, 1e-6)
        assert abs(p.sum() - 1) < 1e-12 and p.min() >= 1e-6
        np.testing.assert_allclose(orc.simplex_proj(p, 1e-6), p, atol=1e-15)  # idempotent


def test_kat_bulk_mean_approx_path(golden):
    g = golden("kat_bulk")
    gem = orc.OracleGEM(BKexpr=g["X"], K=3, init_A_=np.copy(g["init_A"]), MLE_CONV=1e-6,
                        MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=5)
    path = gem.gem()[3]
    np.testing.assert_allclose(path, g["path_elbo"], **RT)
    # the numbers SURVEY App. C quotes
    np.testing.assert_allclose(path[0], -12427.9869411170, rtol=1e-12)
    np.testing.assert_allclose(path[-1], -12173.5932379347, rtol=1e-12)
    np.testing.assert_allclose(gem.A, g["A"], **RT)
    np.testing.assert_allclose(gem.alpha, g["alpha"], **RT)
    for k in ("exp_W", "exp_logW", "exp_Zik", "exp_Zjk"):
        np.testing.assert_allclose(gem.suff_stats[k], g[k], **RT)
    gem2 = orc.OracleGEM(BKexpr=g["X"], K=3, MLE_CONV=1e-6, MLE_maxiter=500, EM_CONV=1e-6,
                         EM_maxiter=3)
    np.testing.assert_allclose(gem2.gem()[3], g["path_elbo_sym"], **RT)


def test_bulk_markers_mean_step(golden):
    g = golden("kat_bulk")
    bk = orc.OracleGibbsBK(g["init_A"], g["mk_alpha"], g["X"], g["mk_iMarkers"])
    bk.init_gibbs()
    np.testing.assert_allclose(bk.W, g["mk_W0"], **RT)
    bk.gibbs(mean_approx=True)
    for k in ("exp_W", "exp_logW", "exp_Zik", "exp_Zjk"):
        np.testing.assert_allclose(bk.suff_stats[k], g["mk_" + k], **RT)


def _inject(sc, g, T):
    sc._zero_stats()
    for t in range(T):
        sc.S, sc.w = g["inj_S"][t], g["inj_w"][t]
        sc.kappa, sc.tau = g["inj_kappa"][t], g["inj_tau"][t]
        sc.accumulate(T)


def test_kat_sc_injected_stats_and_mstep(golden):
    g = golden("kat_sc")
    Y, X, G, K = g["Y"], g["X"], g["G"], 3
    itype = [np.where(G == k)[0] for k in range(K)]
    sc = orc.OracleGibbsSC(g["init_A"], g["init_pkappa"], g["init_ptau"], Y, G, itype,
                           rng=np.random.RandomState(0))
    _inject(sc, g, 4)
    for k in ("exp_S", "exp_kappa", "exp_tau", "exp_kappasq", "exp_tausq", "coeffA",
              "coeffAsq", "exp_elbo_const"):
        np.testing.assert_allclose(sc.suff_stats[k], g["st_" + k], **RT)
    np.testing.assert_allclose(sc.suff_stats["exp_elbo_const"], -1299.9097533145, rtol=1e-12)
    # posterior of (kappa, tau) given injected (w, S)
    for t in range(4):
        s_as = (sc.A[:, G].T * g["inj_S"][t]).sum(axis=1)
        for l in (0, 7, 49):
            m, s = orc.sc_kappa_tau_posterior(g["inj_w"][t][l], g["inj_S"][t][l], sc.A[:, G[l]],
                                              s_as[l], sc.N, sc.pkappa, sc.ptau)
            np.testing.assert_allclose(m, g["post_mean"][t, l], **RT)
            np.testing.assert_allclose(s, g["post_cov"][t, l], **RT)
    bk = orc.OracleGibbsBK(g["init_A"], g["init_alpha"], X)
    bk.init_gibbs()
    bk.gibbs(mean_approx=True)
    st = dict(sc.suff_stats)
    st.update(bk.suff_stats)
    mle = orc.OracleMLE(X, Y, G, K, itype, True, True, g["init_A"], g["init_alpha"],
                        g["init_pkappa"], g["init_ptau"], 1e-6, 1, 1e-6, 500)
    mle.update_suff_stats(st)
    np.testing.assert_allclose(mle.u, g["u"], **RT)
    np.testing.assert_allclose(mle.gd_coeffConst, g["gd_coeffConst"], **RT)
    np.testing.assert_allclose(mle.gd_coeffAinv, g["gd_coeffAinv"], **RT)
    np.testing.assert_allclose(mle.compute_elbo(), g["elbo_E"], **RT)
    np.testing.assert_allclose(mle.compute_elbo(), -12519.5704402744, rtol=1e-12)
    mle.opt_kappa_tau()
    np.testing.assert_allclose(mle.pkappa, g["pkappa"], **RT)
    np.testing.assert_allclose(mle.ptau, g["ptau"], **RT)
    assert mle.opt_alpha() == int(g["niter_alpha"])
    np.testing.assert_allclose(mle.alpha, g["alpha"], **RT)
    mle.opt_A_u()
    np.testing.assert_allclose(mle.A, g["A"], **RT)
    np.testing.assert_allclose(mle.compute_elbo(), g["elbo_M"], **RT)
    np.testing.assert_allclose(mle.compute_elbo(), -12395.0667662578, rtol=1e-12)


def test_kat_sconly(golden):
    g = golden("kat_sconly")
    Y, G, K = g["Y"], g["G"], 3
    itype = [np.where(G == k)[0] for k in range(K)]
    N = Y.shape[1]
    pk, pt = np.array([-1., 0.01]), np.array([N, 0.01 * N], dtype=float)
    sc = orc.OracleGibbsSC(g["init_A"], pk, pt, Y, G, itype, rng=np.random.RandomState(0))
    _inject(sc, g, 3)
    mle = orc.OracleMLE(None, Y, G, K, itype, False, True, g["init_A"], None, pk, pt, 1e-6, 1,
                        1e-6, 60)
    mle.update_suff_stats(dict(sc.suff_stats))
    np.testing.assert_allclose(mle.compute_elbo(), g["elbo_E"], **RT)
    mle.opt_kappa_tau()
    mle.opt_A_u()
    np.testing.assert_allclose(mle.A, g["A"], **RT)
    np.testing.assert_allclose(mle.compute_elbo(), g["elbo_M"], **RT)


def test_seeded_gibbs_em_bit_parity(golden):
    """Same numpy.random call order as the reference => identical seeded run."""
    g = golden("kat_gibbs_seeded")
    np.random.seed(int(g["seed"]))
    gem = orc.OracleGEM(BKexpr=g["X"], SCexpr=g["Y"], K=3, G=g["G"], burnin=2, sample=3, thin=1,
                        bk_mean_approx=False, MLE_CONV=1e-6, MLE_maxiter=50, EM_CONV=1e-6,
                        EM_maxiter=2, rng=np.random)
    path = gem.gem()[3]
    np.testing.assert_allclose(path, g["path_elbo"], **RT)
    np.testing.assert_allclose(gem.A, g["A"], **RT)
    np.testing.assert_allclose(gem.alpha, g["alpha"], **RT)
    np.testing.assert_allclose(gem.pkappa, g["pkappa"], **RT)
    for k in ("exp_S", "exp_kappa", "exp_tau", "coeffA", "coeffAsq", "exp_W", "exp_Zik"):
        np.testing.assert_allclose(gem.suff_stats[k], g[k], **RT)


def test_pg1_devroye_moments():
    """The PG(1,z) restatement vs. closed-form mean/variance/Laplace transform."""
    rs = np.random.RandomState(11)
    n = 6000
    for psi in (0.0, 0.7, 2.5, 6.0, 25.0):
        x = np.array([orc.pg1_devroye(psi, rs) for _ in range(n)])
        m, v = float(orc.pg1_mean(psi)), float(orc.pg1_var(psi))
        assert abs(x.mean() - m) < 5 * np.sqrt(v / n), (psi, x.mean(), m)
        assert abs(x.var() - v) < 0.15 * v, (psi, x.var(), v)
        t = 1.3
        lap = np.cosh(psi / 2) / np.cosh(np.sqrt((psi ** 2 / 2 + t) / 2))
        est = np.exp(-t * x)
        assert abs(est.mean() - lap) < 5 * est.std() / np.sqrt(n)
