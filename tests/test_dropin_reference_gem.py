"""The drop-in claim of INTEGRATION.md sec. 2 (missing #4).

CPU side (only where /root/reference exists, i.e. in the build container): the reference's OWN
`gem.py`, with its three star imports (gem.py:60-62) pointed at `ursm_b200` exactly as
INTEGRATION.md shows, is exec'd; every constructor / method call it makes on the hot-path
classes must bind against this package's signatures, and building the samplers without a GPU
must fail loudly (no CPU fallback).

GPU side: `tests/ref_style_driver.py` drives the classes through the reference's API surface
only (the GPU box has no reference tree): the deterministic bulk-only known answer KAT-B and the
joint demo fit, compared with the fixtures generated from the reference."""
import ast
import inspect
import os
import re

import numpy as np
import pytest

REF = "/root/reference"
HOT = ("LogitNormalGibbs_SC", "LogitNormalGibbs_BK", "LogitNormalMLE")


def _reference_gem_source():
    src = open(os.path.join(REF, "gem.py")).read()
    # the one edit inside the reference tree
    src, n = re.subn(r"from utils import \*.*\nfrom e_step_gibbs import \*.*\nfrom m_step import \*.*\n",
                     "from ursm_b200.utils import simplex_proj, std_row\n"
                     "from ursm_b200.e_step_gibbs import LogitNormalGibbs_SC, LogitNormalGibbs_BK\n"
                     "from ursm_b200.m_step import LogitNormalMLE\n", src)
    assert n == 1, "gem.py:60-62 no longer looks like the star imports INTEGRATION.md describes"
    return src.replace("xrange", "range")  # the reference is Python 2


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "gem.py")),
                    reason="reference tree not present (GPU box)")
def test_reference_gem_binds_to_this_package():
    import ursm_b200.e_step_gibbs as es
    import ursm_b200.m_step as ms
    from ursm_b200 import _lib
    src = _reference_gem_source()
    ns = {"__name__": "reference_gem"}
    exec(compile(src, "reference/gem.py", "exec"), ns)
    assert ns["LogitNormalGibbs_SC"] is es.LogitNormalGibbs_SC
    assert ns["LogitNormalGibbs_BK"] is es.LogitNormalGibbs_BK
    assert ns["LogitNormalMLE"] is ms.LogitNormalMLE
    ours = {"LogitNormalGibbs_SC": es.LogitNormalGibbs_SC, "LogitNormalGibbs_BK": es.LogitNormalGibbs_BK,
            "LogitNormalMLE": ms.LogitNormalMLE}
    owner = {"Gibbs_SC": es.LogitNormalGibbs_SC, "Gibbs_BK": es.LogitNormalGibbs_BK,
             "mle": ms.LogitNormalMLE}
    seen = set()
    for node in ast.walk(ast.parse(src)):
        if not isinstance(node, ast.Call):
            continue
        f = node.func
        args = [object()] * len(node.args)
        kwargs = {k.arg: object() for k in node.keywords}
        if isinstance(f, ast.Name) and f.id in ours:           # constructor calls
            inspect.signature(ours[f.id].__init__).bind(object(), *args, **kwargs)
            seen.add(f.id)
        elif (isinstance(f, ast.Attribute) and isinstance(f.value, ast.Attribute)
              and isinstance(f.value.value, ast.Name) and f.value.value.id == "self"
              and f.value.attr in owner):                      # self.Gibbs_SC.gibbs(...), self.mle.x()
            meth = getattr(owner[f.value.attr], f.attr)        # AttributeError = missing method
            inspect.signature(meth).bind(object(), *args, **kwargs)
            seen.add(f.value.attr + "." + f.attr)
    assert set(HOT) <= seen
    for need in ("Gibbs_SC.init_gibbs", "Gibbs_SC.update_parameters", "Gibbs_SC.gibbs",
                 "Gibbs_BK.init_gibbs", "Gibbs_BK.update_parameters", "Gibbs_BK.gibbs",
                 "mle.update_suff_stats", "mle.compute_elbo", "mle.opt_kappa_tau", "mle.opt_alpha",
                 "mle.opt_A_u"):
        assert need in seen, need
    # attributes gem.py reads back from the M-step object (gem.py:173-183)
    for attr in re.findall(r"self\.mle\.([A-Za-z_]+)\b(?!\()", src):
        assert attr in ("pkappa", "ptau", "alpha", "A"), attr
    # and the reference's driver class, built on top of this package, refuses to run without
    # a GPU instead of falling back to anything
    import torch
    if not torch.cuda.is_available():
        d = np.load(os.path.join(os.path.dirname(__file__), "golden", "demo_data.npz"))
        with pytest.raises(_lib.UrsmError, match="CUDA"):  # (its init_para already projects A)
            gem = ns["LogitNormalGEM"](BKexpr=d["X"], SCexpr=d["Y"], K=3, G=d["G"], burnin=2,
                                       sample=2, EM_maxiter=1)
            gem.gem()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "gem.py")),
                    reason="reference tree not present (GPU box)")
def test_reference_gem_runs_on_this_package():  # pragma: no cover  (needs reference AND a GPU)
    src = _reference_gem_source()
    ns = {"__name__": "reference_gem"}
    exec(compile(src, "reference/gem.py", "exec"), ns)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_bulk.npz"))
    gem = ns["LogitNormalGEM"](BKexpr=g["X"], K=3, init_A=np.copy(g["init_A"]), MLE_CONV=1e-6,
                               MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=5)
    np.testing.assert_allclose(gem.gem()[3], g["path_elbo"], rtol=1e-8)


def _classes():
    import ursm_b200.e_step_gibbs as es
    import ursm_b200.m_step as ms
    return es.LogitNormalGibbs_SC, es.LogitNormalGibbs_BK, ms.LogitNormalMLE


@pytest.mark.gpu
def test_reference_style_driver_kat_bulk(golden):
    """KAT-B through nothing but the reference's API surface."""
    from ref_style_driver import fit
    from ursm_b200.utils import simplex_proj
    g = golden("kat_bulk")
    init_A = np.copy(g["init_A"])
    for k in range(3):  # gem.py:253-254: a user-supplied init_A is projected column by column
        init_A[:, k] = simplex_proj(init_A[:, k], 1e-6)
    out = fit(_classes(), BKexpr=g["X"], SCexpr=None, G=None, K=3, init_A=init_A,
              init_alpha=np.ones(3), init_pkappa=None, init_ptau=None, MLE_CONV=1e-6,
              MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=5)
    np.testing.assert_allclose(out["path_elbo"], g["path_elbo"], rtol=1e-8)
    np.testing.assert_allclose(out["A"], g["A"], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(out["alpha"], g["alpha"], rtol=1e-7)
    np.testing.assert_allclose(out["suff_stats"]["exp_W"], g["exp_W"], rtol=1e-6, atol=1e-8)


@pytest.mark.gpu
def test_reference_style_driver_joint_demo(golden):
    """The `make run` configuration of the reference (makefile:22-31) through the reference's
    API surface: estimates within the bands of the reference's own reruns."""
    from oracle import ursm_oracle as orc
    from ref_style_driver import fit
    d, f = golden("demo_data"), golden("stat_demo_fits")
    Y, X, G = d["Y"], d["X"], d["G"]
    N = Y.shape[1]
    itype = [np.where(G == k)[0] for k in range(3)]
    init_A = orc.init_A(Y, X, 3, itype, 1e-6)
    out = fit(_classes(), BKexpr=X, SCexpr=Y, G=G, K=3, init_A=init_A, init_alpha=np.ones(3),
              init_pkappa=np.array([-1., 0.01]), init_ptau=np.array([N, 0.01 * N], dtype=float),
              burnin=10, sample=20, MLE_CONV=1e-6, MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=10)
    assert out["niter"] == 10
    assert np.all(np.abs(out["path_elbo"] - f["mean_path"]) < 6 * f["sd_path"] + 1.0)
    A, Aref = out["A"], f["mean_A"]
    assert np.linalg.norm(A - Aref) / np.linalg.norm(Aref) < 0.02
    np.testing.assert_allclose(out["alpha"], f["mean_alpha"], rtol=0.02)
    assert np.abs(out["suff_stats"]["exp_W"] - f["mean_exp_W"]).max() < 0.01
    eS = np.asarray(out["suff_stats"]["exp_S"])
    assert eS.shape == Y.shape and np.all(eS[Y > 0] == 1)


@pytest.mark.gpu
def test_mstep_accepts_statistics_of_another_sampler():
    """m_step.py:50-89 takes ANY dict with the reference's keys: here the numpy oracle's Gibbs
    samplers (host arrays, no device handles) feed this package's M-step, which must then agree
    with the oracle's M-step on the same dict (missing #4)."""
    from oracle import ursm_oracle as orc
    import ursm_b200.m_step as ms
    N, K, L, M = 120, 3, 30, 8
    sim = orc.simulate(N=N, K=K, L=L, M=M, seed=4)
    Y, X, G = sim["Y"], sim["X"], sim["G"]
    itype = [np.where(G == k)[0] for k in range(K)]
    A = orc.init_A(Y, X, K, itype, 1e-6)
    pk, pt, al = np.array([-1., 0.01]), np.array([N, 0.01 * N], dtype=float), np.ones(K)
    rng = np.random.RandomState(3)
    osc = orc.OracleGibbsSC(A, pk, pt, Y, G, itype, rng=rng)
    osc.init_gibbs()
    osc.gibbs(burnin=2, sample=7, thin=1)
    obk = orc.OracleGibbsBK(A, al, X, None, rng=rng)
    obk.init_gibbs()
    obk.gibbs(burnin=2, sample=5, thin=1, mean_approx=False)
    st = dict(osc.suff_stats)
    st.update(obk.suff_stats)
    assert not any(k.startswith("_ursm") for k in st)
    mle = ms.LogitNormalMLE(X, Y, G, K, itype, True, True, A, al, pk, pt, 1e-6, 1, 1e-8, 6)
    omle = orc.OracleMLE(X, Y, G, K, itype, True, True, A, al, pk, pt, 1e-6, 1, 1e-8, 6)
    mle.update_suff_stats(dict(st))
    omle.update_suff_stats(dict(st))
    np.testing.assert_allclose(mle.u, omle.u, rtol=1e-10)
    np.testing.assert_allclose(mle.gd_coeffAinv, omle.gd_coeffAinv, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(mle.gd_coeffConst, omle.gd_coeffConst, rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(mle.compute_elbo(), omle.compute_elbo(), rtol=1e-9)
    mle.opt_kappa_tau()
    omle.opt_kappa_tau()
    np.testing.assert_allclose(mle.pkappa, omle.pkappa, rtol=1e-9)
    assert mle.opt_alpha() == omle.opt_alpha()
    mle.opt_A_u()
    omle.opt_A_u()
    np.testing.assert_allclose(mle.A, omle.A, rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(mle.compute_elbo(), omle.compute_elbo(), rtol=1e-8)
    # a dict that is not the statistic of a Gibbs run is refused, not quantised
    bad = dict(st)
    bad["exp_S"] = np.asarray(st["exp_S"]) * 0.999 + 1e-4 * rng.rand(L, N)
    from ursm_b200 import _lib
    with pytest.raises(_lib.UrsmError, match="multiple of 1/sample"):
        mle.update_suff_stats(bad)
