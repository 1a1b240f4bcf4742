"""Generate the golden vectors in this directory by RUNNING THE REFERENCE ITSELF.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

Nothing from the reference is copied into the repo: its five modules are read
from `/root/reference`, the three mechanical Python-3 edits SURVEY App. C lists
(`xrange`->`range`, `.iteritems()`->`.items()`, `np.float`->`float`) are applied
to the source text in memory, and the result is exec'd into throw-away module
objects.  `pypolyagamma` (absent, un-vendored C++ wheel) is replaced by a shim
whose `pgdrawvpar` is the oracle's Devroye sampler, so seeded stochastic runs of
reference and oracle consume identical random streams.

Outputs (committed): `kat_*.npz` -- inputs are either the demo CSVs, copied into
the npz as arrays so the tests do not need `/root/reference`, or seeded synthetic
data; outputs are what the reference code computed.
"""
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ursm_oracle as orc  # noqa: E402


def load_reference(ref_dir):
    """exec the reference modules (py3-fixed in memory) and return them."""
    shim = types.ModuleType("pypolyagamma")
    shim.get_omp_num_threads = lambda: 1
    shim.PyPolyaGamma = orc.PG1Sampler
    shim.pgdrawvpar = orc.pgdrawv
    sys.modules["pypolyagamma"] = shim
    mods = {}
    for name in ("utils", "e_step_gibbs", "m_step", "gem"):
        with open(os.path.join(ref_dir, name + ".py")) as fh:
            src = fh.read()
        src = src.replace("xrange", "range").replace(".iteritems()", ".items()")
        src = re.sub(r"np\.float\)", "float)", src)
        mod = types.ModuleType(name)
        mod.__file__ = os.path.join(ref_dir, name + ".py")
        sys.modules[name] = mod
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
        mods[name] = mod
    return mods


def demo_data(ref_dir):
    d = os.path.join(ref_dir, "demo", "demo_data")
    ld = lambda f, **kw: np.loadtxt(os.path.join(d, f), delimiter=",", **kw)
    return dict(Y=ld("demo_single_cell_rnaseq_counts.csv"),
                X=ld("demo_bulk_rnaseq_counts.csv"),
                G=ld("demo_single_cell_types.csv", dtype=int),
                A_true=ld("demo_profile_matrix.csv"),
                W_true=ld("demo_bulk_mix_proportions.csv"),
                S_true=ld("demo_single_cell_dropout_status.csv"))


def kat_simplex(ref):
    rs = np.random.RandomState(7)
    cases = []
    for n, min_y, scale in [(5, 1e-6, 1.0), (100, 1e-6, 0.05), (257, 1e-3, 0.01),
                            (1000, 1e-6, 1e-3), (64, 1e-2, 5.0), (33, 1e-6, 1e-9)]:
        y = rs.normal(1.0 / n, scale, size=n)
        cases.append((y, min_y, ref["utils"].simplex_proj(y, min_y)))
    # degenerate: nothing above the floor (iBest stays -1 in the reference loop)
    y = np.full(4, -3.0)
    cases.append((y, 0.3, ref["utils"].simplex_proj(y, 0.3)))
    out = {}
    for n, (y, m, p) in enumerate(cases):
        out["y%d" % n], out["min%d" % n], out["proj%d" % n] = y, m, p
    out["ncases"] = len(cases)
    return out


def kat_bulk(ref, demo):
    """KAT-B: deterministic bulk-only mean-approx fit."""
    gem = ref["gem"].LogitNormalGEM(BKexpr=demo["X"], K=3, init_A=np.copy(demo["A_true"]),
                                    MLE_CONV=1e-6, MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=5)
    _, _, _, path = gem.gem()
    out = dict(X=demo["X"], init_A=demo["A_true"], path_elbo=path, A=gem.A, alpha=gem.alpha,
               exp_W=gem.suff_stats["exp_W"], exp_logW=gem.suff_stats["exp_logW"],
               exp_Zik=gem.suff_stats["exp_Zik"], exp_Zjk=gem.suff_stats["exp_Zjk"])
    # symmetric start (no init_A)
    gem2 = ref["gem"].LogitNormalGEM(BKexpr=demo["X"], K=3, MLE_CONV=1e-6, MLE_maxiter=500,
                                     EM_CONV=1e-6, EM_maxiter=3)
    out["path_elbo_sym"] = gem2.gem()[3]
    # one mean-approx E-step from a non-uniform W, with markers
    iM = np.array([[0, 0], [3, 1], [6, 2]])
    bk = ref["e_step_gibbs"].LogitNormalGibbs_BK(demo["A_true"], np.array([1.5, 2.0, 0.7]),
                                                 demo["X"], iM)
    bk.init_gibbs()
    out["mk_iMarkers"], out["mk_alpha"], out["mk_W0"] = iM, bk.alpha, np.copy(bk.W)
    bk.gibbs(mean_approx=True)
    for k, v in bk.suff_stats.items():
        out["mk_" + k] = v
    return out


def kat_sc(ref, demo):
    """KAT-SC: injected (w, S, kappa, tau) -> stats -> M-step."""
    Y, X, G, K = demo["Y"], demo["X"], demo["G"], 3
    gem = ref["gem"].LogitNormalGEM(BKexpr=X, SCexpr=Y, K=K, G=G, MLE_CONV=1e-6,
                                    MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=1,
                                    burnin=1, sample=1)
    np.random.seed(99)
    gem.init_gibbs()
    gem.init_mle()
    sc, bk, mle = gem.Gibbs_SC, gem.Gibbs_BK, gem.mle
    L, N = Y.shape
    rng = np.random.RandomState(12345)
    T = 4
    inj = dict(S=[], w=[], kappa=[], tau=[])
    post_m, post_s = [], []
    sc.init_suffStats()
    for t in range(T):
        S = ((Y > 0) | (rng.rand(L, N) < 0.4)).astype(int)
        w = rng.gamma(2.0, 0.1, (L, N))
        kappa = rng.normal(-1, 0.3, (1, L))
        tau = rng.normal(100, 1, (1, L))
        for k_, v_ in (("S", S), ("w", w), ("kappa", kappa), ("tau", tau)):
            inj[k_].append(np.copy(v_))
        sc.S, sc.w, sc.kappa, sc.tau = S, w, kappa, tau
        sc.sum_AS = (np.transpose(sc.A[:, sc.G]) * sc.S).sum(axis=1)
        # posterior mean / covariance of (kappa, tau) given this (w, S): capture
        # the arguments the reference hands to multivariate_normal (:376)
        cap = []
        orig = np.random.multivariate_normal
        np.random.multivariate_normal = lambda m, s: (cap.append((m, s)), (0.0, 0.0))[1]
        keep = (np.copy(sc.kappa), np.copy(sc.tau))
        sc.draw_kappa_tau()
        np.random.multivariate_normal = orig
        sc.kappa, sc.tau = keep
        post_m.append(np.array([c[0] for c in cap]))
        post_s.append(np.array([c[1] for c in cap]))
        sc.update_suffStats(T)
    bk.update_parameters(gem.A, gem.alpha)
    bk.gibbs(mean_approx=True)
    st = {}
    st.update(sc.suff_stats)
    st.update(bk.suff_stats)
    out = dict(Y=Y, X=X, G=G, init_A=np.copy(gem.init_A), init_pkappa=gem.init_pkappa,
               init_ptau=gem.init_ptau, init_alpha=gem.init_alpha,
               inj_S=np.array(inj["S"]), inj_w=np.array(inj["w"]),
               inj_kappa=np.array(inj["kappa"]), inj_tau=np.array(inj["tau"]),
               post_mean=np.array(post_m), post_cov=np.array(post_s))
    for k_, v_ in st.items():
        out["st_" + k_] = np.array(v_)
    mle.update_suff_stats(st)
    out.update(u=np.copy(mle.u), gd_coeffConst=np.copy(mle.gd_coeffConst),
               gd_coeffAinv=np.copy(mle.gd_coeffAinv), gd_coeffA=np.copy(mle.gd_coeffA),
               elbo_const=mle.elbo_const, elbo_E=mle.compute_elbo())
    mle.opt_kappa_tau()
    out.update(pkappa=np.copy(mle.pkappa), ptau=np.copy(mle.ptau), elbo_kt=mle.compute_elbo())
    out["niter_alpha"] = mle.opt_alpha()
    out.update(alpha=np.copy(mle.alpha), elbo_alpha=mle.compute_elbo())
    mle.opt_A_u()
    out.update(A=np.copy(mle.A), u_M=np.copy(mle.u), elbo_M=mle.compute_elbo(),
               gd_coeffConst_M=np.copy(mle.gd_coeffConst))
    return out


def kat_gibbs_seeded(ref):
    """A short *seeded* joint Gibbs-EM run of the reference (tiny synthetic data,
    proper bulk Gibbs).  The oracle must reproduce it bit-for-bit because it makes
    the same numpy.random calls in the same order."""
    sim = orc.simulate(N=24, K=3, L=9, M=5, seed=3, anchor=1, anti=1)
    Y, X, G = sim["Y"], sim["X"], sim["G"]
    np.random.seed(2024)
    gem = ref["gem"].LogitNormalGEM(BKexpr=X, SCexpr=Y, K=3, G=G, burnin=2, sample=3, thin=1,
                                    bk_mean_approx=False, MLE_CONV=1e-6, MLE_maxiter=50,
                                    EM_CONV=1e-6, EM_maxiter=2)
    _, _, _, path = gem.gem()
    st = gem.suff_stats
    return dict(Y=Y, X=X, G=G, seed=2024, path_elbo=path, A=gem.A, alpha=gem.alpha,
                pkappa=gem.pkappa, ptau=gem.ptau, exp_S=st["exp_S"], exp_kappa=st["exp_kappa"],
                exp_tau=st["exp_tau"], coeffA=st["coeffA"], coeffAsq=st["coeffAsq"],
                exp_elbo_const=st["exp_elbo_const"], exp_W=st["exp_W"],
                exp_Zik=st["exp_Zik"])


def kat_sconly(ref, demo):
    """sc-only injected path (hasBK False): init_step = 0.01/L branch, no alpha."""
    Y, G, K = demo["Y"][:20], demo["G"][:20].copy(), 3
    G[:4] = 0
    G[4:9] = 1  # three populated types among the first 20 demo cells
    gem = ref["gem"].LogitNormalGEM(SCexpr=Y, K=K, G=G, MLE_CONV=1e-6, MLE_maxiter=60,
                                    EM_maxiter=1, burnin=1, sample=1)
    np.random.seed(5)
    gem.init_gibbs()
    gem.init_mle()
    sc, mle = gem.Gibbs_SC, gem.mle
    L, N = Y.shape
    rng = np.random.RandomState(77)
    sc.init_suffStats()
    inj = dict(S=[], w=[], kappa=[], tau=[])
    for t in range(3):
        sc.S = ((Y > 0) | (rng.rand(L, N) < 0.3)).astype(int)
        sc.w = rng.gamma(1.5, 0.1, (L, N))
        sc.kappa = rng.normal(-1, 0.2, (1, L))
        sc.tau = rng.normal(N, 2, (1, L))
        sc.update_suffStats(3)
        for k_ in inj:
            inj[k_].append(np.copy(getattr(sc, k_)))
    st = dict(sc.suff_stats)
    out = dict(Y=Y, G=G, init_A=np.copy(gem.init_A), inj_S=np.array(inj["S"]), inj_w=np.array(inj["w"]),
               inj_kappa=np.array(inj["kappa"]), inj_tau=np.array(inj["tau"]))
    for k_, v_ in st.items():
        out["st_" + k_] = np.array(v_)
    mle.update_suff_stats(st)
    out["elbo_E"] = mle.compute_elbo()
    mle.opt_kappa_tau()
    mle.opt_A_u()
    out.update(A=mle.A, pkappa=mle.pkappa, ptau=mle.ptau, elbo_M=mle.compute_elbo())
    return out


def _use_c_pg(ref):
    """Statistical fixtures need long chains: swap the shim's sampler for the compiled
    restatement (oracle/pg_devroye.c) -- same distribution, ~100x faster."""
    ref["e_step_gibbs"].ppg.PyPolyaGamma = orc.CPGSampler
    ref["e_step_gibbs"].ppg.pgdrawvpar = orc.pgdrawv_c


def stat_chains(ref):
    """Long Gibbs chains of the REFERENCE classes on a small problem, several independent
    repeats: posterior-mean statistics with their between-chain spread.  The CUDA chains
    must agree within Monte-Carlo error (north-star check #2/#3)."""
    _use_c_pg(ref)
    sim = orc.simulate(N=60, K=3, L=12, M=6, seed=11, anchor=1, anti=1)
    Y, X, G = sim["Y"], sim["X"], sim["G"]
    itype = [np.where(G == k)[0] for k in range(3)]
    A = orc.init_A(Y, X, 3, itype, 1e-6)
    pk, pt, alpha = np.array([-1.0, 0.25]), np.array([60.0, 36.0]), np.array([1.5, 1.0, 2.0])
    burnin, sample, reps = 200, 3000, 12
    keys_sc = ("exp_S", "exp_kappa", "exp_tau", "exp_kappasq", "exp_tausq", "coeffA", "coeffAsq",
               "exp_elbo_const")
    keys_bk = ("exp_Zik", "exp_Zjk", "exp_logW", "exp_W")
    acc = {k: [] for k in keys_sc + keys_bk}
    for rep in range(reps):
        np.random.seed(1000 + rep)
        sc = ref["e_step_gibbs"].LogitNormalGibbs_SC(A, pk, pt, Y, G, itype)
        sc.init_gibbs()
        sc.gibbs(burnin=burnin, sample=sample, thin=1)
        for k in keys_sc:
            acc[k].append(np.array(sc.suff_stats[k], dtype=float))
        bk = ref["e_step_gibbs"].LogitNormalGibbs_BK(A, alpha, X, None)
        bk.init_gibbs()
        bk.gibbs(burnin=burnin, sample=sample, thin=1, mean_approx=False)
        for k in keys_bk:
            acc[k].append(np.array(bk.suff_stats[k], dtype=float))
    out = dict(Y=Y, X=X, G=G, A=A, pkappa=pk, ptau=pt, alpha=alpha, burnin=burnin, sample=sample,
               reps=reps)
    for k, v in acc.items():
        v = np.array(v)
        out["mean_" + k] = v.mean(axis=0)
        out["sd_" + k] = v.std(axis=0, ddof=1)
    return out


def demo_fits(ref, demo):
    """Reruns of the reference's `make run` configuration (burnin 10 / sample 20 / EM 10,
    joint demo data, default mean-approx bulk): estimates with their run-to-run spread."""
    _use_c_pg(ref)
    keep = dict(A=[], alpha=[], pkappa=[], ptau=[], path=[], exp_W=[], exp_S=[], exp_kappa=[],
                exp_tau=[])
    for rep in range(6):
        np.random.seed(500 + rep)
        gem = ref["gem"].LogitNormalGEM(BKexpr=demo["X"], SCexpr=demo["Y"], K=3, G=demo["G"],
                                        burnin=10, sample=20, thin=1, bk_mean_approx=True,
                                        MLE_CONV=1e-6, MLE_maxiter=500, EM_CONV=1e-6, EM_maxiter=10)
        path = gem.gem()[3]
        keep["A"].append(gem.A)
        keep["alpha"].append(gem.alpha)
        keep["pkappa"].append(gem.pkappa)
        keep["ptau"].append(gem.ptau)
        keep["path"].append(path)
        keep["exp_W"].append(gem.suff_stats["exp_W"])
        keep["exp_S"].append(gem.suff_stats["exp_S"])
        keep["exp_kappa"].append(gem.suff_stats["exp_kappa"])
        keep["exp_tau"].append(gem.suff_stats["exp_tau"])
    out = {}
    for k, v in keep.items():
        v = np.array(v)
        out["mean_" + k] = v.mean(axis=0)
        out["sd_" + k] = v.std(axis=0, ddof=1)
    out["reps"] = 6
    return out


def main():
    args = [x for x in sys.argv[1:] if not x.startswith("--")]
    ref_dir = args[0] if args else "/root/reference"
    ref = load_reference(ref_dir)
    demo = demo_data(ref_dir)
    np.savez_compressed(os.path.join(HERE, "demo_data.npz"), **demo)
    np.savez_compressed(os.path.join(HERE, "kat_simplex.npz"), **kat_simplex(ref))
    np.savez_compressed(os.path.join(HERE, "kat_bulk.npz"), **kat_bulk(ref, demo))
    np.savez_compressed(os.path.join(HERE, "kat_sc.npz"), **kat_sc(ref, demo))
    np.savez_compressed(os.path.join(HERE, "kat_sconly.npz"), **kat_sconly(ref, demo))
    np.savez_compressed(os.path.join(HERE, "kat_gibbs_seeded.npz"), **kat_gibbs_seeded(ref))
    if "--no-stat" not in sys.argv:
        np.savez_compressed(os.path.join(HERE, "stat_chains.npz"), **stat_chains(ref))
        np.savez_compressed(os.path.join(HERE, "stat_demo_fits.npz"), **demo_fits(ref, demo))
    k = np.load(os.path.join(HERE, "kat_sc.npz"))
    b = np.load(os.path.join(HERE, "kat_bulk.npz"))
    print("KAT-B  path_elbo", b["path_elbo"])
    print("KAT-SC elbo_E %.10f elbo_M %.10f" % (k["elbo_E"], k["elbo_M"]))


if __name__ == "__main__":
    main()
