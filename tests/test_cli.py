"""scUnif.py contract (scUnif.py:60-249): flags, validation errors (CPU), and -- on a
GPU -- the end-to-end run with the reference's `make run` command line and the
gemout_*.csv layout."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_demo(tmp_path, golden):
    d = golden("demo_data")
    paths = {}
    for key, name in (("Y", "sc.csv"), ("X", "bk.csv"), ("G", "types.csv")):
        p = str(tmp_path / name)
        np.savetxt(p, d[key], delimiter=",", fmt="%d" if key == "G" else "%.18e")
        paths[key] = p
    return d, paths


def test_flags_and_defaults_match_reference():
    from ursm_b200.scUnif import build_parser
    p = build_parser()
    a = p.parse_args([])
    assert (a.number_of_cell_types, a.burn_in_length, a.gibbs_sample_number, a.gibbs_thinning) == (3, 50, 50, 1)
    assert (a.Mstep_maxiter, a.EM_maxiter, a.mininimal_A) == (500, 50, 1e-6)
    assert a.Mstep_convergence_tol == 1e-6 and a.EM_convergence_tol == 1e-6
    assert a.bk_mean_approx is True and a.est_alpha is True
    assert a.output_prefix == "gemout_" and a.output_directory == "out/" and a.verbose_level == 1
    # short and long spellings (the R wrapper uses the long ones, scUnif_wrapper.R:118-124)
    b = p.parse_args(["--single_cell_expr_file", "a", "--bulk_expr_file", "b",
                      "--single_cell_type_file", "c", "--number_of_cell_types", "4",
                      "--burn_in_length", "7", "--gibbs_sample_number", "9", "--EM_maxiter", "2",
                      "--no_mean_approx", "--no_est_alpha", "--initial_kappa_mean_var", "-1", "0.5"])
    c = p.parse_args(["-sc", "a", "-bk", "b", "-ctype", "c", "-K", "4", "-burnin", "7", "-sample", "9",
                      "-EM_maxiter", "2", "-no_mean_approx", "-no_est_alpha", "-pkappa", "-1", "0.5"])
    assert vars(b) == vars(c)
    assert c.bk_mean_approx is False and c.est_alpha is False and c.initial_kappa_mean_var == [-1.0, 0.5]


def _run_cli(args, cwd):
    return subprocess.run([sys.executable, os.path.join(ROOT, "scUnif.py")] + args, cwd=cwd,
                          capture_output=True, text=True, timeout=600)


def test_validation_errors_exit_1(tmp_path, golden):
    """Messages and exit code of scUnif.py:188-219 -- raised before any device work."""
    d, paths = _write_demo(tmp_path, golden)
    log = ["-log", str(tmp_path / "log.txt")]
    r = _run_cli(log, str(tmp_path))
    assert r.returncode == 1 and "Must provide at least one of single cell or bulk data" in r.stderr
    r = _run_cli(["-sc", paths["Y"]] + log, str(tmp_path))
    assert r.returncode == 1 and "Must provide cell type information" in r.stderr
    bad = str(tmp_path / "bk_bad.csv")
    np.savetxt(bad, d["X"][:, :50], delimiter=",")
    r = _run_cli(["-sc", paths["Y"], "-ctype", paths["G"], "-bk", bad] + log, str(tmp_path))
    assert r.returncode == 1 and "must have same number of genes" in r.stderr
    r = _run_cli(["-sc", paths["Y"], "-ctype", paths["G"], "-K", "2"] + log, str(tmp_path))
    assert r.returncode == 1 and "can only take values in {0, ..., K-1}" in r.stderr
    short = str(tmp_path / "types_short.csv")
    np.savetxt(short, d["G"][:10], fmt="%d")
    r = _run_cli(["-sc", paths["Y"], "-ctype", short] + log, str(tmp_path))
    assert r.returncode == 1 and "Mismatched cell dimensions" in r.stderr


@pytest.mark.gpu
def test_make_run_command_line(tmp_path, golden):
    """`make run` of the reference (makefile:22-31) through the drop-in scUnif.py."""
    d, paths = _write_demo(tmp_path, golden)
    out = tmp_path / "out"
    r = _run_cli(["-sc", paths["Y"], "-ctype", paths["G"], "-bk", paths["X"], "-K", "3",
                  "-burnin", "10", "-sample", "20", "-EM_maxiter", "4", "-outdir", str(out),
                  "-log", str(out / "log.log"), "-seed", "5"], str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    L, N = d["Y"].shape
    M = d["X"].shape[0]
    shapes = {"est_A": (N, 3), "exp_S": (L, N), "est_pkappa": (2,), "est_kappa": (L,),
              "est_ptau": (2,), "est_tau": (L,), "est_alpha": (3,), "exp_W": (3, M)}
    got = {}
    for name, shape in shapes.items():
        path = out / ("gemout_%s.csv" % name)
        assert path.exists(), name
        got[name] = np.loadtxt(str(path), delimiter=",")
        assert got[name].shape == shape, (name, got[name].shape)
        assert np.isfinite(got[name]).all()
    assert sorted(os.listdir(str(out))) == sorted(["gemout_%s.csv" % n for n in shapes] + ["log.log"])
    np.testing.assert_allclose(got["est_A"].sum(axis=0), 1.0, rtol=1e-9)
    np.testing.assert_allclose(got["exp_W"].sum(axis=0), 1.0, rtol=1e-9)
    assert got["est_A"].min() >= 1e-6 and got["exp_S"].min() >= 0 and got["exp_S"].max() <= 1
    assert np.all(got["exp_S"][d["Y"] > 0] == 1)
    first = open(str(out / "gemout_est_A.csv")).readline().split(",")[0]
    assert "e" in first and len(first) >= 20        # np.savetxt default '%.18e'
    log = open(str(out / "log.log")).read()
    for needle in ("Gibbs-EM for 3 cell types.", "50 bulk samples on 100 genes are loaded.",
                   "50 single cells on 100 genes are loaded.", "1-th EM iteration started...",
                   "E-step finished: elbo=", "M-step finished: elbo=", "Gibbs-EM finished in"):
        assert needle in log, needle
    # same seed: every E-step's chains are a pure function of (seed, parameters, data), but
    # the statistics are summed with floating-point atomics and the reference's projected-
    # gradient M-step amplifies last-bit differences (its iteration counts change), so a
    # whole fit repeats to Monte-Carlo-level differences, not bitwise
    out2 = tmp_path / "out2"
    r = _run_cli(["-sc", paths["Y"], "-ctype", paths["G"], "-bk", paths["X"], "-K", "3",
                  "-burnin", "10", "-sample", "20", "-EM_maxiter", "4", "-outdir", str(out2),
                  "-log", str(out2 / "log.log"), "-seed", "5"], str(tmp_path))
    assert r.returncode == 0
    S2 = np.loadtxt(str(out2 / "gemout_exp_S.csv"), delimiter=",")
    A2 = np.loadtxt(str(out2 / "gemout_est_A.csv"), delimiter=",")
    assert np.abs(S2 - got["exp_S"]).mean() < 0.05
    assert np.corrcoef(A2.ravel(), got["est_A"].ravel())[0, 1] > 0.9995
    # recovery of the simulation truth (demo_plots.py's eyeball check, numerically)
    assert np.corrcoef(got["est_A"].ravel(), d["A_true"].ravel())[0, 1] > 0.95
    assert np.corrcoef(got["exp_W"].T.ravel(), d["W_true"].ravel())[0, 1] > 0.97
