"""oracle/pg_devroye.c (the compiled restatement of pypolyagamma's PG(1,z) sampler
used for the CPU baseline) against closed-form moments and the Python restatement."""
import numpy as np
from scipy import stats

from oracle import ursm_oracle as orc


def _draw(psi, n, nthreads=1, seed=5):
    samplers = [orc.CPGSampler(seed + t) for t in range(nthreads)]
    out = np.empty(n)
    orc.pgdrawv_c(samplers, np.ones(n), np.full(n, float(psi)), out)
    return out


def test_c_mass_matches_python():
    for z in (0.0, 0.3, 1.0, 1.5625, 2.0, 5.0, 20.0):
        np.testing.assert_allclose(orc._pg_c().pg_mass_texpon(z), orc.pg_mass_texpon(z), rtol=1e-12)


def test_c_sampler_moments():
    n = 300000
    for psi in (0.0, 0.5, 2.0, 3.3, 8.0, 40.0, -6.0):
        x = _draw(psi, n)
        m, v = float(orc.pg1_mean(psi)), float(orc.pg1_var(psi))
        assert abs(x.mean() - m) < 4.5 * np.sqrt(v / n), (psi, x.mean(), m)
        assert abs(x.var() / v - 1) < 0.02, (psi, x.var(), v)
        t = 1.7
        a = abs(psi)
        b = np.sqrt((a * a / 2 + t) / 2)
        lap = np.exp(a / 2 - b) * (1 + np.exp(-a)) / (1 + np.exp(-2 * b))
        est = np.exp(-t * x)
        assert abs(est.mean() - lap) < 4.5 * est.std() / np.sqrt(n)


def test_c_sampler_vs_python_ks_and_threads():
    rs = np.random.RandomState(1)
    for psi in (1.0, 4.0):
        y = np.array([orc.pg1_devroye(psi, rs) for _ in range(5000)])
        x = _draw(psi, 40000, nthreads=min(4, orc.pg_c_max_threads()))
        assert stats.ks_2samp(x, y).pvalue > 1e-3
