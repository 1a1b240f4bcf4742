"""Sampler arithmetic of ursm_b200/csrc/samplers.cuh, compiled for the HOST
(tests/host/host_samplers.cpp includes the very header the kernels inline) and
checked against exact moments / pmfs.  No GPU needed."""
import ctypes

import numpy as np
import pytest
from scipy import stats
from scipy.special import expit

from oracle import ursm_oracle as orc


@pytest.fixture(scope="module")
def lib(hostlib):
    return hostlib


def _dbl(n):
    return (ctypes.c_double * n)()


def test_philox_known_answer(lib):
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0
    out = (ctypes.c_uint32 * 8)()
    lib.h_philox(ctypes.c_uint64(0), 0, 0, 0, out, 2)
    assert [hex(v) for v in out[:4]] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    # second block = counter word 3 incremented, must differ
    assert list(out[4:]) != list(out[:4])
    # all-ones counter/key vector from the same KAT file
    lib.h_philox(ctypes.c_uint64(0xffffffffffffffff), 0xffffffff, 0xffffffff, 0xffffffff, out, 1)
    # c3 starts at 0 in our stream layout, so only check determinism + avalanche here
    a = list(out[:4])
    lib.h_philox(ctypes.c_uint64(0xffffffffffffffff), 0xffffffff, 0xffffffff, 0xfffffffe, out, 1)
    assert a != list(out[:4])


def test_uniform_open_interval(lib):
    n = 200000
    out = _dbl(n)
    lib.h_uniform(n, ctypes.c_uint64(5), out)
    u = np.frombuffer(out, dtype=np.float64)
    assert u.min() > 0 and u.max() < 1
    assert abs(u.mean() - 0.5) < 4 / np.sqrt(12 * n)
    assert stats.kstest(u, "uniform").pvalue > 1e-3


def test_pg_mixture_weight_is_the_envelope_mass_ratio(lib):
    """mass_right = p/(p + q) with p, q the integrals of the envelope pieces the sampler
    proposes from (samplers.cuh header), integrated numerically here."""
    from scipy.integrate import quad
    t = 0.64
    for z in (0.0, 0.1, 0.5, 1.0, 1.5, 1.5624, 1.5626, 2.0, 3.0, 5.0, 8.0):
        fz = np.pi ** 2 / 8 + z * z / 2
        p = quad(lambda x: np.pi / 2 * np.exp(-fz * x), t, np.inf)[0]
        a0 = lambda x: np.pi / 2 * (2 / (np.pi * x)) ** 1.5 * np.exp(-1 / (2 * x))
        if z < 1 / t:   # untilted truncated Levy piece
            q = quad(a0, 0, t)[0]
        else:           # tilted piece on the whole half line (inverse Gaussian)
            q = quad(lambda x: a0(x) * np.exp(-z * z * x / 2), 0, np.inf)[0]
        np.testing.assert_allclose(lib.h_pg_mass(z), p / (p + q), rtol=2e-6)
    assert lib.h_pg_mass(50.0) == 0.0 and lib.h_pg_mass(500.0) == 0.0


def test_ndtri_sq(lib):
    from scipy.special import ndtri
    for v in (1e-8, 1e-6, 1e-4, 0.0016, 0.0018, 0.01, 0.1056, 0.3, 0.5 - 1e-3, 0.7, 0.999, 1 - 1e-6):
        want = ndtri(np.float64(np.float32(v))) ** 2
        np.testing.assert_allclose(lib.h_ndtri_sq(v), want, rtol=(1e-5 if v < 1e-5 else 3e-6), atol=1e-9)


@pytest.mark.parametrize("psi", [0.0, 0.3, 1.0, 2.0, 3.2, 5.0, 9.0, 20.0, 60.0, 300.0, -4.0, 3000.0])
def test_pg1_moments(lib, psi):
    """E, Var and Laplace transform of the fp32 device sampler vs closed form."""
    n = 400000
    out = _dbl(n)
    lib.h_pg_draws(psi, n, ctypes.c_uint64(123 + int(abs(psi) * 10)), out)
    x = np.frombuffer(out, dtype=np.float64)
    m, v = float(orc.pg1_mean(psi)), float(orc.pg1_var(psi))
    assert np.all(x > 0)
    assert abs(x.mean() - m) < 4.5 * np.sqrt(v / n) + 2e-7 * m, (psi, x.mean(), m)
    assert abs(x.var() / v - 1) < 0.02, (psi, x.var(), v)
    for t in (0.5, 4.0):
        a = abs(psi)
        # cosh(a/2)/cosh(b) computed stably for large arguments
        b = np.sqrt((a * a / 2 + t) / 2)
        lap = np.exp(a / 2 - b) * (1 + np.exp(-a)) / (1 + np.exp(-2 * b))
        est = np.exp(-t * x)
        assert abs(est.mean() - lap) < 4.5 * est.std() / np.sqrt(n) + 1e-6 * lap, (psi, t)


def test_pg1_distribution_vs_oracle(lib):
    """Two-sample KS: fp32 Philox sampler vs the float64 oracle restatement."""
    rs = np.random.RandomState(3)
    for psi in (0.5, 2.8, 7.0):
        n = 20000
        out = _dbl(n)
        lib.h_pg_draws(psi, n, ctypes.c_uint64(99), out)
        x = np.frombuffer(out, dtype=np.float64)
        y = np.array([orc.pg1_devroye(psi, rs) for _ in range(6000)])  # Devroye's original
        assert stats.ks_2samp(x, y).pvalue > 1e-3, psi


def test_s_threshold_is_the_bernoulli(lib):
    """s_other > T(u)  <=>  u < expit(psi - R log1p(a/s_other))  (e_step_gibbs.py:337-338)."""
    rs = np.random.RandomState(0)
    agree = total = 0
    for _ in range(20000):
        psi = rs.normal(0, 6)
        a = np.exp(rs.normal(-9, 1.5))
        R = float(rs.choice([0, 1, 50, 4000, 40000]))
        so = np.exp(rs.normal(-1, 1))
        u = rs.uniform(1e-6, 1 - 1e-6)
        T = lib.h_s_threshold(psi, u, a, R)
        b = expit(psi - R * np.log1p(a / so)) if R > 0 else expit(psi)
        if abs(u - b) < 2e-5 * max(b, 1e-3):  # razor edge: fp32 vs fp64 may disagree
            continue
        total += 1
        agree += int((so > T) == (u < b))
    assert total > 15000 and agree == total


@pytest.mark.parametrize("n,p", [(1, 0.3), (7, 0.5), (40, 0.02), (50, 0.9), (300, 0.2),
                                 (5000, 0.013), (100000, 0.4), (2000000, 0.001), (3000, 0.97),
                                 (155, 0.2), (64, 0.5), (1000, 0.031), (20000, 0.5),
                                 # the modal-inversion regime of the bulk split kernel
                                 (50, 0.2), (45, 0.11), (1023, 0.03), (20, 0.5), (10, 0.05),
                                 (3, 0.001), (600, 0.05), (2, 0.5), (1, 0.5), (60, 0.75),
                                 (1023, 0.999), (500, 0.0001)])
@pytest.mark.parametrize("which", ["f64", "f32_state_machine"])
def test_binomial_pmf(lib, n, p, which):
    cnt = 200000
    out = (ctypes.c_int * cnt)()
    fn = lib.h_binomial if which == "f64" else lib.h_binomial_f32
    fn(n, p, cnt, ctypes.c_uint64(n * 7 + 1), out)
    x = np.frombuffer(out, dtype=np.int32)
    assert x.min() >= 0 and x.max() <= n
    assert abs(x.mean() - n * p) < 4.5 * np.sqrt(n * p * (1 - p) / cnt)
    # relative sd of a sample variance: sqrt((kurtosis - 1) / cnt); rare events dominate it
    v = n * p * (1 - p)
    assert abs(x.var() / v - 1) < max(0.03, 4.5 * np.sqrt((2 + (1 - 6 * p * (1 - p)) / v) / cnt))
    # chi-square against the exact pmf over the bulk of the support
    lo, hi = stats.binom.ppf([1e-4, 1 - 1e-4], n, p).astype(int)
    edges = np.unique(np.linspace(lo, hi + 1, 30).astype(int))
    obs = np.histogram(x, bins=np.concatenate([[-1], edges, [n + 1]]) + 0.0)[0]
    cdf = stats.binom.cdf(np.concatenate([edges - 1, [n]]), n, p)
    exp = np.diff(np.concatenate([[0], cdf])) * cnt
    keep = exp > 5
    chi = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stats.chi2.sf(chi, keep.sum() - 1) > 1e-4, (n, p, chi)


def test_binomial_mode_inversion_mass_deficit_is_negligible(lib):
    """The fp32 masses of the modal search sum to 1 - O(1e-6): a uniform beyond the sum is
    retried (unbiased), and that must be rare."""
    lib.h_binom_mode_misses.restype = ctypes.c_int
    lib.h_binom_mode_misses.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_uint64]
    for n, r in ((50, 0.2), (1000, 0.03), (1023, 0.031), (5, 0.5), (300, 0.1), (64, 0.5)):
        miss = lib.h_binom_mode_misses(n, r, 400000, ctypes.c_uint64(3))
        assert miss <= 8, (n, r, miss)   # <= 2e-5


@pytest.mark.parametrize("shape", [0.05, 0.5, 1.0, 2.5, 40.0, 3000.0])
def test_gamma_mt(lib, shape):
    n = 200000
    out = _dbl(n)
    lib.h_gamma(shape, n, ctypes.c_uint64(17), out)
    x = np.frombuffer(out, dtype=np.float64)
    assert abs(x.mean() - shape) < 4.5 * np.sqrt(shape / n)
    assert stats.kstest(x, "gamma", args=(shape,)).pvalue > 1e-4


def test_normal_pair(lib):
    n = 100000
    out = _dbl(2 * n)
    lib.h_normal(n, ctypes.c_uint64(4), out)
    x = np.frombuffer(out, dtype=np.float64)
    assert stats.kstest(x, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(x[0::2], x[1::2])[0, 1]) < 0.02
