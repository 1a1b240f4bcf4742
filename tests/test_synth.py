"""the synthetic-data generators against moments of the REFERENCE generator
(tests/golden/stat_synth.npz, produced by running demo/demo_simulate_data.py:45-114 itself,
see tests/golden/make_synth_golden.py).  `oracle.simulate` is checked on the CPU, the device
generator `ursm_b200.synth` on the GPU.  Tolerances are two-sample z-scores."""
import numpy as np
import pytest

from oracle import ursm_oracle as orc


def _z(a, va, na, b, vb, nb):
    """z-scores of the difference of two sample means with sample variances va, vb."""
    sd = np.sqrt(va / na + vb / nb)
    ok = sd > 0
    return ((a - b)[ok] / sd[ok]).ravel()


def _check(g, X, Y, S, G, kind):
    N, K = int(g["N"]), int(g["K"])
    M, L = X.shape[0], Y.shape[0]
    # bulk: depth ~ Poisson(50 N); per-gene mean and variance of the counts
    assert abs(X.sum(axis=1).mean() - 50 * N) < 5 * np.sqrt(50 * N / M)
    assert abs(X.sum(axis=1).var() / (50 * N) - 1) < 0.2
    z = _z(X.mean(axis=0), X.var(axis=0), M, g["bk_mean"], g["bk_var"], int(g["M"]))
    # (genes move together through the Dirichlet mixing weights: the z-scores are strongly
    # correlated across genes, so only a loose bound on their mean square)
    assert np.abs(z).max() < 5.5 and np.mean(z * z) < 3.0, (kind, np.abs(z).max(), np.mean(z * z))
    r = X.var(axis=0)[g["bk_var"] > 1] / g["bk_var"][g["bk_var"] > 1]
    assert abs(np.median(r) - 1) < 0.1
    # single cells: depth ~ Poisson(2 N); dropout indicator rate, zero rate and mean count per
    # gene and type
    assert abs(Y.sum(axis=1).mean() - 2 * N) < 5 * np.sqrt(2 * N / L)
    for k in range(K):
        rows = G == k
        n, nref = rows.sum(), int(g["cells_per_type"][k])
        p, pref = S[rows].mean(axis=0), g["sc_S_rate"][k]
        z = _z(p, p * (1 - p), n, pref, pref * (1 - pref), nref)
        assert np.abs(z).max() < 5.5, (kind, "S rate", k, np.abs(z).max())
        q, qref = (Y[rows] == 0).mean(axis=0), g["sc_zero_rate"][k]
        z = _z(q, q * (1 - q), n, qref, qref * (1 - qref), nref)
        assert np.abs(z).max() < 5.5, (kind, "zero rate", k, np.abs(z).max())
        z = _z(Y[rows].mean(axis=0), Y[rows].var(axis=0), n, g["sc_mean"][k], g["sc_var"][k], nref)
        assert np.abs(z).max() < 5.5 and np.mean(z * z) < 3.0, (kind, "mean", k, np.mean(z * z))
    assert np.all(S[Y > 0] == 1)   # reads only where the gene is "on"


def test_oracle_simulate_matches_reference_generator(golden):
    g = golden("stat_synth")
    N, K = int(g["N"]), int(g["K"])
    sim = orc.simulate(N=N, K=K, L=2400, M=2400, seed=0)
    # orc.simulate draws its own A: rebuild the counts on the fixture's A
    rs = np.random.RandomState(5)
    A = g["A"]
    W = rs.dirichlet(np.full(K, 2.0), 2400)
    X = np.stack([rs.multinomial(rs.poisson(50 * N), w.dot(A.T)) for w in W]).astype(float)
    G = np.arange(2400) % K
    kap, tau = rs.normal(-1, 0.5, 2400), rs.normal(1.5 * N, 0.15 * N, 2400)
    from scipy.special import expit
    S = (rs.random_sample((2400, N)) < expit(A[:, G].T * tau[:, None] + kap[:, None])).astype(int)
    Y = np.stack([rs.multinomial(max(rs.poisson(2 * N), 1), A[:, G[l]] * S[l] / (A[:, G[l]] * S[l]).sum())
                  for l in range(2400)]).astype(float)
    _check(g, X, Y, S, G, "numpy restatement")
    # and the packaged simulate(): same depth laws and invariants on its own A
    assert abs(sim["X"].sum(axis=1).mean() - 50 * N) < 5 * np.sqrt(50 * N / 2400)
    assert abs(sim["Y"].sum(axis=1).mean() - 2 * N) < 5 * np.sqrt(2 * N / 2400)
    assert np.all(sim["S"][sim["Y"] > 0] == 1)
    np.testing.assert_allclose(sim["A"].sum(axis=0), 1.0, rtol=1e-12)


@pytest.mark.gpu
def test_device_generator_matches_reference_generator(golden):
    from ursm_b200 import synth
    g = golden("stat_synth")
    N, K = int(g["N"]), int(g["K"])
    A = synth.profile_matrix(N, K, seed=0, anchor=2, anti=2)
    np.testing.assert_array_equal(A, g["A"])          # the fixture was drawn on this matrix
    M = L = 3000
    X, W = synth.bulk_on_device(A, M, seed=3)
    G = (np.arange(L) % K).astype(np.int64)
    Y, S, kap, tau = synth.cells_on_device(A, G, seed=3)
    X, Y, S = X.double().cpu().numpy(), Y.double().cpu().numpy(), S.cpu().numpy().astype(int)
    assert abs(kap.mean() + 1) < 5 * 0.5 / np.sqrt(L) and abs(kap.std() - 0.5) < 0.05
    assert abs(tau.mean() / (1.5 * N) - 1) < 5 * 0.1 / np.sqrt(L) and abs(tau.std() / (0.15 * N) - 1) < 0.1
    np.testing.assert_allclose(W.sum(axis=1), 1.0, rtol=1e-12)
    _check(g, X, Y, S, G, "device generator")
    # sharded generation: rows are keyed by (seed, row_offset) blocks -- different offsets give
    # different cells, the same offset the same cells
    Y2 = synth.cells_on_device(A, G[:64], seed=3, row_offset=0)[0].double().cpu().numpy()
    np.testing.assert_array_equal(Y2, synth.cells_on_device(A, G[:64], seed=3, row_offset=0)[0]
                                  .double().cpu().numpy())


@pytest.mark.gpu
def test_device_generator_degenerate_cells():
    """Every cell gets at least one read (the reference divides by the depth, m_step.py:125): cells
    whose Poisson draws all came up zero get one read on their most probable gene, cells without
    any expressed gene fall back to the profile itself; shards of rows do not depend on who draws
    them (row_offset keys the streams)."""
    from ursm_b200 import synth
    A = synth.profile_matrix(64, 3, seed=1)
    G = (np.arange(300) % 3).astype(np.int64)
    Y, S, kap, tau = synth.cells_on_device(A, G, seed=2, sc_depth=0.002)      # ~0.13 reads per cell
    Y, S = Y.double().cpu().numpy(), S.cpu().numpy()
    assert np.all(Y.sum(axis=1) >= 1) and (Y.sum(axis=1) == 1).mean() > 0.7
    one = Y.sum(axis=1) == 1
    empty_fill = one & (S.sum(axis=1) > 0)
    # the filled read sits on an expressed gene
    assert np.all((Y[empty_fill] > 0) <= (S[empty_fill] > 0))
    Yd, Sd, _, _ = synth.cells_on_device(A, G, seed=2, kappa=-60.0, kappa_sd=0.0, tau_scale=0.0)
    assert Sd.sum().item() == 0 and np.all(Yd.double().cpu().numpy().sum(axis=1) >= 1)
    # rows 100..199 drawn as their own shard equal rows 100..199 of the whole
    Y2 = synth.cells_on_device(A, G[100:200], seed=2, sc_depth=0.002, row_offset=100)[0]
    np.testing.assert_array_equal(Y2.double().cpu().numpy(), Y[100:200])
    X, W = synth.bulk_on_device(A, 40, seed=2)
    X2, W2 = synth.bulk_on_device(A, 15, seed=2, row_offset=25)
    np.testing.assert_array_equal(X2.cpu().numpy(), X[25:40].cpu().numpy())
    np.testing.assert_array_equal(W2, W[25:40])
