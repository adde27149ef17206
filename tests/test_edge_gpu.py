"""GPU edge cases of the continuous path: empty / single-event / K = 1 data, ties, tile-boundary sizes, windows
covering the whole history, extreme impulse parameters.  Every case is compared with the oracle."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu


def _check_all(proc, om, t, nodes, T, kind, dtmax, sparse_modes=("0",), monkeypatch=None):
    n = len(t)
    u = np.random.default_rng(1).random(n)
    for mode in sparse_modes:
        if monkeypatch is not None:
            monkeypatch.setenv("NHP_SPARSE", mode)
        d = proc.upload((t, nodes, T))
        assert nhp.loglikelihood(proc, d, recursive=False) == pytest.approx(om.loglik(t, nodes, T, recursive=False), rel=1e-10)
        np.testing.assert_allclose(nhp.event_intensity(proc, d), om.event_intensity(t, nodes), rtol=1e-10)
        par, pn = nhp.resample_parents(proc, d, u=u)
        opar, opn = om.resample_parents(t, nodes, u) if n else (np.zeros(0, np.int64), np.zeros(0, np.int64))
        np.testing.assert_array_equal(par, opar)
        np.testing.assert_array_equal(pn, opn)
        st = nhp.sufficient_statistics(proc, d)
        ost = orc.suffstats(1 if kind == "ln" else 0, t, nodes, opar, opn, proc.ndims(), dtmax)
        for key in ("M0", "Mn", "Mnm"):
            np.testing.assert_array_equal(st[key], ost[key])
        np.testing.assert_allclose(st["S1"], ost["S1"], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 65, 255, 256, 257, 1023, 1025])
def test_sizes_around_tile_boundaries(n, monkeypatch):
    K = 5
    t, nodes, T = synth.poisson_stream(max(n, 1), K, 20.0, 100 + n)
    t, nodes = t[:n], nodes[:n]
    proc, om = make_ln(K, 3, density=0.5, wmax=0.2)
    _check_all(proc, om, t, nodes, T, "ln", 1.0, sparse_modes=("0", "1"), monkeypatch=monkeypatch)


def test_single_node_process():
    t, nodes, T = synth.poisson_stream(3000, 1, 30.0, 7)
    proc, om = make_ln(1, 5, wmax=0.5)
    _check_all(proc, om, t, nodes, T, "ln", 1.0)
    pe, oe = make_exp(1, 6, wmax=0.5, dtmax=0.7)
    _check_all(pe, oe, t, nodes, T, "exp", 0.7)


def test_all_events_at_the_same_time():
    K = 3
    t = np.full(200, 1.25)
    nodes = (np.arange(200) % K + 1).astype(np.int64)
    proc, om = make_exp(K, 8, wmax=0.3, dtmax=2.0)   # dt = 0 contributes theta*w for Exponential (quirk Q9)
    _check_all(proc, om, t, nodes, 3.0, "exp", 2.0)
    pl, ol = make_ln(K, 9, wmax=0.3)                   # ... and exactly 0 for LogitNormal
    assert nhp.loglikelihood(pl, (t, nodes, 3.0)) == pytest.approx(ol.loglik(t, nodes, 3.0), rel=1e-12)
    par, _ = nhp.resample_parents(pl, (t, nodes, 3.0), seed=1)
    assert np.all(par == 0)


def test_window_covers_whole_history():
    K, n = 4, 1500
    t, nodes, T = synth.poisson_stream(n, K, 10.0, 12)
    proc, om = make_ln(K, 13, wmax=0.001, dtmax=10.0 * T)  # every earlier event is inside every window
    _check_all(proc, om, t, nodes, T, "ln", 10.0 * T)


def test_extreme_impulse_parameters():
    K, n = 3, 4000
    t, nodes, T = synth.poisson_stream(n, K, 40.0, 14)
    rng = np.random.default_rng(3)
    lam0 = np.array([1e-3, 5.0, 0.7])
    W = rng.uniform(0, 0.5, (K, K))
    for mu, tau in ((np.full((K, K), 9.0), np.full((K, K), 50.0)), (np.full((K, K), -12.0), np.full((K, K), 1e-3)), (rng.normal(size=(K, K)) * 4, rng.uniform(1e-2, 1e2, (K, K)))):
        proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W))
        om = orc.Cont(1, lam0, W, mu, tau, dtmax=1.0)
        assert nhp.loglikelihood(proc, (t, nodes, T)) == pytest.approx(om.loglik(t, nodes, T), rel=1e-10)
    for theta in (np.full((K, K), 500.0), np.full((K, K), 1e-4), rng.uniform(1e-3, 1e3, (K, K))):
        proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.ExponentialImpulseResponse(theta, dtmax=2.0), nhp.DenseWeightModel(W))
        om = orc.Cont(0, lam0, W, theta, dtmax=2.0)
        assert nhp.loglikelihood(proc, (t, nodes, T), recursive=False) == pytest.approx(om.loglik(t, nodes, T, recursive=False), rel=1e-10)


def test_empty_data_everywhere():
    K = 3
    proc, om = make_ln(K, 2, density=0.5, wmax=0.2)
    empty = (np.zeros(0), np.zeros(0, np.int64), 4.0)
    assert nhp.loglikelihood(proc, empty) == pytest.approx(-np.sum(proc.baseline.lam) * 4.0, rel=1e-15)
    assert nhp.event_intensity(proc, empty).shape == (0,)
    par, pn = nhp.resample_parents(proc, empty, seed=1)
    assert par.shape == (0,) and pn.shape == (0,)
    st = nhp.sufficient_statistics(proc, empty, parents=np.zeros(0, np.int64))
    assert np.all(st["Mnm"] == 0) and np.all(st["M0"] == 0)
    lam = nhp.intensity(proc, empty, np.array([0.5, 1.0]))
    np.testing.assert_allclose(lam, np.tile(proc.baseline.lam, (2, 1)))
    proc.network = nhp.BernoulliNetworkModel(0.25, K)
    u = np.random.default_rng(0).random((K, K))
    A = nhp.resample_adjacency_matrix_(proc, empty, u=u)
    np.testing.assert_array_equal(A, (u <= 0.25).astype(float))  # no data: the posterior of A is its prior
