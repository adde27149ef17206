"""GPU tests of the device branching simulator (csrc/cont_rand.cu; rand(process, duration), continuous.jl:16-37): parity with
the reference is distributional, so the checks are moments with Monte-Carlo error bars and structural invariants."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu


def stationary_counts(proc, T):
    """E[N_k(T)] ~ ((I - Weff^T)^-1 lambda0)_k T (edge effects are O(dtmax / T))."""
    W = proc.weights.W if proc.adjacency_matrix is None else proc.adjacency_matrix * proc.weights.W
    return np.linalg.solve(np.eye(W.shape[0]) - W.T, proc.baseline.lam) * T


@pytest.mark.parametrize("kind,K,density", [("ln", 5, None), ("ln", 12, 0.4), ("exp", 6, None)])
def test_sample_is_sorted_bounded_and_has_the_stationary_mean(kind, K, density):
    T = 4000.0
    proc, _ = (make_ln(K, 3, density=density, wmax=0.6 / K) if kind == "ln" else make_exp(K, 3, density=density, wmax=0.6 / K, dtmax=np.inf))
    d = nhp.rand_device(proc, T, seed=11)
    t, nodes, dur = d.download()
    assert dur == T and t.size == d.n_own > 1000
    assert np.all(np.diff(t) >= 0) and t[0] >= 0 and t[-1] <= T
    assert nodes.min() >= 1 and nodes.max() <= K
    counts = np.bincount(nodes - 1, minlength=K).astype(float)
    mean = stationary_counts(proc, T)
    # counts of a Hawkes process are over-dispersed: var ~ mean / (1 - branching ratio)^2
    assert np.all(np.abs(counts - mean) < 6.0 * np.sqrt(mean) / (1 - 0.6)), (counts, mean)


def test_device_sample_matches_host_simulator_in_distribution():
    """Two independent implementations of the same cluster process: event counts over many short samples agree."""
    K, T, reps = 3, 60.0, 120
    proc, _ = make_ln(K, 5, wmax=0.25)
    host = np.array([nhp.rand(proc, T, np.random.default_rng(100 + r))[0].size for r in range(reps)], dtype=float)
    dev = np.array([nhp.rand_device(proc, T, seed=500 + r).n_own for r in range(reps)], dtype=float)
    se = np.sqrt(host.var() / reps + dev.var() / reps)
    assert abs(host.mean() - dev.mean()) < 4.5 * se, (host.mean(), dev.mean(), se)
    assert 0.6 < dev.var() / host.var() < 1.6


def test_lags_follow_the_impulse_response():
    """One parent node exciting one child node only: the child's lags to the latest parent event are LogitNormal draws (checked on
    the logit scale: mean mu, variance 1/tau) when parent events are sparse."""
    K = 2
    lam0 = np.array([0.005, 0.0])
    W = np.array([[0.0, 0.9], [0.0, 0.0]])
    mu, tau = np.full((K, K), 0.4), np.full((K, K), 4.0)
    proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W))
    t, nodes, T = nhp.rand_device(proc, 2000000.0, seed=3).download()
    tp, tc = t[nodes == 1], t[nodes == 2]
    assert tc.size > 3000
    idx = np.searchsorted(tp, tc) - 1
    lag = tc - tp[idx]
    ok = (lag > 0) & (lag < 1)
    z = np.log(lag[ok] / (1 - lag[ok]))
    assert ok.mean() > 0.95
    assert abs(z.mean() - 0.4) < 5 * 0.5 / np.sqrt(z.size) + 0.02
    assert abs(z.var() - 0.25) < 0.04  # ~0.5 % of the children sit behind a later parent event: their lag to the latest one is not a draw
    assert abs(tc.size / tp.size - 0.9) < 0.06


def test_loglik_and_chain_on_a_device_sample():
    """The handle nhp_cont_rand returns feeds the sweeps directly; the log-likelihood equals the oracle's on the downloaded copy."""
    K = 8
    proc, om = make_ln(K, 9, density=0.5, wmax=1.0 / K)
    d = nhp.rand_device(proc, 500.0, seed=2)
    t, nodes, T = d.download()
    assert nhp.loglikelihood(proc, d) == pytest.approx(om.loglik(t, nodes, T), rel=1e-10)
    res = nhp.mcmc_(proc, d, nsteps=4, seed=1, device_draws=True)
    assert all(np.all(np.isfinite(s)) for s in res.samples)


def test_sample_too_large_is_refused():
    proc, _ = make_ln(3, 1, wmax=0.1)
    with pytest.raises(nhp.NHPError):
        nhp.rand_device(proc, 1000.0, seed=1, max_events=100)
