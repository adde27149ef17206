"""GPU parity tests: intensity(process, data, times) (continuous.jl:76-96) and the adjacency Gibbs
sampler (continuous.jl:444-519) against the oracle's literal restatement."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,K,density", [("ln", 3, None), ("ln", 40, 0.3), ("exp", 5, None), ("exp", 300, 0.2)])
def test_intensity_at_query_times(kind, K, density):
    n = 4000
    t, nodes, T = synth.poisson_stream(n, K, 50.0, 11)
    proc, om = (make_ln(K, 3, density=density) if kind == "ln" else make_exp(K, 3, density=density, wmax=0.5 / K, dtmax=1.5))
    rng = np.random.default_rng(1)
    tq = np.sort(np.concatenate([rng.uniform(0, T, 200), t[:5], [0.0, T, T + 5.0]]))  # includes exact event times (strict window)
    lam = nhp.intensity(proc, (t, nodes, T), tq)
    ref = om.intensity(t, nodes, tq)
    assert lam.shape == (tq.size, K)
    np.testing.assert_allclose(lam, ref, rtol=1e-10)
    one = nhp.intensity(proc, (t, nodes, T), float(tq[7]))
    np.testing.assert_allclose(one, ref[7], rtol=1e-10)
    with pytest.raises(ValueError):
        nhp.intensity(proc, (t, nodes, T), -1.0)


def test_intensity_full_history_exponential():
    K, n = 4, 3000
    t, nodes, T = synth.poisson_stream(n, K, 20.0, 2)
    proc, om = make_exp(K, 5, wmax=0.2)  # dtmax = Inf: full history (cut-off horizon inside 1e-14)
    tq = np.linspace(0.0, T, 101)
    np.testing.assert_allclose(nhp.intensity(proc, (t, nodes, T), tq), om.intensity(t, nodes, tq), rtol=1e-10)


@pytest.mark.parametrize("kind,K,n,rate,rho", [("ln", 4, 600, 12.0, 0.5), ("ln", 9, 1500, 40.0, 0.2), ("exp", 5, 800, 15.0, 0.4)])
def test_adjacency_sampler_matches_oracle(kind, K, n, rate, rho):
    t, nodes, T = synth.poisson_stream(n, K, rate, 21 + K)
    if kind == "ln":
        proc, om = make_ln(K, 7, density=0.5, wmax=1.5 / K)
    else:
        proc, om = make_exp(K, 7, density=0.5, wmax=1.5 / K, dtmax=1.0)
    proc.network = nhp.BernoulliNetworkModel(rho, K)
    A0 = proc.adjacency_matrix.copy()
    mismatches = 0
    for rep in range(3):
        u = np.random.default_rng(100 + rep).random((K, K))
        proc.adjacency_matrix = A0.copy()
        A_gpu = nhp.resample_adjacency_matrix_(proc, (t, nodes, T), u=u).copy()
        A_ref = om.resample_adjacency(A0, np.full((K, K), rho), t, nodes, T, u)
        mismatches += int(np.count_nonzero(A_gpu != A_ref))
    assert mismatches == 0
    assert set(np.unique(proc.adjacency_matrix)) <= {0.0, 1.0}


def test_adjacency_sampler_updates_context_tables():
    """After the sweep the log-likelihood uses the new adjacency matrix."""
    K, n = 6, 1000
    t, nodes, T = synth.poisson_stream(n, K, 20.0, 4)
    proc, om = make_ln(K, 9, density=0.5, wmax=1.0 / K)
    d = proc.upload((t, nodes, T))
    nhp.resample_adjacency_matrix_(proc, d, seed=5, counter=1)
    om2 = orc.Cont(1, proc.baseline.lam, proc.weights.W, proc.impulses.mu, proc.impulses.tau, A=proc.adjacency_matrix, dtmax=1.0)
    assert nhp.loglikelihood(proc, d) == pytest.approx(om2.loglik(t, nodes, T), rel=1e-10)


def test_adjacency_dense_network_keeps_all_links():
    K, n = 4, 500
    t, nodes, T = synth.poisson_stream(n, K, 10.0, 8)
    proc, _ = make_ln(K, 2, density=0.5, wmax=0.3)
    proc.network = nhp.DenseNetworkModel(K)  # link probability 1 -> ll0 = -Inf -> A = 1 (networks.jl:20-32)
    A = nhp.resample_adjacency_matrix_(proc, (t, nodes, T), seed=1)
    assert np.all(A == 1.0)


def test_full_gibbs_sweep_runs_and_moves_parameters():
    K = 3
    proc, _ = make_ln(K, 12, density=0.7, wmax=0.4)
    t, nodes, T = nhp.rand(proc, 300.0, np.random.default_rng(2))
    res = nhp.mcmc_(proc, (t, nodes, T), nsteps=5, seed=3)
    assert len(res.samples) == 5 and all(np.all(np.isfinite(s)) for s in res.samples)
    assert not np.allclose(res.samples[0], res.samples[-1])


def test_config5_shape_large_K_exponential_recursive():
    """config 5 shape (K = 5000 Exponential standard process, full-history semantics) at an oracle-sized N:
    the cut-off-horizon sweep against the reference's O(N K) recursion."""
    K, n = 5000, 20000
    t, nodes, T = synth.poisson_stream(n, K, 3.2, 17)
    proc, om = make_exp(K, 18, wmax=0.5 / K)
    d = proc.upload((t, nodes, T))
    assert nhp.loglikelihood(proc, d, recursive=True) == pytest.approx(om.loglik(t, nodes, T, recursive=True), rel=1e-10)
    u = np.random.default_rng(2).random(n)
    par, pn = nhp.resample_parents(proc, d, u=u)
    assert par.shape == (n,) and np.all(par < np.arange(1, n + 1))


def test_adjacency_column_partition_reproduces_full_sweep():
    """Multi-GPU form: ranks own interleaved columns; merging the owned columns equals the full sweep."""
    K, n = 7, 1200
    t, nodes, T = synth.poisson_stream(n, K, 30.0, 33)
    proc, om = make_ln(K, 34, density=0.5, wmax=1.0 / K)
    proc.network = nhp.BernoulliNetworkModel(0.3, K)
    A0 = proc.adjacency_matrix.copy()
    u = np.random.default_rng(8).random((K, K))
    d = proc.upload((t, nodes, T))
    full = nhp.resample_adjacency_matrix_(proc, d, u=u).copy()
    merged = A0.copy()
    R = 3
    for r in range(R):
        proc.adjacency_matrix = A0.copy()
        part = nhp.resample_adjacency_matrix_(proc, d, u=u, col_begin=r, col_stride=R)
        other = [c for c in range(K) if c % R != r]
        np.testing.assert_array_equal(part[:, other], A0[:, other])  # columns of other ranks untouched
        merged[:, r::R] = part[:, r::R]
    np.testing.assert_array_equal(merged, full)
    np.testing.assert_array_equal(full, om.resample_adjacency(A0, np.full((K, K), 0.3), t, nodes, T, u))


def test_adjacency_cache_survives_a_moving_exponential_horizon(monkeypatch):
    """Exponential with dtmax = Inf: the cut-off horizon moves with the parameters, the cached pair structure (built with a
    margin) is reused on one resident data handle, and the matrices still equal the oracle's (full history) given the uniforms."""
    K, n, rho = 5, 900, 0.4
    t, nodes, T = synth.poisson_stream(n, K, 15.0, 31)
    proc, _ = make_exp(K, 7, density=0.5, wmax=1.5 / K)   # dtmax = Inf
    proc.network = nhp.BernoulliNetworkModel(rho, K)
    d = proc.upload((t, nodes, T))
    A0 = proc.adjacency_matrix.copy()
    theta0 = proc.impulses.theta.copy()
    for rep, scale in enumerate((1.0, 1.1, 0.95, 1.3)):   # theta changes => horizon changes (within and beyond the margin)
        proc.impulses.theta = theta0 * scale
        proc.adjacency_matrix = A0.copy()
        u = np.random.default_rng(200 + rep).random((K, K))
        A_gpu = nhp.resample_adjacency_matrix_(proc, d, u=u).copy()
        om = orc.Cont(0, proc.baseline.lam, proc.weights.W, proc.impulses.theta, A=A0, dtmax=np.inf)
        A_ref = om.resample_adjacency(A0, np.full((K, K), rho), t, nodes, T, u)
        assert np.array_equal(A_gpu, A_ref), rep
