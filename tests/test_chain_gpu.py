"""Posterior moments from full Gibbs chains (north_star correctness item 4): a chain driven by libnhp
(parent sweep + statistics + adjacency on the GPU, conjugate draws on the host) against a chain driven by
the CPU oracle with the same conjugate-draw code.  Different uniform streams, so agreement is within
Monte-Carlo error."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth

pytestmark = pytest.mark.gpu


def _model(seed):
    rng = np.random.default_rng(seed)
    K = 2
    lam0 = np.array([0.8, 1.2])
    W = np.array([[0.3, 0.1], [0.15, 0.25]])
    mu = np.array([[0.0, 0.5], [-0.5, 0.2]])
    tau = np.array([[1.0, 1.5], [0.8, 1.2]])
    return nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0.copy()), nhp.LogitNormalImpulseResponse(mu.copy(), tau.copy(), 1.0),
                                               nhp.DenseWeightModel(W.copy()))


def _oracle_sweep(proc, t, nodes, T, rng):
    """resample!(process, data) with the oracle computing parents and statistics (continuous.jl:202-208)."""
    imp = proc.impulses
    om = orc.Cont(1, proc.baseline.lam, proc.weights.W, imp.mu, imp.tau, dtmax=imp.dtmax)
    par, pn = om.resample_parents(t, nodes, rng.random(t.size))
    st = orc.suffstats(1, t, nodes, par, pn, proc.ndims(), imp.dtmax)
    proc.baseline.resample_(st["M0"], T, rng)
    proc.weights.resample_(st["Mn"], st["Mnm"], rng)
    proc.impulses.resample_(st["Mnm"], st["S1"], st["S2"], rng)
    return proc.params()


def test_posterior_moments_match_oracle_chain():
    truth = _model(0)
    t, nodes, T = nhp.rand(truth, 250.0, np.random.default_rng(11))
    assert 300 < t.size < 2000
    nsteps, burn = 600, 100
    gpu = _model(0)
    res = nhp.mcmc_(gpu, (t, nodes, T), nsteps=nsteps, seed=5)
    g = np.array(res.samples)[burn:]
    cpu = _model(0)
    rng = np.random.default_rng(6)
    c = np.array([_oracle_sweep(cpu, t, nodes, T, rng) for _ in range(nsteps)])[burn:]
    # params = [lambda0 (2); mu (4); tau (4); W (4)]; compare lambda0 and W (well identified), loosely mu/tau
    def check(idx, nsig):
        for k in idx:
            # batch-means standard error (chains are autocorrelated)
            nb = 10
            se = lambda x: np.std(x.reshape(nb, -1).mean(axis=1), ddof=1) / np.sqrt(nb)
            d = abs(g[:, k].mean() - c[:, k].mean())
            tol = nsig * np.hypot(se(g[:, k]), se(c[:, k])) + 1e-3
            assert d < tol, (k, g[:, k].mean(), c[:, k].mean(), tol)
    check([0, 1], 5.0)
    check(range(10, 14), 5.0)
    check(range(2, 10), 6.0)


def test_mle_readme_example_runs():
    """config 1: README example (README.md:27-38): 2-node exponential process, loglikelihood + mle!."""
    K = 2
    proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))),
                                               nhp.DenseWeightModel(0.1 * np.ones((K, K))))
    t, nodes, T = nhp.rand(proc, 1000.0, np.random.default_rng(0))
    om = orc.Cont(0, np.ones(K), 0.1 * np.ones((K, K)), np.ones((K, K)))
    ll_true = nhp.loglikelihood(proc, (t, nodes, T))
    assert ll_true == pytest.approx(om.loglik(t, nodes, T, recursive=True), rel=1e-10)
    fit = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))),
                                              nhp.DenseWeightModel(0.1 * np.ones((K, K))))
    res = nhp.mle_(fit, (t, nodes, T), max_iter=60, seed=1)
    assert res.maximum >= ll_true - 1e-6          # the optimiser ends at least as high as the generating parameters
    assert np.all(np.abs(fit.baseline.lam - 1.0) < 0.3)


@pytest.mark.parametrize("network", [False, True])
def test_device_trace_holds_the_chain_samples(network):
    """mcmc_device_: the whole chain on the device with the device-side sample trace (nhp_cont_trace_*) gives the same samples as
    the device chain that pulls every sweep's parameters to the host (same Philox keys), in the order of params(process)."""
    import ctypes
    K, nsteps = 5, 7
    lam0, W, mu, tau, A = synth.ln_params(K, 3, wmax=0.4, density=0.6 if network else None)

    def make():
        base, imp, wts = nhp.HomogeneousProcess(lam0.copy()), nhp.LogitNormalImpulseResponse(mu.copy(), tau.copy(), 1.0), nhp.DenseWeightModel(W.copy())
        if network:
            return nhp.ContinuousNetworkHawkesProcess(base, imp, wts, A.copy(), nhp.BernoulliNetworkModel(0.5, K))
        return nhp.ContinuousStandardHawkesProcess(base, imp, wts)

    p0 = make()
    t, nodes, T = nhp.rand(p0, 300.0, np.random.default_rng(4))
    a, b = make(), make()
    ref = nhp.mcmc_(a, (t, nodes, T), nsteps=nsteps, seed=9, device_draws=True)
    got = nhp.mcmc_device_(b, (t, nodes, T), nsteps=nsteps, seed=9)
    assert len(got.samples) == nsteps
    for x, y in zip(ref.samples, got.samples):  # the float statistics are accumulated with atomics: equal up to the summation order
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a.params(), b.params(), rtol=1e-9, atol=1e-12)
    # capacity and state errors
    ctx = b._ctx()
    assert ctx.lib.nhp_cont_trace_push(ctx.h) < 0  # no open trace
    b._push(ctx)
    ctx.check(ctx.lib.nhp_cont_trace_begin(ctx.h, 1))
    ctx.check(ctx.lib.nhp_cont_trace_push(ctx.h))
    assert ctx.lib.nhp_cont_trace_push(ctx.h) < 0  # full
    n, cap = ctypes.c_int64(), ctypes.c_int64()
    ctx.check(ctx.lib.nhp_cont_trace_count(ctx.h, ctypes.byref(n), ctypes.byref(cap)))
    assert (n.value, cap.value) == (1, 1)
    ctx.check(ctx.lib.nhp_cont_trace_free(ctx.h))
