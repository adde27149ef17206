"""Device-side conjugate draws (nhp_cont_resample_params; SURVEY section 8f "next").  The reference's samplers are not
reproducible from uniforms, so parity is distributional: for FIXED statistics the draws must have the moments of the
conjugate posteriors of baselines.jl:72-77, weights.jl:59-64, impulses.jl:68-73 / 204-214, and a chain that keeps the
whole sweep on the device must agree with the host-draw chain within Monte-Carlo error."""
import ctypes

import numpy as np
import pytest

import nhp_b200 as nhp
import synth
from nhp_b200.continuous import _hyper, _read_stats, _resample_parents
from nhp_b200.core import _ptr
from test_chain_gpu import _model
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu
R = 600  # draws per parameter


def _draws(proc, data, R):
    ctx = proc._ctx()
    d = proc.upload(data)
    proc._push(ctx)
    _resample_parents(ctx, d, 3, 0, None, False)
    st = _read_stats(ctx, d, proc.ndims())
    hy = _hyper(proc)
    out = []
    for r in range(R):
        ctx.check(ctx.lib.nhp_cont_resample_params(ctx.h, d.h, 11, r, float(d.duration), _ptr(hy), hy.size, 1))
        nhp.pull_params_(proc, ctx)
        imp = proc.impulses
        out.append((proc.baseline.lam.copy(), proc.weights.W.copy(), imp.p1().copy(), None if imp.p2() is None else imp.p2().copy()))
    return st, [np.array([o[k] for o in out]) if out[0][k] is not None else None for k in range(4)]


def _check_gamma(x, shape, rate, name, nsig=5.0):
    mean, var = shape / rate, shape / rate ** 2
    se = np.sqrt(var / x.shape[0])
    assert np.all(np.abs(x.mean(axis=0) - mean) < nsig * se + 1e-12), name
    # sd of the sample-variance ratio is sqrt((2 + 6/shape) / R) <= 0.115 at R = 600: a 50 % band is > 4 sigma
    assert np.all(np.abs(x.var(axis=0, ddof=1) / var - 1.0) < 0.5), name + " variance"


def test_exponential_posterior_moments():
    K = 3
    proc, _ = make_exp(K, 4, wmax=0.3)
    data = synth.poisson_stream(3000, K, 8.0, 2)
    st, (lam, W, th, _) = _draws(proc, data, R)
    b, w, imp = proc.baseline, proc.weights, proc.impulses
    _check_gamma(lam, b.alpha0 + st["M0"], b.beta0 + data[2], "lambda0")
    _check_gamma(W, w.kappa + st["Mnm"], (w.nu + st["Mn"])[:, None] * np.ones((K, K)), "W")
    _check_gamma(th, imp.alpha + st["Mnm"], imp.beta + st["S1"], "theta")


def test_logitnormal_posterior_moments():
    K = 3
    proc, _ = make_ln(K, 4, wmax=0.3)
    data = synth.poisson_stream(3000, K, 8.0, 2)
    st, (lam, W, mu, tau) = _draws(proc, data, R)
    b, w, imp = proc.baseline, proc.weights, proc.impulses
    _check_gamma(lam, b.alpha0 + st["M0"], b.beta0 + data[2], "lambda0")
    _check_gamma(W, w.kappa + st["Mnm"], (w.nu + st["Mn"])[:, None] * np.ones((K, K)), "W")
    m = st["Mnm"]
    X = st["S1"] / m
    a = imp.alpha0 + m / 2.0
    bb = st["S2"] / 2.0 + m * imp.kappamu / (m + imp.kappamu) * (X - imp.mumu) ** 2 / 2.0
    _check_gamma(tau, a, bb, "tau")
    mun = (imp.kappamu * imp.mumu + m * X) / (imp.kappamu + m)
    var_mu = bb / ((imp.kappamu + m) * (a - 1.0))  # E[1 / (kappa tau)]
    assert np.all(np.abs(mu.mean(axis=0) - mun) < 5.0 * np.sqrt(var_mu / R)), "mu"
    assert np.all(np.abs(mu.var(axis=0, ddof=1) / var_mu - 1.0) < 0.5), "mu variance"


def test_draws_are_keyed_by_seed_and_counter():
    K = 5
    proc, _ = make_ln(K, 1)
    ctx = proc._ctx()
    d = proc.upload(synth.poisson_stream(2000, K, 10.0, 5))
    proc._push(ctx)
    _resample_parents(ctx, d, 3, 0, None, False)
    hy = _hyper(proc)
    got = []
    for seed, counter in ((7, 1), (7, 1), (7, 2), (8, 1)):
        ctx.check(ctx.lib.nhp_cont_resample_params(ctx.h, d.h, seed, counter, float(d.duration), _ptr(hy), hy.size, 1))
        got.append(nhp.pull_params_(proc, ctx))
    # same key -> same draws (S2 is re-accumulated with atomics by the second pass, so equal up to its summation order)
    np.testing.assert_allclose(got[0], got[1], rtol=1e-9)
    assert not np.allclose(got[0], got[2], rtol=1e-3) and not np.allclose(got[0], got[3], rtol=1e-3)
    assert np.all(got[0][:K] > 0) and np.all(np.isfinite(got[0]))
    # the rebuilt tables are the pulled parameters: the log-likelihood agrees with a fresh push of the same values
    ll = ctypes.c_double()
    ctx.check(ctx.lib.nhp_cont_loglik(ctx.h, d.h, 0, ctypes.byref(ll)))
    assert ll.value == pytest.approx(nhp.loglikelihood(proc, d), rel=1e-12)


def test_device_chain_matches_host_chain():
    truth = _model(0)
    t, nodes, T = nhp.rand(truth, 250.0, np.random.default_rng(11))
    nsteps, burn = 600, 100
    dev, host = _model(0), _model(0)
    gd = np.array(nhp.mcmc_(dev, (t, nodes, T), nsteps=nsteps, seed=5, device_draws=True).samples)[burn:]
    gh = np.array(nhp.mcmc_(host, (t, nodes, T), nsteps=nsteps, seed=9).samples)[burn:]
    nb = 10
    se = lambda x: np.std(x.reshape(nb, -1).mean(axis=1), ddof=1) / np.sqrt(nb)
    for k in range(gd.shape[1]):
        tol = 6.0 * np.hypot(se(gd[:, k]), se(gh[:, k])) + 1e-3
        assert abs(gd[:, k].mean() - gh[:, k].mean()) < tol, (k, gd[:, k].mean(), gh[:, k].mean(), tol)


def test_device_chain_network_runs_and_thins():
    K = 6
    proc, _ = make_ln(K, 3, density=0.5, wmax=0.2)
    data = synth.poisson_stream(3000, K, 10.0, 8)
    res = nhp.mcmc_(proc, data, nsteps=12, seed=2, device_draws=True, store_every=5)
    assert len(res.samples) == 3  # sweeps 5, 10 and the last one
    assert all(np.all(np.isfinite(s)) for s in res.samples)
    assert set(np.unique(proc.adjacency_matrix)) <= {0.0, 1.0}
