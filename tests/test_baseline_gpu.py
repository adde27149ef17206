"""Inhomogeneous (LogGaussianCoxProcess) baselines inside the sweeps (csrc/cont_baseline.cu; baselines.jl:187-336,
utils/interpolation.jl:27-36).  The oracle restates the homogeneous reference path only, so the checks here are built from it by
linearity (the baseline enters every intensity additively) and from an independent NumPy statement of the formulas."""
import ctypes

import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth

pytestmark = pytest.mark.gpu


def curves(K, G, T, seed):
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, T, G)
    lam = np.exp(rng.normal(0.0, 0.5, (K, 1)) + 0.6 * np.sin(2 * np.pi * rng.uniform(0.5, 2.0, (K, 1)) * x[None, :] / T + rng.uniform(0, 6, (K, 1))))
    return x, lam


def models(K, seed, density, x, lam):
    lam0, W, mu, tau, A = synth.ln_params(K, seed, wmax=1.5 / (K * (density or 1.0)), density=density)
    imp = lambda: nhp.LogitNormalImpulseResponse(mu.copy(), tau.copy(), 1.0)
    if A is None:
        mk = lambda base: nhp.ContinuousStandardHawkesProcess(base, imp(), nhp.DenseWeightModel(W.copy()))
    else:
        mk = lambda base: nhp.ContinuousNetworkHawkesProcess(base, imp(), nhp.DenseWeightModel(W.copy()), A.copy(), nhp.BernoulliNetworkModel(density, K))
    return mk(nhp.LogGaussianCoxProcess(x, lam)), mk(nhp.HomogeneousProcess(np.ones(K))), orc.Cont(1, np.ones(K), W, mu, tau, A=A, dtmax=1.0), (W, mu, tau, A)


@pytest.mark.parametrize("K,n,rate,density,child", [(6, 4000, 30.0, None, False), (40, 6000, 60.0, 0.1, False), (9, 3000, 25.0, None, True)])
def test_intensity_and_loglik_with_grid_baseline(monkeypatch, K, n, rate, density, child):
    """lambda_i(grid) = lambda_i(lambda0 = 1) - 1 + lambda0_{c_i}(t_i), and ll = sum log lambda_i - sum_k trapz - compensator,
    through the dense, the sparse-adjacency and the child-major sweep."""
    if child:
        monkeypatch.setenv("NHP_CHILD", "1")
    t, nodes, T = synth.poisson_stream(n, K, rate, 3)
    x, lam = curves(K, 33, T, 4)
    pg, ph, om, (W, mu, tau, A) = models(K, 5, density, x, lam)
    b_i = np.array([np.interp(ti, x, lam[c - 1]) for ti, c in zip(t, nodes)])
    lam_h = om.event_intensity(t, nodes)
    d = pg.upload((t, nodes, T))
    lam_g = nhp.event_intensity(pg, d)
    np.testing.assert_allclose(lam_g, lam_h - 1.0 + b_i, rtol=1e-11)
    Weff = W if A is None else W * A
    comp = float(np.sum(np.bincount(nodes - 1, minlength=K) * Weff.sum(axis=1)))
    trapz = float(np.sum(0.5 * (lam[:, 1:] + lam[:, :-1]) * np.diff(x)[None, :]))
    ll_ref = float(np.sum(np.log(lam_h - 1.0 + b_i))) - trapz - comp
    assert nhp.loglikelihood(pg, d, recursive=False) == pytest.approx(ll_ref, rel=1e-11)
    # back to a homogeneous model on the same context: the curves are gone
    assert nhp.loglikelihood(ph, (t, nodes, T), recursive=False) == pytest.approx(om.loglik(t, nodes, T, recursive=False), rel=1e-10)
    d.free()


def _ln_pdf(dt, mu, tau, D=1.0):
    z = np.log(dt / (D - dt))
    return np.sqrt(tau / (2 * np.pi)) * np.exp(-0.5 * tau * (z - mu) ** 2) * D / (dt * (D - dt)) * D  # impulses.jl:174-178 with the D^2 of the device table


@pytest.mark.parametrize("density", [None, 0.3])
def test_parents_and_slice_likelihood_with_grid_baseline(density):
    """Parent draws given u against a NumPy statement of parents.jl:25-46 with lambda0_c(t_i) as the last weight; then the
    elliptical-slice likelihood of candidate curves (baselines.jl:247-254) from the device-resident assignment."""
    K, n = 5, 1200
    t, nodes, T = synth.poisson_stream(n, K, 18.0, 8)
    x, lam = curves(K, 21, T, 9)
    pg, _, _, (W, mu, tau, A) = models(K, 10, density, x, lam)
    Weff = W if A is None else W * A
    u = np.random.default_rng(2).random(n)
    ref = np.zeros(n, dtype=np.int64)
    for i in range(n):
        ws, js = [], []
        for j in range(i - 1, -1, -1):
            dt = t[i] - t[j]
            if not (t[j] > t[i] - 1.0):
                break
            p, c = nodes[j] - 1, nodes[i] - 1
            ws.append(Weff[p, c] * _ln_pdf(dt, mu[p, c], tau[p, c]) if 0.0 < dt < 1.0 else 0.0)
            js.append(j + 1)
        b = np.interp(t[i], x, lam[nodes[i] - 1])
        S = sum(ws) + b
        cum, tgt = 0.0, u[i] * S
        for w, j in zip(ws, js):
            cum += w
            if cum > tgt:
                ref[i] = j
                break
    ref[0] = 0
    d = pg.upload((t, nodes, T))
    par, pn = nhp.resample_parents(pg, d, u=u)
    assert np.count_nonzero(par != ref) == 0
    # slice likelihood of candidate curves
    x2, cand = curves(K, 17, T, 12)
    ctx = pg._ctx()
    from nhp_b200.core import _f64, _ptr
    ll = np.empty(K)
    ctx.check(ctx.lib.nhp_cont_baseline_loglik(ctx.h, d.h, x2.size, _ptr(_f64(x2)), _ptr(_f64(cand.ravel())), _ptr(ll)))
    for k in range(K):
        sel = (par == 0) & (nodes == k + 1)
        want = np.sum(np.log(np.interp(t[sel], x2, cand[k]))) - np.sum(0.5 * (cand[k, 1:] + cand[k, :-1]) * np.diff(x2))
        assert ll[k] == pytest.approx(want, rel=1e-12, abs=1e-10)
    d.free()


def test_query_times_support_and_chain():
    K, n = 4, 2500
    t, nodes, T = synth.poisson_stream(n, K, 20.0, 14)
    x, lam = curves(K, 25, T, 15)
    pg, ph, _, _ = models(K, 16, None, x, lam)
    tq = np.linspace(0.0, T, 301)
    got = nhp.intensity(pg, (t, nodes, T), tq)
    base = nhp.intensity(ph, (t, nodes, T), tq)
    want = base - 1.0 + np.stack([np.interp(tq, x, lam[k]) for k in range(K)], axis=1)
    np.testing.assert_allclose(got, want, rtol=1e-11)
    # an event beyond the grid is the reference's DomainError
    short = nhp.LogGaussianCoxProcess(x[:-3], lam[:, :-3])
    pbad = nhp.ContinuousStandardHawkesProcess(short, pg.impulses, pg.weights)
    with pytest.raises(nhp.NHPError):
        nhp.loglikelihood(pbad, (t, nodes, T), recursive=False)
    # the gradient and the adjacency sampler take a homogeneous baseline
    with pytest.raises(nhp.NHPError):
        nhp.loglikelihood_gradient(pg, (t, nodes, T), recursive=False)
    # a few Gibbs sweeps with the elliptical-slice update of the curves
    rng = np.random.default_rng(1)
    d = pg.upload((t, nodes, T))
    ll0 = nhp.loglikelihood(pg, d, recursive=False)
    for s in range(4):
        nhp.resample_(pg, d, rng, seed=3, counter=s)
    assert np.all(np.isfinite(pg.baseline.lam_grid)) and np.all(pg.baseline.lam_grid > 0)
    assert np.isfinite(nhp.loglikelihood(pg, d, recursive=False)) and np.isfinite(ll0)
    d.free()
