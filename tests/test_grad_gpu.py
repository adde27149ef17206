"""GPU parity of the analytic log-likelihood gradient (extension for `mle!`; nhp_cont_loglik_grad) against the CPU
oracle's gradient, which tests/test_oracle.py pins by finite differences.  Tolerance: 1e-10 of the plane's scale
(the terms are sums of FP64 products accumulated by atomics in a different order)."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu
G_TOL = 1e-10


def check(proc, om, data, recursive, with_p2):
    t, nodes, T = data
    ll, g = nhp.loglikelihood_gradient(proc, data, recursive=recursive)
    llo, g0, gW, g1, g2 = om.loglik_grad(t, nodes, T, recursive=recursive)
    assert ll == pytest.approx(llo, rel=1e-10)
    assert ll == pytest.approx(nhp.loglikelihood(proc, data, recursive=recursive), rel=1e-12)
    for name, a, b in (("lambda0", g["lambda0"], g0), ("W", g["W"], gW), ("p1", g["p1"], g1)) + ((("p2", g["p2"], g2),) if with_p2 else ()):
        scale = max(np.max(np.abs(b)), 1.0)
        assert np.max(np.abs(a - b)) <= G_TOL * scale, name


@pytest.mark.parametrize("K,n,rate,density", [(3, 400, 6.0, None), (40, 6000, 60.0, None), (40, 6000, 60.0, 0.2), (300, 20000, 200.0, 0.05)])
def test_gradient_logitnormal(K, n, rate, density):
    proc, om = make_ln(K, 11, density=density)
    check(proc, om, synth.poisson_stream(n, K, rate, 3), False, True)


@pytest.mark.parametrize("K,n,rate,density,recursive,dtmax", [(2, 500, 3.0, None, True, np.inf), (2, 500, 3.0, None, False, 2.0), (30, 4000, 20.0, None, True, np.inf),
                                                              (30, 4000, 20.0, 0.3, True, np.inf), (30, 4000, 20.0, 0.3, False, 1.0)])
def test_gradient_exponential(K, n, rate, density, recursive, dtmax):
    proc, om = make_exp(K, 7, density=density, dtmax=dtmax, wmax=0.5 / K)
    check(proc, om, synth.poisson_stream(n, K, rate, 4), recursive, False)


def test_gradient_is_additive_over_time_shards():
    """Two shards with a dtmax halo: the shard gradients add up to the full gradient (the multi-GPU contract)."""
    K, n = 20, 8000
    proc, _ = make_ln(K, 2)
    t, nodes, T = synth.poisson_stream(n, K, 40.0, 9)
    ll, g = nhp.loglikelihood_gradient(proc, (t, nodes, T))
    ctx = proc._ctx()
    cut = n // 2
    halo = int(np.searchsorted(t, t[cut] - 1.0, side="left"))
    d0 = nhp.ContinuousData(ctx, t[:cut], nodes[:cut], T, K)
    d1 = nhp.ContinuousData(ctx, t[halo:], nodes[halo:], T, K, n_halo=cut - halo, index_base=halo, flags=0)
    ll0, g0 = nhp.loglikelihood_gradient(proc, d0)
    ll1, g1 = nhp.loglikelihood_gradient(proc, d1)
    assert ll0 + ll1 == pytest.approx(ll, rel=1e-12)
    for k in ("lambda0", "W", "p1", "p2"):
        np.testing.assert_allclose(g0[k] + g1[k], g[k], rtol=1e-9, atol=1e-9 * max(1.0, np.max(np.abs(g[k]))))


def test_mle_with_analytic_gradient_matches_finite_difference_mle():
    """README example (K = 2 Exponential): both optimisers reach the same optimum; the analytic one needs far fewer sweeps."""
    K = 2
    true = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))),
                                               nhp.DenseWeightModel(0.1 * np.ones((K, K))))
    data = nhp.rand(true, 300.0, np.random.default_rng(0))
    res = {}
    for mode in ("analytic", "finite"):
        proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))),
                                                   nhp.DenseWeightModel(0.1 * np.ones((K, K))))
        res[mode] = nhp.mle_(proc, data, guess=np.full(proc.params().size, 0.5), gradient=mode, max_iter=300)
    assert res["analytic"].maximum == pytest.approx(res["finite"].maximum, abs=1e-3)
    assert res["analytic"].maximum >= nhp.loglikelihood(true, data) - 1e-6


def _grad_close(g, og, tol=G_TOL):
    for name, a, b in (("lambda0", g["lambda0"], og[1]), ("W", g["W"], og[2]), ("p1", g["p1"], og[3])) + ((("p2", g["p2"], og[4]),) if g["p2"] is not None else ()):
        assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1.0), name


@pytest.mark.parametrize("n", [0, 1, 2, 257])
def test_gradient_edge_sizes(n):
    K = 4
    t, nodes, T = synth.poisson_stream(max(n, 1), K, 20.0, 50 + n)
    t, nodes = t[:n], nodes[:n]
    for proc, om, rec in (make_ln(K, 3, density=0.5, wmax=0.2) + (False,), make_exp(K, 3, wmax=0.2) + (True,)):
        ll, g = nhp.loglikelihood_gradient(proc, (t, nodes, T), recursive=rec)
        og = om.loglik_grad(t, nodes, T, recursive=rec)
        assert ll == pytest.approx(og[0], rel=1e-10)
        _grad_close(g, og)


def test_gradient_single_node_ties_and_empty_network():
    t, nodes, T = synth.poisson_stream(2000, 1, 30.0, 7)
    proc, om = make_ln(1, 5, wmax=0.5)
    _grad_close(nhp.loglikelihood_gradient(proc, (t, nodes, T))[1], om.loglik_grad(t, nodes, T))
    # simultaneous events: dt = 0 contributes theta * w for the Exponential (quirk Q9) and its derivative terms
    K = 3
    tt = np.full(120, 1.25)
    nn = (np.arange(120) % K + 1).astype(np.int64)
    pe, oe = make_exp(K, 8, wmax=0.3, dtmax=2.0)
    _grad_close(nhp.loglikelihood_gradient(pe, (tt, nn, 3.0), recursive=False)[1], oe.loglik_grad(tt, nn, 3.0, recursive=False))
    # a network without links: only the compensator terms remain (dW = 0 where A = 0, dlambda0 = sum 1/lambda0 - T)
    lam0, W, mu, tau, _ = synth.ln_params(K, 4)
    A = np.zeros((K, K))
    pn = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W), A,
                                            nhp.BernoulliNetworkModel(0.5, K))
    t3, n3, T3 = synth.poisson_stream(500, K, 10.0, 3)
    ll, g = nhp.loglikelihood_gradient(pn, (t3, n3, T3))
    assert np.all(g["W"] == 0.0) and np.all(g["p1"] == 0.0) and np.all(g["p2"] == 0.0)
    cnt = np.bincount(n3 - 1, minlength=K)
    np.testing.assert_allclose(g["lambda0"], cnt / lam0 - T3, rtol=1e-12)


def test_gradient_null_outputs_and_state():
    import ctypes
    K = 3
    proc, _ = make_ln(K, 1)
    ctx = proc._ctx()
    d = proc.upload(synth.poisson_stream(300, K, 10.0, 1))
    proc._push(ctx)
    ll = ctypes.c_double()
    ctx.check(ctx.lib.nhp_cont_loglik_grad(ctx.h, d.h, 0, ctypes.byref(ll), None, None, None, None))
    assert ll.value == pytest.approx(nhp.loglikelihood(proc, d), rel=1e-12)
    # the gradient planes share the statistics buffers: a parent sweep invalidates them
    nhp.resample_parents(proc, d, seed=1)
    with pytest.raises(nhp.NHPError):
        ctx.check(ctx.lib.nhp_cont_loglik_grad_read(ctx.h, d.h, ctypes.byref(ll), None, None, None, None))


def test_gradient_closed_form_two_events():
    """Same known answer as tests/test_oracle.py::test_gradient_oracle_closed_form_two_events, through the C ABI."""
    lam0, W, th, T = 0.7, 0.4, 1.3, 5.0
    proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.array([lam0])), nhp.ExponentialImpulseResponse(np.array([[th]])),
                                               nhp.DenseWeightModel(np.array([[W]])))
    data = (np.array([1.0, 2.0]), np.array([1, 1]), T)
    l2 = lam0 + W * th * np.exp(-th)
    for rec in (True, False):
        ll, g = nhp.loglikelihood_gradient(proc, data, recursive=rec)
        assert ll == pytest.approx(np.log(lam0) + np.log(l2) - lam0 * T - 2 * W, rel=1e-12)
        assert g["lambda0"][0] == pytest.approx(1 / lam0 + 1 / l2 - T, rel=1e-12)
        assert g["W"][0, 0] == pytest.approx(th * np.exp(-th) / l2 - 2.0, rel=1e-12)
        assert g["p1"][0, 0] == pytest.approx(W * np.exp(-th) * (1 - th) / l2, rel=1e-12)
