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
