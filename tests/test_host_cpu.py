"""CPU tests of the host-side mirror (no GPU, no libnhp calls): parameter vectors, the cluster simulator and the
conjugate draws that stay on the host."""
import numpy as np
import pytest

import nhp_b200 as nhp
from nhp_b200 import discrete as D


def _std(K=3, kind="ln"):
    rng = np.random.default_rng(0)
    base = nhp.HomogeneousProcess(rng.uniform(0.5, 1.5, K))
    imp = (nhp.LogitNormalImpulseResponse(rng.normal(size=(K, K)), rng.uniform(0.5, 2, (K, K)), 1.0) if kind == "ln"
           else nhp.ExponentialImpulseResponse(rng.uniform(0.5, 2, (K, K))))
    return nhp.ContinuousStandardHawkesProcess(base, imp, nhp.DenseWeightModel(rng.uniform(0, 0.2, (K, K))))


@pytest.mark.parametrize("kind", ["ln", "exp"])
def test_params_roundtrip_matches_julia_vec_order(kind):
    p = _std(3, kind)
    x = p.params()
    K = 3
    assert x.size == K + (2 if kind == "ln" else 1) * K * K + K * K  # [baseline; impulses; weights]  continuous.jl:116-119
    # vec() is column-major: the second entry of the impulse block is element [2,1] (1-based) = [1,0]
    first = p.impulses.mu if kind == "ln" else p.impulses.theta
    assert x[K + 1] == first[1, 0]
    y = np.random.default_rng(1).uniform(0.1, 1.0, x.size)
    p.params_(y)
    np.testing.assert_array_equal(p.params(), y)
    with pytest.raises(ValueError):
        p.impulses.params_(np.ones(5))  # impulses.jl:44-45 / 155-156
    with pytest.raises(ValueError):
        p.weights.params_(np.ones(5))   # weights.jl:10-11


def test_network_params_layout():
    p = _std(2)
    A = np.array([[1.0, 0.0], [1.0, 1.0]])
    net = nhp.ContinuousNetworkHawkesProcess(p.baseline, p.impulses, p.weights, A, nhp.BernoulliNetworkModel(0.3, 2))
    x = net.params()  # [rho; lambda0; W; theta; vec(A)]  continuous.jl:325-333
    assert x[0] == 0.3 and x.size == 1 + 2 + 4 + 8 + 4
    np.testing.assert_array_equal(x[-4:], [1.0, 1.0, 0.0, 1.0])
    assert net.isstable()


def test_rand_cluster_simulator_rate():
    """Stationary rate of a Hawkes process: (I - W^T)^-1 lambda0; the simulator follows continuous.jl:16-37."""
    p = _std(2)
    p.weights.W = np.array([[0.2, 0.1], [0.05, 0.3]])
    p.baseline.lam = np.array([1.0, 0.5])
    t, nodes, T = nhp.rand(p, 4000.0, np.random.default_rng(3))
    assert np.all(np.diff(t) >= 0) and nodes.min() >= 1 and nodes.max() <= 2 and T == 4000.0
    expect = np.linalg.solve(np.eye(2) - p.weights.W.T, p.baseline.lam)
    got = np.bincount(nodes - 1, minlength=2) / T
    np.testing.assert_allclose(got, expect, rtol=0.08)


def test_conjugate_draws_follow_survey_section_10():
    rng = np.random.default_rng(5)
    K = 2
    base = nhp.HomogeneousProcess(np.ones(K))
    M0 = np.array([400.0, 100.0])
    draws = []
    for _ in range(400):
        base.resample_(M0, 200.0, rng)
        draws.append(base.lam.copy())
    # lambda_k ~ Gamma(alpha0 + M0, 1/(beta0 + T))  (baselines.jl:72-77)
    np.testing.assert_allclose(np.mean(draws, axis=0), (1 + M0) / (1 + 200.0), rtol=0.02)
    w = nhp.DenseWeightModel(np.ones((K, K)))
    Mn, Mnm = np.array([50.0, 10.0]), np.array([[20.0, 5.0], [1.0, 0.0]])
    dw = []
    for _ in range(600):
        w.resample_(Mn, Mnm, rng)
        dw.append(w.W.copy())
    # W[p,c] ~ Gamma(kappa + Mnm, 1/(nu + Mn[p]))  (weights.jl:59-64): the rate uses the PARENT's event count
    np.testing.assert_allclose(np.mean(dw, axis=0), (1 + Mnm) / (1 + Mn)[:, None], rtol=0.08)
    ln = nhp.LogitNormalImpulseResponse(np.zeros((K, K)), np.ones((K, K)), 1.0)
    M = np.array([[30.0, 0.0], [4.0, 9.0]])
    ln.resample_(M, M * 0.3, M * 0.5, rng)  # statistics with empty cells: NaN guards of impulses.jl:207,210 (quirk Q5)
    assert np.all(np.isfinite(ln.mu)) and np.all(ln.tau > 0)


def test_component_validation():
    with pytest.raises(ValueError):
        nhp.HomogeneousProcess([1.0, -0.1])            # DomainError baselines.jl:32
    with pytest.raises(ValueError):
        nhp.HomogeneousProcess([1.0], alpha0=0.0)
    with pytest.raises(ValueError):
        D.DiscreteHomogeneousProcess([1.0], dt=0.0)    # baselines.jl:372
    with pytest.raises(ValueError):
        D.DiscreteGaussianImpulseResponse(np.full((2, 2, 3), 0.5), 4)  # rows must sum to one (impulses.jl:282)
    th = np.full((2, 2, 3), 1.0 / 3.0)
    with pytest.raises(ValueError):
        D.DiscreteStandardHawkesProcess(D.DiscreteHomogeneousProcess([1.0, 1.0], dt=0.5), D.DiscreteGaussianImpulseResponse(th, 4), nhp.DenseWeightModel(np.ones((2, 2))))


def test_discrete_params_layout():
    th = np.random.default_rng(0).dirichlet(np.ones(3), (2, 2))
    W = np.array([[0.1, 0.2], [0.3, 0.4]])
    p = D.DiscreteStandardHawkesProcess(D.DiscreteHomogeneousProcess([1.0, 2.0]), D.DiscreteGaussianImpulseResponse(th, 4), nhp.DenseWeightModel(W))
    x = p.params()  # [lambda0; vec(W .* theta)]  discrete.jl:174-182
    assert x.size == 2 + 12
    assert x[2 + 1] == pytest.approx(W[1, 0] * th[1, 0, 0])


@pytest.mark.parametrize("kind", ["ln", "exp"])
def test_gradient_vector_follows_params_order(kind):
    """gradient_vector flattens d/d(lambda0, impulse params, W) in the order of params(process) (continuous.jl:116-119), so
    x + eps * gradient_vector(...) perturbs exactly the entries the gradient belongs to."""
    from nhp_b200.continuous import gradient_vector, _hyper
    K = 3
    p = _std(K, kind)
    rng = np.random.default_rng(2)
    g = dict(lambda0=rng.normal(size=K), W=rng.normal(size=(K, K)), p1=rng.normal(size=(K, K)), p2=rng.normal(size=(K, K)) if kind == "ln" else None)
    v = gradient_vector(p, g)
    assert v.size == p.params().size
    q = _std(K, kind)
    q.params_(v)  # reading the flattened gradient back through params_ must put every block where it came from
    np.testing.assert_array_equal(q.baseline.lam, g["lambda0"])
    np.testing.assert_array_equal(q.weights.W, g["W"])
    np.testing.assert_array_equal(q.impulses.p1(), g["p1"])
    if kind == "ln":
        np.testing.assert_array_equal(q.impulses.p2(), g["p2"])
    h = _hyper(p)  # [alpha0, beta0, kappa, nu, impulse hyper-parameters] as nhp_cont_resample_params expects
    assert h.size == (8 if kind == "ln" else 6) and np.all(h[:4] == [p.baseline.alpha0, p.baseline.beta0, p.weights.kappa, p.weights.nu])


def test_log_gaussian_cox_process_host_methods():
    """Host side of the LogGaussianCoxProcess baseline (baselines.jl:187-336): interpolation support, trapezoid integral, parameter
    vector; the placeholder rates handed to nhp_cont_params_set are the time averages of the curves."""
    x = np.linspace(0.0, 10.0, 6)
    lam = np.array([[1.0, 2.0, 3.0, 2.0, 1.0, 1.0], [0.5] * 6])
    b = nhp.LogGaussianCoxProcess(x, lam, m=0.0, sigma=1.0, eta=2.0)
    assert b.ndims() == 2 and b.params().size == 12
    assert b.intensity(0, 1.0) == pytest.approx(1.5)                       # interpolation.jl:31
    assert b.intensity(0, 10.0) == pytest.approx(1.0)                      # the last grid point belongs to the support
    with pytest.raises(ValueError):
        b.intensity(0, 10.5)                                               # DomainError interpolation.jl:29
    np.testing.assert_allclose(b.integrated_intensity(), [2.0 * (0.5 * 1 + 2 + 3 + 2 + 1 + 0.5 * 1), 5.0])
    np.testing.assert_allclose(b.lam, b.integrated_intensity() / 10.0)
    assert np.isfinite(b.logprior())
    with pytest.raises(ValueError):
        nhp.LogGaussianCoxProcess(x, lam[:, :-1])
