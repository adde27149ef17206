"""GPU parity tests of the continuous hot path: libnhp (through the C ABI / nhp_b200 mirror) against
the CPU oracle on identical seeded inputs.  Tolerances (north_star): log-likelihood and intensities
<= 1e-10 relative (FP64); parent indices bit-exact given the same uniforms; counts bit-exact given the
same parents; float statistics <= 1e-12 relative (atomic summation order differs)."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth

pytestmark = pytest.mark.gpu
LL_RTOL = 1e-10
STAT_RTOL = 1e-12


def make_ln(K, seed, density=None, wmax=None, dtmax=1.0):
    lam0, W, mu, tau, A = synth.ln_params(K, seed, wmax=wmax, density=density)
    base, imp, wts = nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, dtmax), nhp.DenseWeightModel(W)
    if A is None:
        proc = nhp.ContinuousStandardHawkesProcess(base, imp, wts)
    else:
        proc = nhp.ContinuousNetworkHawkesProcess(base, imp, wts, A, nhp.BernoulliNetworkModel(density, K))
    return proc, orc.Cont(1, lam0, W, mu, tau, A=A, dtmax=dtmax)


def make_exp(K, seed, density=None, wmax=None, dtmax=np.inf):
    lam0, W, theta, A = synth.exp_params(K, seed, wmax=wmax, density=density)
    base, imp, wts = nhp.HomogeneousProcess(lam0), nhp.ExponentialImpulseResponse(theta, dtmax=dtmax), nhp.DenseWeightModel(W)
    if A is None:
        proc = nhp.ContinuousStandardHawkesProcess(base, imp, wts)
    else:
        proc = nhp.ContinuousNetworkHawkesProcess(base, imp, wts, A, nhp.BernoulliNetworkModel(density, K))
    return proc, orc.Cont(0, lam0, W, theta, A=A, dtmax=dtmax)


# ---------------------------------------------------------------------------- known answers
def test_kat_a_exponential(kat):
    k = kat["A"]
    proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(k["lambda0"]), nhp.ExponentialImpulseResponse(np.array(k["theta"])),
                                               nhp.DenseWeightModel(np.array(k["W"])))
    data = (k["events"], k["nodes"], k["duration"])
    np.testing.assert_allclose(nhp.event_intensity(proc, data), k["intensities"], rtol=1e-13)
    assert nhp.loglikelihood(proc, data, recursive=True) == pytest.approx(k["ll"], rel=1e-13)
    assert nhp.loglikelihood(proc, data, recursive=False) == pytest.approx(k["ll"], rel=1e-13)


def test_kat_b_logitnormal(kat):
    k = kat["B"]
    mu, tau = np.full((2, 2), k["mu"]), np.full((2, 2), k["tau"])
    base, wts = nhp.HomogeneousProcess(k["lambda0"]), nhp.DenseWeightModel(np.array(k["W"]))
    data = (k["events"], k["nodes"], k["duration"])
    std = nhp.ContinuousStandardHawkesProcess(base, nhp.LogitNormalImpulseResponse(mu, tau, k["dtmax"]), wts)
    np.testing.assert_allclose(nhp.event_intensity(std, data), k["intensities"], rtol=1e-13)
    assert nhp.loglikelihood(std, data) == pytest.approx(k["ll"], rel=1e-13)
    net = nhp.ContinuousNetworkHawkesProcess(base, nhp.LogitNormalImpulseResponse(mu, tau, k["dtmax"]), wts, np.array(k["A"]),
                                             nhp.BernoulliNetworkModel(0.5, 2))
    np.testing.assert_allclose(nhp.event_intensity(net, data), k["intensities_network"], rtol=1e-13)
    assert nhp.loglikelihood(net, data) == pytest.approx(k["ll_network"], rel=1e-13)


def test_kat_c_d_parents_and_stats(kat):
    kb, kc, kd = kat["B"], kat["C"], kat["D"]
    proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(kb["lambda0"]),
                                               nhp.LogitNormalImpulseResponse(np.ones((2, 2)), np.ones((2, 2)), 1.0), nhp.DenseWeightModel(np.array(kb["W"])))
    data = (kb["events"], kb["nodes"], kb["duration"])
    for d in kc["draws"]:
        u = np.full(6, 0.999999)
        u[3] = d["u"]
        par, pn = nhp.resample_parents(proc, data, u=u)
        assert par[3] == d["parent"] and par[0] == 0 and pn[0] == 0
    st = nhp.sufficient_statistics(proc, data, parents=np.array(kd["parents"]))
    np.testing.assert_array_equal(st["M0"], kd["M0"])
    np.testing.assert_array_equal(st["Mn"], kd["Mn"])
    np.testing.assert_array_equal(st["Mnm"], kd["Mnm"])
    np.testing.assert_allclose(st["S1"], kd["Xsum"], rtol=1e-13)
    np.testing.assert_allclose(st["S2"], kd["V"], rtol=1e-12, atol=1e-300)


# ---------------------------------------------------------------------------- loglik / intensity parity
@pytest.mark.parametrize("K,n,rate,density", [(2, 2000, 2.5, None), (50, 50000, 100.0, None), (50, 50000, 100.0, 0.1),
                                              (300, 40000, 64.0, 0.05), (7, 5000, 800.0, 0.5)])
@pytest.mark.parametrize("G", [None, 1, 4, 32])
def test_logitnormal_loglik_parity(K, n, rate, density, G, monkeypatch):
    if G is not None:
        monkeypatch.setenv("NHP_G", str(G))
    t, nodes, T = synth.poisson_stream(n, K, rate, 10 + K)
    proc, om = make_ln(K, 20 + K, density=density)
    data = (t, nodes, T)
    ll = nhp.loglikelihood(proc, data)
    ref = om.loglik(t, nodes, T)
    assert ll == pytest.approx(ref, rel=LL_RTOL)
    if G is None:
        np.testing.assert_allclose(nhp.event_intensity(proc, data), om.event_intensity(t, nodes), rtol=LL_RTOL)


@pytest.mark.parametrize("K,n,rate,density,dtmax", [(2, 2500, 2.5, None, np.inf), (20, 30000, 30.0, None, np.inf), (20, 30000, 30.0, 0.3, np.inf),
                                                    (20, 30000, 30.0, None, 2.0)])
@pytest.mark.parametrize("recursive", [True, False])
def test_exponential_loglik_parity(K, n, rate, density, dtmax, recursive):
    t, nodes, T = synth.poisson_stream(n, K, rate, 30 + K)
    proc, om = make_exp(K, 40 + K, density=density, wmax=0.5 / K, dtmax=dtmax)
    ll = nhp.loglikelihood(proc, (t, nodes, T), recursive=recursive)
    ref = om.loglik(t, nodes, T, recursive=recursive)
    assert ll == pytest.approx(ref, rel=LL_RTOL)
    if not recursive:
        np.testing.assert_allclose(nhp.event_intensity(proc, (t, nodes, T)), om.event_intensity(t, nodes), rtol=LL_RTOL)


def test_loglik_on_true_hawkes_sample():
    """Data drawn from the model itself (reference cluster simulator, continuous.jl:16-37)."""
    proc, om = make_ln(4, 5, wmax=0.2)
    t, nodes, T = nhp.rand(proc, 400.0, np.random.default_rng(1))
    assert t.size > 1000
    assert nhp.loglikelihood(proc, (t, nodes, T)) == pytest.approx(om.loglik(t, nodes, T), rel=LL_RTOL)


def test_edge_cases():
    proc, om = make_ln(3, 1, wmax=0.3)
    # empty data: ll = -sum(lambda0) T
    assert nhp.loglikelihood(proc, (np.zeros(0), np.zeros(0, np.int64), 5.0)) == pytest.approx(-3 * 5.0, rel=1e-15)
    # single event, equal times (dt = 0 contributes 0 for LogitNormal, theta for Exponential: quirk Q9), event at t = 0 (quirk Q6)
    t = np.array([0.0, 0.0, 0.5, 0.5, 0.5, 1.2, 1.49999, 1.5])
    nodes = np.array([1, 2, 3, 1, 1, 2, 3, 3], dtype=np.int64)
    assert nhp.loglikelihood(proc, (t, nodes, 2.0)) == pytest.approx(om.loglik(t, nodes, 2.0), rel=1e-12)
    pe, oe = make_exp(3, 2, wmax=0.3)
    for rec in (True, False):
        assert nhp.loglikelihood(pe, (t, nodes, 2.0), recursive=rec) == pytest.approx(oe.loglik(t, nodes, 2.0, recursive=rec), rel=1e-12)
    # window longer than the shared-memory staging capacity -> global-memory fallback path
    tt, nn, T = synth.poisson_stream(30000, 3, 2000.0, 9)
    pl, ol = make_ln(3, 3, wmax=0.0005, dtmax=10.0)
    assert nhp.loglikelihood(pl, (tt, nn, T)) == pytest.approx(ol.loglik(tt, nn, T), rel=LL_RTOL)


def test_error_behaviour(ctx):
    proc, _ = make_ln(3, 1)
    with pytest.raises(nhp.NHPError):  # node outside 1..K
        nhp.loglikelihood(proc, (np.array([0.1, 0.2]), np.array([1, 4]), 1.0))
    with pytest.raises(nhp.NHPError):  # unsorted times
        nhp.loglikelihood(proc, (np.array([0.3, 0.2]), np.array([1, 2]), 1.0))
    with pytest.raises(ValueError):  # DomainError baselines.jl:32
        nhp.HomogeneousProcess([-1.0])
    bad = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(3)), nhp.LogitNormalImpulseResponse(np.ones((2, 2)), np.ones((2, 2)), 1.0),
                                              nhp.DenseWeightModel(np.ones((3, 3))))
    with pytest.raises(ValueError):
        nhp.loglikelihood(bad, (np.array([0.1]), np.array([1]), 1.0))


# ---------------------------------------------------------------------------- parents + statistics
@pytest.mark.parametrize("kind", ["ln", "exp"])
@pytest.mark.parametrize("K,n,rate,density", [(2, 3000, 4.0, None), (50, 40000, 100.0, None), (40, 30000, 64.0, 0.1)])
@pytest.mark.parametrize("G", [None, 1, 8, 32])
def test_parents_bit_exact_given_uniforms(kind, K, n, rate, density, G, monkeypatch):
    if G is not None:
        monkeypatch.setenv("NHP_G", str(G))
    t, nodes, T = synth.poisson_stream(n, K, rate, 50 + K)
    if kind == "ln":
        proc, om = make_ln(K, 60 + K, density=density)
    else:
        proc, om = make_exp(K, 60 + K, density=density, wmax=0.5 / K, dtmax=1.5)
    u = np.random.default_rng(7).random(n)
    d = proc.upload((t, nodes, T))
    par, pn = nhp.resample_parents(proc, d, u=u, with_loglik=True)
    opar, opn = om.resample_parents(t, nodes, u)
    assert np.count_nonzero(par != opar) == 0
    np.testing.assert_array_equal(pn, opn)
    # statistics of that assignment: counts bit-exact, float sums to 1e-12
    assert nhp.sweep_loglikelihood(proc, d) == pytest.approx(om.loglik(t, nodes, T, recursive=False), rel=1e-10)  # fused with the sweep
    st = nhp.sufficient_statistics(proc, d)
    ost = orc.suffstats(1 if kind == "ln" else 0, t, nodes, opar, opn, K, proc.impulses.dtmax)
    for key in ("M0", "Mn", "Mnm"):
        np.testing.assert_array_equal(st[key], ost[key])
    np.testing.assert_allclose(st["S1"], ost["S1"], rtol=STAT_RTOL, atol=1e-12)
    np.testing.assert_allclose(st["S2"], ost["S2"], rtol=1e-10, atol=1e-12)
    # re-importing the same parents reproduces the same statistics
    st2 = nhp.sufficient_statistics(proc, d, parents=opar)
    np.testing.assert_array_equal(st2["Mnm"], ost["Mnm"])
    np.testing.assert_allclose(st2["S1"], ost["S1"], rtol=STAT_RTOL, atol=1e-12)


def test_parents_philox_matches_host_philox():
    """Device Philox4x32-10 uniforms == the numpy restatement, so the oracle fed with them gives identical parents."""
    K, n = 10, 20000
    t, nodes, T = synth.poisson_stream(n, K, 40.0, 3)
    proc, om = make_ln(K, 4, wmax=0.05)
    par, _ = nhp.resample_parents(proc, (t, nodes, T), seed=1234, counter=9)
    u = synth.philox_uniform(1234, np.arange(n, dtype=np.uint64), 9)
    opar, _ = om.resample_parents(t, nodes, u)
    assert np.count_nonzero(par != opar) == 0
    par2, _ = nhp.resample_parents(proc, (t, nodes, T), seed=1234, counter=10)
    assert np.count_nonzero(par2 != par) > 0


def test_parent_distribution_matches_weights():
    """Size-independent property: over many sweeps the empirical parent frequencies of one event follow lambda_s / sum."""
    kb = dict(events=[0.1, 0.4, 0.9, 1.3], nodes=[1, 2, 1, 2])
    proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess([1.0, 2.0]), nhp.LogitNormalImpulseResponse(np.ones((2, 2)), np.ones((2, 2)), 1.0),
                                               nhp.DenseWeightModel(np.array([[0.1, 0.2], [0.2, 0.1]]) * 5))
    d = proc.upload((kb["events"], kb["nodes"], 2.0))
    counts = np.zeros(4)
    nrep = 4000
    for s in range(nrep):
        par, _ = nhp.resample_parents(proc, d, seed=99, counter=s)
        counts[par[3]] += 1
    om = orc.Cont(1, [1.0, 2.0], np.array([[0.1, 0.2], [0.2, 0.1]]) * 5, np.ones((2, 2)), np.ones((2, 2)), dtmax=1.0)
    w = np.array([2.0, 0.0, om.event_intensity([0.4, 1.3], [2, 2])[1] - 2.0, om.event_intensity([0.9, 1.3], [1, 2])[1] - 2.0])
    p = w / w.sum()
    assert np.all(np.abs(counts / nrep - p) < 4 * np.sqrt(p * (1 - p) / nrep) + 1e-9)


def test_halo_shards_reproduce_unsharded():
    """Time shards with a dtmax halo: shares add up to the unsharded log-likelihood; parents are global indices."""
    K, n = 12, 30000
    t, nodes, T = synth.poisson_stream(n, K, 60.0, 5)
    proc, om = make_ln(K, 6, wmax=0.05)
    ctx = proc._ctx()
    ref = nhp.loglikelihood(proc, (t, nodes, T))
    u = np.random.default_rng(3).random(n)
    par_ref, _ = nhp.resample_parents(proc, (t, nodes, T), u=u)
    st_ref = nhp.sufficient_statistics(proc, (t, nodes, T), parents=par_ref)
    total, pars, Mnm = 0.0, [], 0
    bounds = [0, 7000, 7001, 19000, n]
    for r in range(4):
        a, b = bounds[r], bounds[r + 1]
        lo = int(np.searchsorted(t, t[a] - 1.0, side="right")) if a > 0 else 0
        d = nhp.ContinuousData(ctx, t[lo:b], nodes[lo:b], T, K, n_halo=a - lo, index_base=lo, flags=1 if r == 0 else 0)
        total += nhp.loglikelihood(proc, d)
        p, _ = nhp.resample_parents(proc, d, u=u[a:b])
        pars.append(p)
        Mnm = Mnm + nhp.sufficient_statistics(proc, d)["Mnm"]
    assert total == pytest.approx(ref, rel=1e-12)
    np.testing.assert_array_equal(np.concatenate(pars), par_ref)
    np.testing.assert_array_equal(Mnm, st_ref["Mnm"])


@pytest.mark.parametrize("K,n,rate,density", [(2, 3000, 30.0, None), (12, 20000, 200.0, None), (12, 20000, 200.0, 0.4), (40, 30000, 500.0, None)])
def test_recursive_exponential_chunked_scan_matches_oracle_recursion(K, n, rate, density, monkeypatch):
    """recursive_loglikelihood (continuous.jl:241-276 / 407-442): the chunked-scan formulation (cont_exp_scan.cu, forced here)
    and the cut-off-horizon window sweep both reproduce the oracle's sequential recursion, incl. events at t == 0 (quirk Q6)."""
    proc, om = make_exp(K, 5, density=density, wmax=0.6 / K)
    t, nodes, T = synth.poisson_stream(n, K, rate, 12)
    t[:3] = 0.0  # leading events at exactly 0.0 never act as parents
    ref = om.loglik(t, nodes, T, recursive=True)
    d = proc.upload((t, nodes, T))
    for mode in ("1", "0"):
        monkeypatch.setenv("NHP_EXP_SCAN", mode)
        assert nhp.loglikelihood(proc, d, recursive=True) == pytest.approx(ref, rel=LL_RTOL), mode
    monkeypatch.delenv("NHP_EXP_SCAN")
    assert nhp.loglikelihood(proc, d, recursive=True) == pytest.approx(ref, rel=LL_RTOL)
