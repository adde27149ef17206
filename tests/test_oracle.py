"""CPU tests: the oracle against the formula-derived KAT vectors (tests/golden/kat.json, made by
tests/golden/make_kat.py with SciPy) and against the fixtures of the reference's own tests
(/root/reference/test/baselines.jl:10-23, 77-78)."""
import numpy as np
import pytest

import oracle_ffi as orc

RTOL = 1e-13


def test_kat_a_exponential(kat):
    k = kat["A"]
    m = orc.Cont(0, k["lambda0"], np.array(k["W"]), np.array(k["theta"]))
    lam = m.event_intensity(k["events"], k["nodes"])
    np.testing.assert_allclose(lam, k["intensities"], rtol=RTOL)
    assert m.loglik(k["events"], k["nodes"], k["duration"], recursive=False) == pytest.approx(k["ll"], rel=RTOL)
    # recursive and windowed forms agree (continuous.jl:241-276 vs 210-239)
    assert m.loglik(k["events"], k["nodes"], k["duration"], recursive=True) == pytest.approx(k["ll"], rel=RTOL)


def test_kat_b_logitnormal(kat):
    k = kat["B"]
    W = np.array(k["W"])
    mu, tau = np.full((2, 2), k["mu"]), np.full((2, 2), k["tau"])
    m = orc.Cont(1, k["lambda0"], W, mu, tau, dtmax=k["dtmax"])
    np.testing.assert_allclose(m.event_intensity(k["events"], k["nodes"]), k["intensities"], rtol=RTOL)
    assert m.loglik(k["events"], k["nodes"], k["duration"]) == pytest.approx(k["ll"], rel=RTOL)
    mn = orc.Cont(1, k["lambda0"], W, mu, tau, A=np.array(k["A"]), dtmax=k["dtmax"])
    np.testing.assert_allclose(mn.event_intensity(k["events"], k["nodes"]), k["intensities_network"], rtol=RTOL)
    assert mn.loglik(k["events"], k["nodes"], k["duration"]) == pytest.approx(k["ll_network"], rel=RTOL)


def test_kat_c_parent_draws(kat):
    kb, kc = kat["B"], kat["C"]
    m = orc.Cont(1, kb["lambda0"], np.array(kb["W"]), np.full((2, 2), 1.0), np.full((2, 2), 1.0), dtmax=1.0)
    for d in kc["draws"]:
        u = np.full(6, 0.999999)
        u[3] = d["u"]
        par, pn = m.resample_parents(kb["events"], kb["nodes"], u)
        assert par[3] == d["parent"]
        assert pn[3] == (kb["nodes"][d["parent"] - 1] if d["parent"] > 0 else 0)
        assert par[0] == 0 and pn[0] == 0  # index == 1 -> (0, 0)  parents.jl:26-28


def test_kat_d_sufficient_statistics(kat):
    kb, kd = kat["B"], kat["D"]
    nodes = np.array(kb["nodes"])
    par = np.array(kd["parents"])
    pn = np.where(par > 0, nodes[np.maximum(par, 1) - 1], 0)
    st = orc.suffstats(1, kb["events"], nodes, par, pn, 2, 1.0)
    np.testing.assert_array_equal(st["M0"], kd["M0"])
    np.testing.assert_array_equal(st["Mn"], kd["Mn"])
    np.testing.assert_array_equal(st["Mnm"], kd["Mnm"])
    np.testing.assert_allclose(st["S1"], kd["Xsum"], rtol=RTOL)
    np.testing.assert_allclose(st["S2"], kd["V"], rtol=1e-12, atol=1e-300)
    ste = orc.suffstats(0, kb["events"], nodes, par, pn, 2, np.inf)
    np.testing.assert_allclose(ste["duration_mean"], kd["duration_mean"], rtol=RTOL)


def test_kat_e_discrete(kat):
    k = kat["E"]
    phi = orc.disc_basis(k["L"], k["B"], k["dt"])
    np.testing.assert_allclose(phi, k["phi"], rtol=RTOL)
    data = np.array(k["data"], dtype=np.int64)
    conv = orc.disc_convolve(data, phi)
    np.testing.assert_allclose(conv, k["conv"], rtol=RTOL, atol=1e-300)
    m = orc.Disc(k["lambda0"], np.array(k["W"]), np.full((2, 2, 3), 1.0 / 3.0), dt=k["dt"])
    np.testing.assert_allclose(m.intensity(conv), k["lam"], rtol=RTOL)
    assert m.loglik(data, conv) == pytest.approx(k["ll"], rel=1e-12)


def test_reference_fixture_node_counts(kat):
    # test/baselines.jl:10-23
    import ctypes
    for case in kat["ref_tests"]["node_counts"]:
        nodes = np.array(case["nodes"], dtype=np.int64)
        pn = np.array(case["parentnodes"], dtype=np.int64)
        out = np.empty(case["K"])
        orc.lib().orc_baseline_counts(orc._p(nodes), orc._p(pn), ctypes.c_int64(nodes.size), ctypes.c_int64(case["K"]), orc._p(out))
        np.testing.assert_array_equal(out, case["expect"])


def test_reference_fixture_discrete_counts(kat):
    # test/baselines.jl:77-78: sufficient_statistics(process, data) == ([3, 2], 10); nu_sum carries the same row sums
    c = kat["ref_tests"]["disc_suffstats"]
    data = np.array(c["data"], dtype=np.int64)
    conv = orc.disc_convolve(data, orc.disc_basis(4, 3))
    st = orc.disc_vb_stats(data, conv, np.ones(2), np.ones((2, 2, 3)))
    np.testing.assert_array_equal(st["nu_sum"][:, 0], c["expect_Mn"])
    assert data.shape[1] == c["expect_T"]


def test_recursive_equals_windowed_full_history():
    import synth
    t, nodes, T = synth.poisson_stream(400, 3, 5.0, 1)
    lam0, W, theta, _ = synth.exp_params(3, 2, wmax=0.3)
    m = orc.Cont(0, lam0, W, theta)
    assert m.loglik(t, nodes, T, recursive=True) == pytest.approx(m.loglik(t, nodes, T, recursive=False), rel=1e-12)


def test_threads_match_serial():
    import synth
    t, nodes, T = synth.poisson_stream(3000, 5, 50.0, 3)
    lam0, W, mu, tau, A = synth.ln_params(5, 4, wmax=0.2, density=0.5)
    m = orc.Cont(1, lam0, W, mu, tau, A=A, dtmax=1.0)
    u = np.random.default_rng(0).random(t.size)
    orc.set_threads(1)
    ll1 = m.loglik(t, nodes, T)
    p1 = m.resample_parents(t, nodes, u)
    orc.set_threads(4)
    ll4 = m.loglik(t, nodes, T)
    p4 = m.resample_parents(t, nodes, u)
    orc.set_threads(1)
    assert ll4 == pytest.approx(ll1, rel=1e-12)
    np.testing.assert_array_equal(p1[0], p4[0])


def test_philox_reference_vector():
    # Random123 known-answer test for philox4x32-10: counter = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    import synth
    u = synth.philox_uniform(0, np.array([0], dtype=np.uint64), 0)
    x = ((0x6627E8D5 << 32) | 0xE169C58D) >> 11
    assert u[0] == x * 2.0 ** -53


# ---------------------------------------------------------------------------- extension: analytic gradient
def _fd_grad(make, x0, f, h=1e-6):
    g = np.zeros_like(x0)
    for k in range(x0.size):
        xp, xm = x0.copy(), x0.copy()
        xp[k] += h
        xm[k] -= h
        g[k] = (f(make(xp)) - f(make(xm))) / (2 * h)
    return g


@pytest.mark.parametrize("kind,network,recursive", [(0, False, True), (0, False, False), (0, True, True), (0, True, False), (1, False, False), (1, True, False)])
def test_gradient_oracle_matches_finite_differences(kind, network, recursive):
    """The gradient has no reference counterpart: the oracle's analytic gradient is pinned by central differences
    of the oracle's own log-likelihood (incl. the compensator quirks Q3 of the recursive network path)."""
    rng = np.random.default_rng(5 + kind + 2 * network)
    K, n, T = 3, 70, 20.0
    dtmax = 1.5 if kind == 1 else (np.inf if recursive else 2.0)
    ev = np.sort(rng.uniform(0.01, T, n))
    nd = rng.integers(1, K + 1, n)
    A = (rng.random((K, K)) < 0.6).astype(np.float64) if network else None
    lam0, W = rng.uniform(0.5, 1.5, K), rng.uniform(0.05, 0.4, (K, K))
    p1 = rng.uniform(0.5, 2.0, (K, K)) if kind == 0 else rng.uniform(-1, 1, (K, K))
    p2 = None if kind == 0 else rng.uniform(0.5, 2.0, (K, K))
    x0 = np.concatenate([lam0, W.ravel(), p1.ravel()] + ([] if p2 is None else [p2.ravel()]))

    def make(x):
        q2 = None if p2 is None else x[K + 2 * K * K:].reshape(K, K)
        return orc.Cont(kind, x[:K], x[K:K + K * K].reshape(K, K), x[K + K * K:K + 2 * K * K].reshape(K, K), q2, A=A, dtmax=dtmax)

    ll, g0, gW, g1, g2 = make(x0).loglik_grad(ev, nd, T, recursive=recursive)
    assert ll == make(x0).loglik(ev, nd, T, recursive=recursive)
    ana = np.concatenate([g0, gW.ravel(), g1.ravel()] + ([] if p2 is None else [g2.ravel()]))
    num = _fd_grad(make, x0, lambda m: m.loglik(ev, nd, T, recursive=recursive))
    np.testing.assert_allclose(ana, num, rtol=2e-6, atol=2e-6)


def test_gradient_oracle_closed_form_two_events():
    """K = 1, two events at t = 1, 2, Exponential impulse: lambda_2 = lambda0 + W theta exp(-theta);
    ll = log lambda0 + log lambda_2 - lambda0 T - 2 W (the compensator counts W once per event, continuous.jl:219-221)."""
    lam0, W, th, T = 0.7, 0.4, 1.3, 5.0
    om = orc.Cont(0, np.array([lam0]), np.array([[W]]), np.array([[th]]))
    ev, nd = np.array([1.0, 2.0]), np.array([1, 1])
    l2 = lam0 + W * th * np.exp(-th)
    for rec in (True, False):
        ll, g0, gW, g1, _ = om.loglik_grad(ev, nd, T, recursive=rec)
        assert ll == pytest.approx(np.log(lam0) + np.log(l2) - lam0 * T - 2 * W, rel=1e-14)
        assert g0[0] == pytest.approx(1 / lam0 + 1 / l2 - T, rel=1e-14)
        assert gW[0, 0] == pytest.approx(th * np.exp(-th) / l2 - 2.0, rel=1e-14)
        assert g1[0, 0] == pytest.approx(W * np.exp(-th) * (1 - th) / l2, rel=1e-14)


# ---------------------------------------------------------------------------- round-2 vectors (KAT-F..K): rows a5, a9, Q3, a12, a13, a14
def _kat_b_models(kat):
    k = kat["B"]
    W, mu, tau = np.array(k["W"]), np.full((2, 2), k["mu"]), np.full((2, 2), k["tau"])
    return k, orc.Cont(1, k["lambda0"], W, mu, tau, dtmax=k["dtmax"]), orc.Cont(1, k["lambda0"], W, mu, tau, A=np.array(k["A"]), dtmax=k["dtmax"])


def test_kat_f_intensity_at_query_times_strict_window(kat):
    k, std, net = _kat_b_models(kat)
    tq = np.array(kat["F"]["times"])
    np.testing.assert_allclose(std.intensity(k["events"], k["nodes"], tq), kat["F"]["standard"], rtol=RTOL)
    np.testing.assert_allclose(net.intensity(k["events"], k["nodes"], tq), kat["F"]["network"], rtol=RTOL)


def test_kat_g_adjacency_sweep_given_uniforms(kat):
    k, _, net = _kat_b_models(kat)
    for case in kat["G"]["cases"]:
        A = net.resample_adjacency(np.array(kat["G"]["A0"]), np.full((2, 2), case["rho"]), k["events"], k["nodes"], k["duration"], np.array(case["u"]))
        np.testing.assert_array_equal(A, case["A"])


def test_kat_h_recursive_network_loglik_quirk_q3(kat):
    k = kat["H"]
    m = orc.Cont(0, k["lambda0"], np.array(k["W"]), np.array(k["theta"]), A=np.array(k["A"]))
    np.testing.assert_allclose(m.event_intensity(k["events"], k["nodes"]), k["intensities"], rtol=RTOL)
    assert m.loglik(k["events"], k["nodes"], k["duration"], recursive=True) == pytest.approx(k["ll_recursive"], rel=RTOL)
    assert m.loglik(k["events"], k["nodes"], k["duration"], recursive=False) == pytest.approx(k["ll_windowed"], rel=RTOL)
    assert k["ll_recursive"] != k["ll_windowed"]  # the compensator of the recursive form ignores A


def _kat_i_model(kat, A=None):
    k = kat["I"]
    return k, np.array(k["data"], dtype=np.int64), np.array(k["conv"]), orc.Disc(k["lambda0"], np.array(k["W"]), np.array(k["theta"]), dt=1.0, A=A)


def test_kat_i_discrete_gibbs_counts_given_uniforms(kat):
    k, data, conv, m = _kat_i_model(kat)
    np.testing.assert_allclose(orc.disc_convolve(data, orc.disc_basis(4, 3, 1.0)), conv, rtol=RTOL, atol=1e-300)
    np.testing.assert_array_equal(m.gibbs_counts(data, conv, np.array(k["u"])), k["counts"])


def test_kat_j_vb_statistics(kat):
    k, data, conv, _ = _kat_i_model(kat)
    j = kat["J"]
    st = orc.disc_vb_stats(data, conv, np.array(j["e0"]), np.array(j["E"]))
    np.testing.assert_allclose(st["alpha_sum"], j["alpha_sum"], rtol=RTOL)
    np.testing.assert_allclose(st["gamma_sum"], j["gamma_sum"], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(st["kappa_sum"], j["kappa_sum"], rtol=RTOL)
    np.testing.assert_array_equal(st["nu_sum"], j["nu_sum"])


def test_kat_k_discrete_adjacency_given_uniforms(kat):
    A0 = np.array(kat["K"]["A0"])
    k, data, conv, m = _kat_i_model(kat, A=A0)
    for case in kat["K"]["cases"]:
        A = m.resample_adjacency(A0, np.full((2, 2), case["rho"]), data, conv, np.array(case["u"]))
        np.testing.assert_array_equal(A, case["A"])
