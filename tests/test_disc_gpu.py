"""GPU parity tests of the discrete-time path against the oracle (convolve, intensity GEMM, Poisson
log-likelihood, Gibbs parent counts, VB statistics, adjacency Gibbs)."""
import numpy as np
import pytest

import nhp_b200 as nhp
from nhp_b200 import discrete as D
import oracle_ffi as orc

pytestmark = pytest.mark.gpu


def make(N, T, B, L, seed, network=False, rate=0.05):
    rng = np.random.default_rng(seed)
    lam0 = rng.uniform(0.5, 1.5, N) * rate
    W = rng.uniform(0.0, 0.6 / N, (N, N))
    theta = rng.dirichlet(np.ones(B), (N, N))
    A = (rng.random((N, N)) < 0.5).astype(np.float64) if network else None
    data = rng.poisson(rate, (N, T)).astype(np.int64)
    base, imp, wts = D.DiscreteHomogeneousProcess(lam0), D.DiscreteGaussianImpulseResponse(theta, L), nhp.DenseWeightModel(W)
    proc = D.DiscreteNetworkHawkesProcess(base, imp, wts, A, nhp.BernoulliNetworkModel(0.4, N)) if network else D.DiscreteStandardHawkesProcess(base, imp, wts)
    return proc, orc.Disc(lam0, W, theta, dt=1.0, A=A), data


def test_kat_e(kat):
    k = kat["E"]
    theta = np.full((2, 2, 3), 1.0 / 3.0)
    proc = D.DiscreteStandardHawkesProcess(D.DiscreteHomogeneousProcess(k["lambda0"]), D.DiscreteGaussianImpulseResponse(theta, k["L"]), nhp.DenseWeightModel(np.array(k["W"])))
    np.testing.assert_allclose(proc.impulses.basis(), k["phi"], rtol=1e-14)
    data = np.array(k["data"], dtype=np.int64)
    d = proc.upload(data)
    np.testing.assert_allclose(D.convolve(proc, d), k["conv"], rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(D.intensity(proc, d), k["lam"], rtol=1e-13)
    assert D.loglikelihood(proc, d) == pytest.approx(k["ll"], rel=1e-12)


@pytest.mark.parametrize("N,T,B,L,network", [(3, 500, 3, 4, False), (20, 3000, 4, 8, True), (70, 1500, 6, 12, False), (5, 200, 1 + 1, 2, True),
                                             (200, 3000, 6, 12, True)])  # last: BASELINE config 3's shape (N = 200, B = 6, L = 12: the NT = 5 DMMA tile)
def test_convolve_intensity_loglik(N, T, B, L, network):
    proc, om, data = make(N, T, B, L, 3 + N, network)
    d = proc.upload(data)
    conv = D.convolve(proc, d)
    oconv = orc.disc_convolve(data, orc.disc_basis(L, B))
    np.testing.assert_allclose(conv, oconv, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(D.intensity(proc, d), om.intensity(oconv), rtol=1e-10)
    assert D.loglikelihood(proc, d) == pytest.approx(om.loglik(data, oconv), rel=1e-10)


@pytest.mark.parametrize("N,T,B,L", [(33, 700, 5, 40), (64, 300, 7, 32), (9, 1000, 8, 33), (1, 130, 1, 1), (300, 260, 3, 5), (200, 129, 6, 64)])
def test_convolve_row_kernel_shapes(N, T, B, L):
    """Shapes around the limits of the row-segment convolve kernel (k_convolve_rows): 64-bit lag masks (L > 32), partial warps, several
    row groups per CTA (small N), segment boundaries (T just above a multiple of 128), N > 256 (thread-per-(t, p) kernel); the basis is
    the reference's (discrete.jl:171-180 through the oracle); max(0, .) and the lags-ascending sum give the oracle's digits."""
    proc, om, data = make(N, T, B, L, 40 + N, False, rate=0.3)
    d = proc.upload(data)
    conv = D.convolve(proc, d)
    oconv = orc.disc_convolve(data, orc.disc_basis(L, B))
    np.testing.assert_allclose(conv, oconv, rtol=1e-13, atol=1e-300)
    # the fused column sums feed the compensator of the log-likelihood
    assert D.loglikelihood(proc, d) == pytest.approx(om.loglik(data, oconv), rel=1e-10)


@pytest.mark.parametrize("warp", ["1", "0"])
@pytest.mark.parametrize("N,T,B,L,network", [(3, 400, 3, 4, False), (12, 1500, 4, 6, True), (70, 800, 6, 12, False), (200, 1500, 6, 12, True)])
def test_gibbs_counts_match_given_uniforms(N, T, B, L, network, warp, monkeypatch):
    monkeypatch.setenv("NHP_DISC_WARP", warp)
    proc, om, data = make(N, T, B, L, 11 + N, network, rate=0.2)
    d = proc.upload(data)
    D.convolve(proc, d, export=False)
    oconv = orc.disc_convolve(data, orc.disc_basis(L, B))
    u = np.random.default_rng(5).random(int(data.sum()))
    counts = D.resample_parents(proc, d, u=u)
    ref = om.gibbs_counts(data, oconv, u)
    assert counts.sum() == data.sum()
    np.testing.assert_array_equal(counts.sum(axis=1), data.sum(axis=1))  # every event gets exactly one parent
    assert np.count_nonzero(counts != ref) == 0


@pytest.mark.parametrize("tb", ["", "37", "100000"])
def test_gibbs_counts_sparse_network_and_time_blocks(tb, monkeypatch):
    """Sparse-network path (compact non-zero parent lists) incl. a child without any parent, for several time-block sizes
    of the (time block, child) decomposition; bit-exact against the oracle given the uniforms."""
    if tb:
        monkeypatch.setenv("NHP_DISC_TB", tb)
    N, T, B, L = 24, 2500, 4, 6
    rng = np.random.default_rng(77)
    lam0 = rng.uniform(0.5, 1.5, N) * 0.1
    A = (rng.random((N, N)) < 0.12).astype(np.float64)
    A[:, 5] = 0.0   # child 5: baseline only
    A[:, 9] = 1.0   # child 9: every parent
    W = rng.uniform(0.0, 2.0 / N, (N, N))
    theta = rng.dirichlet(np.ones(B), (N, N))
    data = rng.poisson(0.15, (N, T)).astype(np.int64)
    proc = D.DiscreteNetworkHawkesProcess(D.DiscreteHomogeneousProcess(lam0), D.DiscreteGaussianImpulseResponse(theta, L), nhp.DenseWeightModel(W), A,
                                          nhp.BernoulliNetworkModel(0.12, N))
    om = orc.Disc(lam0, W, theta, dt=1.0, A=A)
    d = proc.upload(data)
    D.convolve(proc, d, export=False)
    oconv = orc.disc_convolve(data, orc.disc_basis(L, B))
    u = np.random.default_rng(6).random(int(data.sum()))
    for sparse in ("1", "0"):
        monkeypatch.setenv("NHP_DISC_SPARSE", sparse)
        counts = D.resample_parents(proc, d, u=u)
        assert np.count_nonzero(counts != om.gibbs_counts(data, oconv, u)) == 0
    assert np.all(counts[5, 1:] == 0) and counts[5, 0] == data[5].sum()


def test_gibbs_counts_distribution():
    """Size-independent property: summed over many sweeps the counts follow the expected attribution mass."""
    proc, om, data = make(4, 600, 3, 4, 21, False, rate=0.3)
    d = proc.upload(data)
    conv = D.convolve(proc, d)
    lam = om.intensity(conv)
    N, B = 4, 3
    expect = np.zeros((N, 1 + N * B))
    bump = proc.weights.W[:, :, None] * proc.impulses.theta  # [p, c, b]
    for c in range(N):
        s = data[c].astype(float)
        expect[c, 0] = np.sum(s * proc.baseline.lam[c] / lam[:, c])
        for p in range(N):
            for b in range(B):
                expect[c, 1 + p * B + b] = np.sum(s * conv[:, p, b] * bump[p, c, b] / lam[:, c])
    tot = np.zeros_like(expect)
    nrep = 200
    for r in range(nrep):
        tot += D.resample_parents(proc, d, seed=3, counter=r)
    mean = tot / nrep
    assert np.all(np.abs(mean - expect) < 5 * np.sqrt(expect / nrep + 1e-9) + 0.02)


@pytest.mark.parametrize("warp", ["1", "0"])
@pytest.mark.parametrize("N,T,B,L", [(3, 400, 3, 4), (15, 2000, 5, 10), (70, 800, 6, 12), (200, 1200, 6, 12)])
def test_vb_statistics(N, T, B, L, warp, monkeypatch):
    monkeypatch.setenv("NHP_DISC_WARP", warp)
    proc, om, data = make(N, T, B, L, 31 + N, False, rate=0.1)
    d = proc.upload(data)
    conv = D.convolve(proc, d)
    rng = np.random.default_rng(2)
    e0, E = rng.uniform(0.5, 1.5, N), rng.uniform(0.01, 0.2, (N, N, B))
    st = D.vb_statistics(proc, d, e0, E)
    ref = orc.disc_vb_stats(data, conv, e0, E)
    for k in ("alpha_sum", "kappa_sum", "gamma_sum"):
        np.testing.assert_allclose(st[k], ref[k], rtol=1e-10, atol=1e-13)
    np.testing.assert_array_equal(st["nu_sum"], ref["nu_sum"])
    # fixture of the reference's own tests (test/baselines.jl:77-78): row sums ([3, 2], T = 10)


def test_vb_and_gibbs_steps_run():
    proc, om, data = make(4, 800, 3, 4, 41, False, rate=0.2)
    d = proc.upload(data)
    trace = D.vb_(proc, d, max_steps=3)
    assert len(trace) == 3 and np.all(np.isfinite(trace[-1]))
    res = D.mcmc_(proc, d, nsteps=3, seed=1)
    assert len(res.samples) == 3 and np.all(np.isfinite(res.samples[-1]))
    np.testing.assert_allclose(proc.impulses.theta.sum(axis=2), 1.0, rtol=1e-12)


@pytest.mark.parametrize("N,T,B,L", [(4, 300, 3, 4), (8, 600, 2, 5), (48, 300, 6, 12)])
def test_discrete_adjacency_matches_oracle(N, T, B, L):
    proc, om, data = make(N, T, B, L, 51 + N, True, rate=0.2)
    d = proc.upload(data)
    conv = D.convolve(proc, d)
    A0 = proc.adjacency_matrix.copy()
    mism = 0
    for rep in range(3):
        u = np.random.default_rng(60 + rep).random((N, N))
        proc.adjacency_matrix = A0.copy()
        A = D.resample_adjacency_matrix_(proc, d, u=u).copy()
        ref = om.resample_adjacency(A0, np.full((N, N), 0.4), data, conv, u)
        mism += int(np.count_nonzero(A != ref))
    assert mism == 0


def test_time_shards_with_lag_halo_reproduce_unsharded():
    """Discrete multi-GPU form: contiguous time shards, each preceded by an L-bin halo of counts; the
    log-likelihood shares, Gibbs counts and VB statistics add up to the unsharded results."""
    N, T, B, L = 6, 1200, 3, 5
    proc, om, data = make(N, T, B, L, 77, False, rate=0.15)
    ctx = proc._ctx()
    d = proc.upload(data)
    D.convolve(proc, d, export=False)
    ll_ref = D.loglikelihood(proc, d)
    rng = np.random.default_rng(4)
    e0, E = rng.uniform(0.5, 1.5, N), rng.uniform(0.01, 0.2, (N, N, B))
    vb_ref = D.vb_statistics(proc, d, e0, E)
    cnt_ref = D.resample_parents(proc, d, seed=5, counter=1).sum()
    ll, kappa, alpha, total = 0.0, 0.0, 0.0, 0.0
    bounds = [0, 400, 401, 900, T]
    for r in range(4):
        a, b = bounds[r], bounds[r + 1]
        lo = max(0, a - L)
        sh = D.DiscreteData(ctx, data[:, lo:b], t_halo=a - lo)
        D.convolve(proc, sh, export=False)
        ll += D.loglikelihood(proc, sh)
        st = D.vb_statistics(proc, sh, e0, E)
        kappa = kappa + st["kappa_sum"]
        alpha = alpha + st["alpha_sum"]
        total += D.resample_parents(proc, sh, seed=5, counter=1).sum()
    assert ll == pytest.approx(ll_ref, rel=1e-12)
    np.testing.assert_allclose(kappa, vb_ref["kappa_sum"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(alpha, vb_ref["alpha_sum"], rtol=1e-11)
    assert total == cnt_ref == data.sum()


def test_gibbs_counts_after_adjacency_resample_use_the_new_adjacency():
    """The C-ABI sequence resample_adjacency -> gibbs_counts without a parameter push in between must see the NEW links
    (the compact parent lists of the old adjacency are dropped)."""
    import ctypes
    from nhp_b200.core import _fmat, _ptr
    N, T, B, L = 10, 1500, 3, 5
    rng = np.random.default_rng(3)
    lam0 = rng.uniform(0.05, 0.15, N)
    A = (rng.random((N, N)) < 0.2).astype(np.float64)
    W, theta = rng.uniform(0.2 / N, 2.0 / N, (N, N)), rng.dirichlet(np.ones(B), (N, N))
    data = rng.poisson(0.15, (N, T)).astype(np.int64)
    proc = D.DiscreteNetworkHawkesProcess(D.DiscreteHomogeneousProcess(lam0), D.DiscreteGaussianImpulseResponse(theta, L), nhp.DenseWeightModel(W), A,
                                          nhp.BernoulliNetworkModel(0.5, N))
    d = proc.upload(data)
    oconv = orc.disc_convolve(data, orc.disc_basis(L, B))
    D.convolve(proc, d, export=False)
    A_new = D.resample_adjacency_matrix_(proc, d, seed=4).copy()
    assert not np.array_equal(A_new, A)
    ctx = proc._ctx()
    u = np.random.default_rng(8).random(int(data.sum()))
    counts = np.empty(N * (1 + N * B))
    ctx.check(ctx.lib.nhp_disc_gibbs_counts(ctx.h, d.h, 0, 0, _ptr(u), u.size, _ptr(counts)))  # no params_set in between
    ref = orc.Disc(lam0, W, theta, dt=1.0, A=A_new).gibbs_counts(data, oconv, u)
    assert np.array_equal(counts.reshape(1 + N * B, N).T, ref)


def test_device_conjugate_draws_have_the_posterior_moments():
    """nhp_disc_resample_params: lambda0 / W Gamma draws and the Dirichlet theta from the device-resident counts of one parent sweep
    (discrete.jl:361-367): moments over repeated draws with different Philox counters against the closed forms."""
    rng = np.random.default_rng(11)
    N, T, B, L = 5, 4000, 3, 6
    l0, W, th = rng.uniform(0.05, 0.15, N), rng.uniform(0.0, 0.5 / N, (N, N)), rng.dirichlet(np.ones(B), (N, N))
    data = rng.poisson(0.15, (N, T)).astype(np.int64)
    proc = D.DiscreteStandardHawkesProcess(D.DiscreteHomogeneousProcess(l0), D.DiscreteGaussianImpulseResponse(th, L), nhp.DenseWeightModel(W))
    d = proc.upload(data)
    ctx = proc._ctx()
    u = rng.random(int(data.sum()))
    C = D.resample_parents(proc, d, u=u)                     # [c, k]; the same counts stay on the device
    Mn = data.sum(axis=1).astype(np.float64)
    b, w, imp = proc.baseline, proc.weights, proc.impulses
    hy = np.array([b.alpha0, b.beta0, w.kappa, w.nu, imp.gamma])
    from nhp_b200.core import _ptr
    R = 400
    L0, WW, TH = np.empty((R, N)), np.empty((R, N, N)), np.empty((R, N, N, B))
    for r in range(R):
        lam, Wf, thf = np.empty(N), np.empty(N * N), np.empty(N * N * B)
        # the draws replace the context's parameters, the counts on the device stay those of the sweep above
        ctx.check(ctx.lib.nhp_disc_resample_params(ctx.h, d.h, 5, r, _ptr(Mn), _ptr(hy), 5, _ptr(lam), _ptr(Wf), _ptr(thf)))
        L0[r], WW[r], TH[r] = lam, Wf.reshape(N, N).T, thf.reshape(B, N, N).transpose(2, 1, 0)
    np.testing.assert_allclose(TH.sum(axis=3), 1.0, rtol=1e-12)
    a_l, r_l = b.alpha0 + C[:, 0], b.beta0 + T * proc.dt
    assert np.all(np.abs(L0.mean(0) - a_l / r_l) < 5 * np.sqrt(a_l) / r_l / np.sqrt(R) + 1e-12)
    Cpb = C[:, 1:].reshape(N, N, B)                          # [c, p, b]
    a_w, r_w = w.kappa + Cpb.sum(axis=2).T, (w.nu + Mn)[:, None]
    assert np.all(np.abs(WW.mean(0) - a_w / r_w) < 5 * np.sqrt(a_w) / r_w / np.sqrt(R) + 1e-12)
    g = imp.gamma + Cpb.transpose(1, 0, 2)                   # [p, c, b]
    g0 = g.sum(axis=2, keepdims=True)
    mean, var = g / g0, g * (g0 - g) / (g0 ** 2 * (g0 + 1.0))
    assert np.all(np.abs(TH.mean(0) - mean) < 5 * np.sqrt(var / R) + 1e-12)
    # and the whole device sweep runs as a chain
    for s_ in range(3):
        x = D.resample_on_device_(proc, d, seed=3, counter=s_)
        assert np.all(np.isfinite(x))


@pytest.mark.parametrize("N,T,B,L,network", [(4, 600, 3, 5, False), (30, 2500, 4, 8, True), (200, 1500, 6, 12, True)])
def test_analytic_gradient_of_the_discrete_loglikelihood(N, T, B, L, network):
    """nhp_disc_loglik_grad against a NumPy statement of the gradient (dense contraction over all bins) built on the oracle's convolution,
    and against central differences of the oracle's log-likelihood."""
    proc, om, data = make(N, T, B, L, 90 + N, network, rate=0.08)
    d = proc.upload(data)
    ll, g = D.loglikelihood_gradient(proc, d)
    conv = orc.disc_convolve(data, orc.disc_basis(L, B))          # [t, p, b]
    assert ll == pytest.approx(om.loglik(data, conv), rel=1e-10)
    lam0, W, th = proc.baseline.lam, proc.weights.W, proc.impulses.theta
    A = np.ones((N, N)) if proc.adjacency_matrix is None else proc.adjacency_matrix
    dt = proc.dt
    bump = (A * W)[:, :, None] * th * dt                            # [p, c, b]
    lam = lam0[None, :] * dt + np.einsum("tpb,pcb->tc", conv, bump)
    r = data.T / lam - 1.0                                          # [t, c]
    Gb = np.einsum("tpb,tc->pcb", conv, r)
    np.testing.assert_allclose(g["lambda0"], dt * r.sum(axis=0), rtol=1e-9, atol=1e-9 * T)
    scale_W, scale_T = np.max(np.abs(dt * (th * Gb).sum(axis=2))), np.max(np.abs(W[:, :, None] * dt * Gb))
    np.testing.assert_allclose(g["W"], A * dt * (th * Gb).sum(axis=2), rtol=1e-9, atol=1e-10 * scale_W)
    np.testing.assert_allclose(g["theta"], (A * W)[:, :, None] * dt * Gb, rtol=1e-9, atol=1e-10 * scale_T)
    if N <= 30:  # central differences of the oracle on a few entries
        rng = np.random.default_rng(1)
        for _ in range(4):
            p, c = rng.integers(0, N, 2)
            if A[p, c] == 0.0:
                continue
            h = 1e-3 * max(W[p, c], 1e-3)  # the oracle's ll ~ 1e4 carries ~1e-12 of rounding: a smaller step drowns in it
            Wp, Wm = W.copy(), W.copy()
            Wp[p, c] += h
            Wm[p, c] -= h
            fd = (orc.Disc(lam0, Wp, th, dt=dt, A=proc.adjacency_matrix).loglik(data, conv) - orc.Disc(lam0, Wm, th, dt=dt, A=proc.adjacency_matrix).loglik(data, conv)) / (2 * h)
            assert g["W"][p, c] == pytest.approx(fd, rel=1e-4, abs=1e-3)
