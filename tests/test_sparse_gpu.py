"""GPU parity tests of the sparse-adjacency sweep (csrc/cont_sparse.cu) against the oracle and against
the dense sweep: same results whichever path runs (NHP_SPARSE=1 forces it, =0 disables it)."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,K,n,rate,density,dtmax", [
    ("ln", 60, 40000, 64.0, 0.05, 1.0),      # typical: ~3 hits per window
    ("ln", 1000, 30000, 64.0, 0.05, 1.0),    # BASELINE config 4's exact shape (K = 1000, rho = 0.05, mean window 64: SG = 2) at an oracle-sized N
    ("ln", 1200, 30000, 64.0, 0.05, 1.0),    # K > 1024: bit rows longer than one warp load
    ("ln", 30, 20000, 200.0, 0.6, 1.0),      # dense-ish: slot overflow -> direct path for most events
    ("ln", 12, 20000, 300.0, None, 1.0),     # Standard process forced through the sparse kernel (every pair active)
    ("exp", 40, 30000, 50.0, 0.1, 1.5),
    ("ln", 8, 30000, 3000.0, 0.2, 5.0),      # window longer than the staging buffer -> global-memory direct path
])
def test_sparse_path_matches_oracle_and_dense(kind, K, n, rate, density, dtmax, monkeypatch):
    t, nodes, T = synth.poisson_stream(n, K, rate, 70 + K)
    if kind == "ln":
        proc, om = make_ln(K, 80 + K, density=density, wmax=0.3 / (K * (density or 1.0)), dtmax=dtmax)
    else:
        proc, om = make_exp(K, 80 + K, density=density, wmax=0.3 / (K * (density or 1.0)), dtmax=dtmax)
    d = proc.upload((t, nodes, T))
    u = np.random.default_rng(9).random(n)
    ref_ll = om.loglik(t, nodes, T, recursive=False)
    ref_lam = om.event_intensity(t, nodes)
    ref_par, ref_pn = om.resample_parents(t, nodes, u)
    ost = orc.suffstats(1 if kind == "ln" else 0, t, nodes, ref_par, ref_pn, K, dtmax)
    for mode in ("1", "0"):
        monkeypatch.setenv("NHP_SPARSE", mode)
        assert nhp.loglikelihood(proc, d, recursive=False) == pytest.approx(ref_ll, rel=1e-10)
        np.testing.assert_allclose(nhp.event_intensity(proc, d), ref_lam, rtol=1e-10)
        par, pn = nhp.resample_parents(proc, d, u=u, with_loglik=True)
        assert nhp.sweep_loglikelihood(proc, d) == pytest.approx(ref_ll, rel=1e-10)  # fused with the sweep
        assert np.count_nonzero(par != ref_par) == 0
        np.testing.assert_array_equal(pn, ref_pn)
        st = nhp.sufficient_statistics(proc, d)
        for key in ("M0", "Mn", "Mnm"):
            np.testing.assert_array_equal(st[key], ost[key])
        np.testing.assert_allclose(st["S1"], ost["S1"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(st["S2"], ost["S2"], rtol=1e-10, atol=1e-12)


def test_sparse_path_with_halo_shards(monkeypatch):
    monkeypatch.setenv("NHP_SPARSE", "1")
    K, n = 25, 30000
    t, nodes, T = synth.poisson_stream(n, K, 80.0, 6)
    proc, om = make_ln(K, 8, density=0.1, wmax=0.2)
    ctx = proc._ctx()
    ref = om.loglik(t, nodes, T)
    u = np.random.default_rng(3).random(n)
    ref_par, _ = om.resample_parents(t, nodes, u)
    total, pars = 0.0, []
    bounds = [0, 9000, 21000, n]
    for r in range(3):
        a, b = bounds[r], bounds[r + 1]
        lo = int(np.searchsorted(t, t[a] - 1.0, side="right")) if a > 0 else 0
        d = nhp.ContinuousData(ctx, t[lo:b], nodes[lo:b], T, K, n_halo=a - lo, index_base=lo, flags=1 if r == 0 else 0)
        total += nhp.loglikelihood(proc, d)
        pars.append(nhp.resample_parents(proc, d, u=u[a:b])[0])
    assert total == pytest.approx(ref, rel=1e-11)
    np.testing.assert_array_equal(np.concatenate(pars), ref_par)


def test_sparse_edge_cases(monkeypatch):
    monkeypatch.setenv("NHP_SPARSE", "1")
    proc, om = make_ln(3, 1, density=0.5, wmax=0.3)
    t = np.array([0.0, 0.0, 0.5, 0.5, 0.5, 1.2, 1.49999, 1.5])
    nodes = np.array([1, 2, 3, 1, 1, 2, 3, 3], dtype=np.int64)
    assert nhp.loglikelihood(proc, (t, nodes, 2.0)) == pytest.approx(om.loglik(t, nodes, 2.0), rel=1e-12)
    # an all-zero adjacency: every event is a baseline event
    proc.adjacency_matrix[:] = 0.0
    par, pn = nhp.resample_parents(proc, (t, nodes, 2.0), seed=1)
    assert np.all(par == 0) and np.all(pn == 0)
    assert nhp.loglikelihood(proc, (t, nodes, 2.0)) == pytest.approx(-np.sum(proc.baseline.lam) * 2.0 + np.sum(np.log(proc.baseline.lam[nodes - 1])), rel=1e-13)
