"""GPU parity tests of the child-major dense sweep (csrc/cont_child.cu): same results as the oracle and as
the time-tiled sweep (NHP_CHILD=1 forces it, =0 disables it)."""
import numpy as np
import pytest

import nhp_b200 as nhp
import oracle_ffi as orc
import synth
from test_cont_gpu import make_exp, make_ln

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,K,n,rate,density,dtmax", [
    ("ln", 40, 30000, 80.0, None, 1.0),
    ("exp", 300, 40000, 60.0, None, 1.5),
    ("ln", 7, 20000, 400.0, 0.6, 1.0),      # network model with a dense adjacency (sparse path not applicable)
    ("exp", 3, 3000, 5.0, None, np.inf),    # full history through the cut-off horizon
])
def test_child_major_sweep_matches_oracle(kind, K, n, rate, density, dtmax, monkeypatch):
    monkeypatch.setenv("NHP_SPARSE", "0")
    t, nodes, T = synth.poisson_stream(n, K, rate, 90 + K)
    if kind == "ln":
        proc, om = make_ln(K, 91 + K, density=density, wmax=0.4 / K, dtmax=dtmax)
    else:
        proc, om = make_exp(K, 91 + K, density=density, wmax=0.4 / K, dtmax=dtmax)
    d = proc.upload((t, nodes, T))
    u = np.random.default_rng(4).random(n)
    ref_ll = om.loglik(t, nodes, T, recursive=False)
    ref_lam = om.event_intensity(t, nodes)
    ref_par, ref_pn = om.resample_parents(t, nodes, u)
    ost = orc.suffstats(1 if kind == "ln" else 0, t, nodes, ref_par, ref_pn, K, dtmax)
    for mode in ("1", "0"):
        monkeypatch.setenv("NHP_CHILD", mode)
        assert nhp.loglikelihood(proc, d, recursive=False) == pytest.approx(ref_ll, rel=1e-10)
        np.testing.assert_allclose(nhp.event_intensity(proc, d), ref_lam, rtol=1e-10)
        par, pn = nhp.resample_parents(proc, d, u=u, with_loglik=True)
        assert nhp.sweep_loglikelihood(proc, d) == pytest.approx(ref_ll, rel=1e-10)  # fused with the sweep
        assert np.count_nonzero(par != ref_par) == 0
        st = nhp.sufficient_statistics(proc, d)
        for key in ("M0", "Mn", "Mnm"):
            np.testing.assert_array_equal(st[key], ost[key])
        np.testing.assert_allclose(st["S1"], ost["S1"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(st["S2"], ost["S2"], rtol=1e-10, atol=1e-12)


def test_child_major_with_halo_shards(monkeypatch):
    monkeypatch.setenv("NHP_CHILD", "1")
    monkeypatch.setenv("NHP_SPARSE", "0")
    K, n = 15, 20000
    t, nodes, T = synth.poisson_stream(n, K, 50.0, 8)
    proc, om = make_ln(K, 9, wmax=0.03)
    ctx = proc._ctx()
    ref = om.loglik(t, nodes, T)
    u = np.random.default_rng(3).random(n)
    ref_par, _ = om.resample_parents(t, nodes, u)
    total, pars = 0.0, []
    bounds = [0, 6000, 13000, n]
    for r in range(3):
        a, b = bounds[r], bounds[r + 1]
        lo = int(np.searchsorted(t, t[a] - 1.0, side="right")) if a > 0 else 0
        dd = nhp.ContinuousData(ctx, t[lo:b], nodes[lo:b], T, K, n_halo=a - lo, index_base=lo, flags=1 if r == 0 else 0)
        total += nhp.loglikelihood(proc, dd)
        pars.append(nhp.resample_parents(proc, dd, u=u[a:b])[0])
    assert total == pytest.approx(ref, rel=1e-11)
    np.testing.assert_array_equal(np.concatenate(pars), ref_par)
