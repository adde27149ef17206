"""GPU path against the independent known-answer vectors KAT-F..K (tests/golden/make_kat.py: numpy/scipy restatements of the
cited Julia, written without oracle/): rows a5, a9, Q3, a12, a13, a14 of SURVEY.md section 8 -- so a reading error shared by the
oracle and the kernels cannot pass unnoticed."""
import numpy as np
import pytest

import nhp_b200 as nhp
from nhp_b200 import discrete as D

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def kat_b_process(kat, network):
    k = kat["B"]
    base = nhp.HomogeneousProcess(k["lambda0"])
    imp = nhp.LogitNormalImpulseResponse(np.full((2, 2), k["mu"]), np.full((2, 2), k["tau"]), k["dtmax"])
    wts = nhp.DenseWeightModel(np.array(k["W"]))
    data = (np.array(k["events"]), np.array(k["nodes"]), k["duration"])
    if network:
        return nhp.ContinuousNetworkHawkesProcess(base, imp, wts, np.array(k["A"]), nhp.BernoulliNetworkModel(0.5, 2)), data
    return nhp.ContinuousStandardHawkesProcess(base, imp, wts), data


def test_kat_f_intensity_at_query_times(kat):
    tq = np.array(kat["F"]["times"])
    for network, key in ((False, "standard"), (True, "network")):
        proc, data = kat_b_process(kat, network)
        np.testing.assert_allclose(nhp.intensity(proc, data, tq), kat["F"][key], rtol=RTOL)


def test_kat_g_adjacency_sweep(kat):
    for case in kat["G"]["cases"]:
        proc, data = kat_b_process(kat, True)
        proc.network = nhp.BernoulliNetworkModel(case["rho"], 2)
        proc.adjacency_matrix = np.array(kat["G"]["A0"])
        A = nhp.resample_adjacency_matrix_(proc, data, u=np.array(case["u"]))
        np.testing.assert_array_equal(A, case["A"])


def test_kat_h_recursive_network_loglik(kat):
    k = kat["H"]
    proc = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(k["lambda0"]), nhp.ExponentialImpulseResponse(np.array(k["theta"])),
                                              nhp.DenseWeightModel(np.array(k["W"])), np.array(k["A"]), nhp.BernoulliNetworkModel(0.5, 2))
    data = (np.array(k["events"]), np.array(k["nodes"]), k["duration"])
    np.testing.assert_allclose(nhp.event_intensity(proc, data), k["intensities"], rtol=RTOL)
    assert nhp.loglikelihood(proc, data, recursive=True) == pytest.approx(k["ll_recursive"], rel=RTOL)
    assert nhp.loglikelihood(proc, data, recursive=False) == pytest.approx(k["ll_windowed"], rel=RTOL)


def kat_i_process(kat, A=None):
    k = kat["I"]
    data = np.array(k["data"], dtype=np.int64)
    base, imp, wts = D.DiscreteHomogeneousProcess(k["lambda0"]), D.DiscreteGaussianImpulseResponse(np.array(k["theta"]), 4), nhp.DenseWeightModel(np.array(k["W"]))
    if A is None:
        return k, data, D.DiscreteStandardHawkesProcess(base, imp, wts)
    return k, data, D.DiscreteNetworkHawkesProcess(base, imp, wts, A, nhp.BernoulliNetworkModel(0.5, 2))


def test_kat_i_discrete_gibbs_counts(kat):
    k, data, proc = kat_i_process(kat)
    np.testing.assert_allclose(D.convolve(proc, data), k["conv"], rtol=RTOL, atol=1e-300)
    np.testing.assert_array_equal(D.resample_parents(proc, data, u=np.array(k["u"])), k["counts"])


def test_kat_j_vb_statistics(kat):
    k, data, proc = kat_i_process(kat)
    j = kat["J"]
    st = D.vb_statistics(proc, data, np.array(j["e0"]), np.array(j["E"]))
    np.testing.assert_allclose(st["alpha_sum"], j["alpha_sum"], rtol=RTOL)
    np.testing.assert_allclose(st["gamma_sum"], j["gamma_sum"], rtol=RTOL, atol=1e-300)
    np.testing.assert_allclose(st["kappa_sum"], j["kappa_sum"], rtol=RTOL)
    np.testing.assert_array_equal(st["nu_sum"], j["nu_sum"])


def test_kat_k_discrete_adjacency(kat):
    for case in kat["K"]["cases"]:
        k, data, proc = kat_i_process(kat, A=np.array(kat["K"]["A0"]))
        proc.network = nhp.BernoulliNetworkModel(case["rho"], 2)
        A = D.resample_adjacency_matrix_(proc, data, u=np.array(case["u"]))
        np.testing.assert_array_equal(A, case["A"])
