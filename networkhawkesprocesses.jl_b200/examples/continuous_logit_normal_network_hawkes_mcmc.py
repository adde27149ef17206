"""Gibbs sampling of a continuous logit-normal NETWORK Hawkes process (examples/continuous-logit-normal-network-hawkes.jl):
parents + statistics + adjacency columns on the GPU, conjugate draws on the host."""
import _path  # noqa: F401
import numpy as np

import nhp_b200 as nhp

rng = np.random.default_rng(0)
K, duration, dtmax = 4, 500.0, 1.0
A = (rng.random((K, K)) < 0.5).astype(float)
truth = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.LogitNormalImpulseResponse(np.zeros((K, K)), np.ones((K, K)), dtmax),
                                           nhp.DenseWeightModel(0.25 * np.ones((K, K))), A, nhp.BernoulliNetworkModel(0.5, K))
data = nhp.rand(truth, duration, rng)
model = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.LogitNormalImpulseResponse(np.zeros((K, K)), np.ones((K, K)), dtmax),
                                           nhp.DenseWeightModel(0.1 * np.ones((K, K))), np.ones((K, K)), nhp.BernoulliNetworkModel(0.5, K))
res = nhp.mcmc_(model, data, nsteps=500, seed=1)
s = np.array(res.samples)[100:]
print("events:", len(data[0]), " sweeps/s:", len(res.samples) / res.elapsed)
print("posterior mean lambda0:", s[:, 1:1 + K].mean(axis=0), " truth:", truth.baseline.lam)
print("posterior link probabilities:\n", s[:, -K * K:].mean(axis=0).reshape(K, K).T.round(2), "\ntrue adjacency:\n", A)
