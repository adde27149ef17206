"""The quick-start snippet of README.md, runnable as a check (python tools/readme_snippet.py from the repo root)."""
import sys; sys.path.insert(0, "networkhawkesprocesses.jl_b200")
import numpy as np, nhp_b200 as nhp
K = 50
proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)),
                                           nhp.LogitNormalImpulseResponse(np.zeros((K, K)), np.ones((K, K)), 1.0),
                                           nhp.DenseWeightModel(np.full((K, K), 0.5 / K)))
data = proc.upload(nhp.rand(proc, 200.0, np.random.default_rng(0)))
ll = nhp.loglikelihood(proc, data)
ll2, grads = nhp.loglikelihood_gradient(proc, data)
parents, parentnodes = nhp.resample_parents(proc, data, seed=1)
chain = nhp.mcmc_device_(proc, data, nsteps=200)
print("readme snippet ok", ll, ll2, len(parents), len(chain.samples))
