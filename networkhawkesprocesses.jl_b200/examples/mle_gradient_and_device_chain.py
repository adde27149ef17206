"""The two extensions on either side of the hot path:
 1. `mle!` with the analytic gradient sweep (nhp_cont_loglik_grad) next to the reference's finite-difference behaviour;
 2. a Gibbs chain whose conjugate draws stay on the GPU (nhp_cont_resample_params), parameters pulled every 10th sweep."""
import time

import _path  # noqa: F401
import numpy as np

import nhp_b200 as nhp

rng = np.random.default_rng(0)
K = 3
truth = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(1.5 * np.ones((K, K))),
                                            nhp.DenseWeightModel(0.2 * np.ones((K, K))))
data = nhp.rand(truth, 2000.0, rng)
print("events:", len(data[0]))
for mode in ("analytic", "finite"):
    fit = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))),
                                              nhp.DenseWeightModel(0.1 * np.ones((K, K))))
    t0 = time.time()
    res = nhp.mle_(fit, data, guess=np.full(fit.params().size, 0.5), gradient=mode)
    print(f"mle ({mode:8s} gradient): loglik {res.maximum:.3f} after {res.steps} iterations, {time.time() - t0:.2f} s; lambda0 = {fit.baseline.lam.round(3)}")

chain = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))),
                                            nhp.DenseWeightModel(0.1 * np.ones((K, K))))
res = nhp.mcmc_(chain, data, nsteps=1000, seed=1, device_draws=True, store_every=10)
s = np.array(res.samples)[10:]
print(f"device-resident chain: {1000 / res.elapsed:.0f} sweeps/s, {len(res.samples)} stored samples")
print("posterior mean lambda0:", s[:, :K].mean(axis=0).round(3), " W[0,:]:", s[:, K + K * K:K + K * K + K].mean(axis=0).round(3))
