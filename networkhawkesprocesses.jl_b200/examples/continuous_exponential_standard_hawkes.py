"""README example of the reference (README.md:27-38; examples/continuous-exponential-standard-hawkes.jl):
2-node exponential process, duration 1000, loglikelihood + mle!.  Runs on a B200 through libnhp."""
import _path  # noqa: F401
import numpy as np

import nhp_b200 as nhp

nnodes, duration = 2, 1000.0
process = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(nnodes)), nhp.ExponentialImpulseResponse(np.ones((nnodes, nnodes))),
                                              nhp.DenseWeightModel(0.1 * np.ones((nnodes, nnodes))))
data = nhp.rand(process, duration, np.random.default_rng(0))
truth = process.params()
print("events:", len(data[0]), " loglikelihood at the generating parameters:", nhp.loglikelihood(process, data))
res = nhp.mle_(process, data, seed=0)
print("mle status:", res.status, " maximum:", res.maximum)
print(np.column_stack([truth, res.maximizer]))  # [theta_true theta_estimated], as the reference example prints
