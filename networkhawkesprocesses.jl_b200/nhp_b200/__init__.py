"""nhp_b200 -- host-side mirror of NetworkHawkesProcesses.jl's event-history API over libnhp
(hand-written sm_100a CUDA behind the C ABI of include/nhp.h).  No CPU fallback."""
from ._lib import NHPError, LIB_PATH, load  # noqa: F401
from .core import Context, ContinuousData, default_context  # noqa: F401
from .continuous import (  # noqa: F401
    HomogeneousProcess, LogGaussianCoxProcess, ExponentialImpulseResponse, LogitNormalImpulseResponse, DenseWeightModel, SparseWeightModel,
    DenseNetworkModel, BernoulliNetworkModel, ContinuousStandardHawkesProcess, ContinuousNetworkHawkesProcess,
    loglikelihood, loglikelihood_gradient, gradient_vector, event_intensity, intensity, resample_parents, sweep_loglikelihood, sufficient_statistics, resample_adjacency_matrix_,
    resample_, resample_on_device_, pull_params_, adjacency_info, mcmc_, mcmc_device_, mle_, rand, rand_device, MarkovChainMonteCarlo, MaximumLikelihood)
from . import discrete  # noqa: F401,E402
from .discrete import (  # noqa: F401,E402
    DiscreteHomogeneousProcess, DiscreteGaussianImpulseResponse, DiscreteStandardHawkesProcess, DiscreteNetworkHawkesProcess, DiscreteData)
