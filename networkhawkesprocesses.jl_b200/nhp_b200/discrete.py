"""Host-side mirror of the reference's discrete-time API over libnhp.

  components  DiscreteHomogeneousProcess (baselines.jl:358-382), DiscreteGaussianImpulseResponse (impulses.jl:272-288)
  processes   DiscreteStandardHawkesProcess (discrete.jl:161-170), DiscreteNetworkHawkesProcess (discrete.jl:395-402)
  functions   basis, convolve, intensity, loglikelihood, resample_parents (reduced over t: counts[c, k]),
              update_ (one VB step, discrete.jl:369-375), resample_ (one Gibbs sweep, discrete.jl:361-367 / 416-424),
              resample_adjacency_matrix_, mcmc_, vb_
`data` is an N x T integer matrix (numpy row = node) as in the reference; arrays indexed theta[p, c, b].
"""
import ctypes
import time
import weakref

import numpy as np
from scipy.special import digamma

from .continuous import BernoulliNetworkModel, DenseNetworkModel, DenseWeightModel  # noqa: F401
from .core import _f64, _fmat, _ptr, default_context


class DiscreteHomogeneousProcess:
    """baselines.jl:358-382"""

    def __init__(self, lam, dt=1.0, alpha0=1.0, beta0=1.0):
        self.lam = np.array(lam, dtype=np.float64).reshape(-1)
        if np.any(self.lam < 0):
            raise ValueError("DiscreteHomogeneousProcess: intensity parameter λ must be non-negative")
        if not dt > 0:
            raise ValueError("DiscreteHomogeneousProcess: time step dt must be non-negative")
        self.dt, self.alpha0, self.beta0 = float(dt), float(alpha0), float(beta0)
        self.alphav, self.betav = np.ones_like(self.lam), np.ones_like(self.lam)

    def ndims(self):
        return self.lam.size

    def variational_log_expectation(self):
        return digamma(self.alphav) - np.log(self.betav)  # baselines.jl:454-456


class DiscreteGaussianImpulseResponse:
    """impulses.jl:272-288: theta[p, c, b] (rows sum to one over b), nlags, dt."""

    def __init__(self, theta, nlags, dt=1.0, gamma=1.0):
        self.theta = np.array(theta, dtype=np.float64)
        if not np.all(np.sum(self.theta, axis=2) == 1.0) and not np.allclose(np.sum(self.theta, axis=2), 1.0):
            raise ValueError("Invalid discrete basis parameter.")  # impulses.jl:282
        self.nlags, self.dt, self.gamma = int(nlags), float(dt), float(gamma)
        self.gammav = np.ones_like(self.theta)

    def nbasis(self):
        return self.theta.shape[2]

    def basis(self):
        """impulses.jl:321-335 -> phi[l, b] (computed by libnhp's nhp_disc_basis)."""
        from . import _lib
        L, B = self.nlags, self.nbasis()
        phi = np.empty(L * B)
        rc = _lib.load().nhp_disc_basis(L, B, self.dt, _ptr(phi))
        if rc != 0:
            raise ValueError("invalid basis dimensions")
        return phi.reshape(B, L).T.copy()

    def variational_log_expectation(self):
        return digamma(self.gammav) - digamma(np.sum(self.gammav, axis=2, keepdims=True))  # impulses.jl:373-375


class DiscreteData:
    """Device-resident count matrix + its convolution (nhp_disc)."""

    def __init__(self, ctx, data, t_halo=0):
        self.ctx = ctx
        d = np.asarray(data)
        if d.ndim != 2:
            raise ValueError("data must be an N x T matrix")
        self.N, self.T = d.shape
        flat = np.ascontiguousarray(d.T, dtype=np.int64).ravel()  # data[n + N*t]
        h = ctypes.c_void_p()
        ctx.check(ctx.lib.nhp_disc_upload(ctx.h, _ptr(flat), self.N, self.T, int(t_halo), ctypes.byref(h)))
        self.h = h
        self.convolved = False
        self._fin = weakref.finalize(self, ctx.lib.nhp_disc_free, ctx.h, h)

    def free(self):
        self._fin()


class DiscreteHawkesProcess:
    adjacency_matrix = None

    def ndims(self):
        return self.baseline.ndims()

    def _ctx(self):
        if getattr(self, "ctx", None) is None:
            self.ctx = default_context()
        return self.ctx

    def _theta_flat(self, th):
        return np.ascontiguousarray(np.asarray(th, dtype=np.float64).transpose(2, 1, 0)).ravel()  # theta[p + N*(c + N*b)]

    def _push(self, ctx):
        N, B = self.ndims(), self.impulses.nbasis()
        if self.weights.W.shape != (N, N) or self.impulses.theta.shape != (N, N, B):
            raise ValueError("parameter shapes do not match the number of nodes")
        A = None if self.adjacency_matrix is None else _fmat(self.adjacency_matrix)
        lam, W, th = _f64(self.baseline.lam), _fmat(self.weights.W), self._theta_flat(self.impulses.theta)
        ctx.check(ctx.lib.nhp_disc_params_set(ctx.h, N, B, _ptr(lam), _ptr(W), _ptr(A), _ptr(th), self.dt))

    def upload(self, data):
        return data if isinstance(data, DiscreteData) else DiscreteData(self._ctx(), data)

    def _ready(self, data):
        d = self.upload(data)
        if not d.convolved:
            convolve(self, d, export=False)
        return d


class DiscreteStandardHawkesProcess(DiscreteHawkesProcess):
    """discrete.jl:161-170"""

    def __init__(self, baseline, impulses, weights, dt=1.0):
        if baseline.dt != dt or impulses.dt != dt:
            raise ValueError("Baseline and impulse response time step must match process time step.")  # intent of discrete.jl:167 (quirk Q13)
        self.baseline, self.impulses, self.weights, self.dt = baseline, impulses, weights, float(dt)

    def params(self):
        eff = (self.weights.W[:, :, None] * self.impulses.theta).transpose(2, 1, 0).ravel()  # vec(W .* theta)
        return np.concatenate([self.baseline.lam, eff])  # discrete.jl:178-182

    def variational_params(self):
        b, w, i = self.baseline, self.weights, self.impulses
        return np.concatenate([b.alphav, b.betav, i.gammav.transpose(2, 1, 0).ravel(), w.kappav.T.ravel(), w.nuv.T.ravel()])  # discrete.jl:204-209


class DiscreteNetworkHawkesProcess(DiscreteHawkesProcess):
    """discrete.jl:395-402"""

    def __init__(self, baseline, impulses, weights, adjacency_matrix, network, dt=1.0):
        self.baseline, self.impulses, self.weights, self.network, self.dt = baseline, impulses, weights, network, float(dt)
        self.adjacency_matrix = np.array(adjacency_matrix, dtype=np.float64)

    def params(self):
        return np.concatenate([self.network.params(), self.baseline.lam, self.weights.W.T.ravel(), self.impulses.theta.transpose(2, 1, 0).ravel(),
                               self.adjacency_matrix.T.ravel()])  # discrete.jl:406-414


# ------------------------------------------------------------------------------------------
def convolve(process, data, export=True):
    """discrete.jl:146-151 -> conv[t, n, b]; the result also stays on the device inside `data` (DiscreteData)."""
    ctx = process._ctx()
    d = process.upload(data)
    phi = process.impulses.basis()
    L, B = phi.shape
    ph = np.ascontiguousarray(phi.T).ravel()
    out = np.empty(d.T * d.N * B) if export else None
    ctx.check(ctx.lib.nhp_disc_convolve(ctx.h, d.h, _ptr(ph), L, B, _ptr(out)))
    d.convolved = True
    return out.reshape(B, d.N, d.T).transpose(2, 1, 0).copy() if export else None


def intensity(process, data):
    """discrete.jl:115-129 -> lam[t, c]."""
    ctx = process._ctx()
    d = process._ready(data)
    process._push(ctx)
    lam = np.empty(d.T * d.N)
    ctx.check(ctx.lib.nhp_disc_intensity(ctx.h, d.h, _ptr(lam)))
    return lam.reshape(d.N, d.T).T.copy()


def loglikelihood(process, data):
    """discrete.jl:86-102."""
    ctx = process._ctx()
    d = process._ready(data)
    process._push(ctx)
    ll = ctypes.c_double()
    ctx.check(ctx.lib.nhp_disc_loglik(ctx.h, d.h, ctypes.byref(ll)))
    return ll.value


def loglikelihood_gradient(process, data):
    """Extension for `mle!` (discrete.jl:211-296; the reference differentiates numerically): the log-likelihood and its analytic
    gradient (nhp_disc_loglik_grad).  Returns `(ll, grads)` with `lambda0 [N]`, `W [p, c]`, `theta [p, c, b]`."""
    ctx = process._ctx()
    d = process._ready(data)
    process._push(ctx)
    N, B = d.N, process.impulses.nbasis()
    ll = ctypes.c_double()
    g0, gW, gT = np.empty(N), np.empty(N * N), np.empty(N * N * B)
    ctx.check(ctx.lib.nhp_disc_loglik_grad(ctx.h, d.h, ctypes.byref(ll), _ptr(g0), _ptr(gW), _ptr(gT)))
    return ll.value, dict(lambda0=g0, W=gW.reshape(N, N).T.copy(), theta=gT.reshape(B, N, N).transpose(2, 1, 0).copy())


def resample_parents(process, data, seed=0, counter=0, u=None):
    """parents.jl:82-117 reduced over t: counts[c, k], k = 0 baseline, k = 1 + p*B + b."""
    ctx = process._ctx()
    d = process._ready(data)
    process._push(ctx)
    N, B = d.N, process.impulses.nbasis()
    NK = 1 + N * B
    counts = np.empty(N * NK)
    uu = None if u is None else _f64(u)
    ctx.check(ctx.lib.nhp_disc_gibbs_counts(ctx.h, d.h, int(seed), int(counter), _ptr(uu), 0 if uu is None else uu.size, _ptr(counts)))
    return counts.reshape(NK, N).T.copy()


def vb_statistics(process, data, e0, E):
    """parents.jl:136-177 fused with baselines.jl:444-452, weights.jl:70-91, impulses.jl:355-371."""
    ctx = process._ctx()
    d = process._ready(data)
    N, B = d.N, process.impulses.nbasis()
    a, k, nu, g = np.empty(N), np.empty(N * N), np.empty(N * N), np.empty(N * N * B)
    Ef = process._theta_flat(E)
    ctx.check(ctx.lib.nhp_disc_vb_stats(ctx.h, d.h, _ptr(_f64(e0)), _ptr(Ef), _ptr(a), _ptr(k), _ptr(nu), _ptr(g)))
    f = lambda v: v.reshape(N, N).T.copy()
    return dict(alpha_sum=a, kappa_sum=f(k), nu_sum=f(nu), gamma_sum=g.reshape(B, N, N).transpose(2, 1, 0).copy())


def update_(process, data):
    """One mean-field VB step, `update!` (discrete.jl:369-375)."""
    b, w, imp = process.baseline, process.weights, process.impulses
    if not hasattr(w, "kappav"):
        w.kappav, w.nuv = np.ones_like(w.W), np.ones_like(w.W)
    e0 = np.exp(b.variational_log_expectation())
    ElogW = digamma(w.kappav) - np.log(w.nuv)  # weights.jl:95-97
    E = np.exp(imp.variational_log_expectation() + ElogW[:, :, None])
    st = vb_statistics(process, data, e0, E)
    d = process.upload(data)
    b.alphav = b.alpha0 + st["alpha_sum"]
    b.betav = 1.0 / b.beta0 + d.T * process.dt * np.ones(d.N)  # baselines.jl:450 (as written in the reference)
    w.kappav = w.kappa + st["kappa_sum"]
    w.nuv = w.nu + st["nu_sum"]
    imp.gammav = imp.gamma + st["gamma_sum"]
    return process.variational_params()


def resample_adjacency_matrix_(process, data, seed=0, counter=0, u=None):
    """discrete.jl:426-460; mutates process.adjacency_matrix."""
    ctx = process._ctx()
    d = process._ready(data)
    process._push(ctx)
    N = d.N
    A = _fmat(process.adjacency_matrix).copy()
    rho = _fmat(process.network.link_probability())
    uu = None if u is None else _fmat(u)
    ctx.check(ctx.lib.nhp_disc_resample_adjacency(ctx.h, d.h, _ptr(rho), int(seed), int(counter), _ptr(uu), _ptr(A)))
    process.adjacency_matrix = A.reshape(N, N).T.copy()
    return process.adjacency_matrix


def resample_(process, data, rng, seed=0, counter=0):
    """One Gibbs sweep, `resample!` (discrete.jl:361-367 / 416-424); conjugate draws on the host."""
    d = process._ready(data)
    N, B = d.N, process.impulses.nbasis()
    C = resample_parents(process, d, seed=seed, counter=counter)  # [c, k]
    b, w, imp = process.baseline, process.weights, process.impulses
    b.lam = rng.gamma(b.alpha0 + C[:, 0], 1.0 / (b.beta0 + d.T * process.dt))  # intended form of baselines.jl:413-419 (quirk Q2)
    Cpb = C[:, 1:].reshape(N, N, B)  # [c, p, b]
    Mnm = Cpb.sum(axis=2).T  # [p, c]
    Mn = vb_row_sums(process, d)
    w.W = rng.gamma(w.kappa + Mnm, (1.0 / (w.nu + Mn))[:, None] * np.ones((N, N)))
    g = imp.gamma + Cpb.transpose(1, 0, 2)  # [p, c, b]
    th = rng.gamma(g, 1.0)
    imp.theta = th / th.sum(axis=2, keepdims=True)  # Dirichlet (impulses.jl:337-353)
    if process.adjacency_matrix is not None:
        resample_adjacency_matrix_(process, d, seed=seed, counter=counter + (1 << 40))
        process.network.resample_(process.adjacency_matrix, rng)
    return process.params()


def resample_on_device_(process, data, rng=None, seed=0, counter=0):
    """One Gibbs sweep, `resample!` (discrete.jl:361-367 / 416-424), with the conjugate draws on the device: the counts of the parent
    sweep stay there (nhp_disc_gibbs_counts with no host copy) and nhp_disc_resample_params draws lambda0, W and the Dirichlet theta
    from them (Philox: seed, counter); the adjacency sweep and the network draw follow as in `resample_`.  `rng` only feeds the
    host-side network draw."""
    d = process._ready(data)
    ctx = process._ctx()
    N, B = d.N, process.impulses.nbasis()
    b, w, imp = process.baseline, process.weights, process.impulses
    Mn = _f64(vb_row_sums(process, d))  # events per node (cached with the data)
    process._push(ctx)
    ctx.check(ctx.lib.nhp_disc_gibbs_counts(ctx.h, d.h, int(seed), int(counter), None, 0, None))
    hy = np.array([b.alpha0, b.beta0, w.kappa, w.nu, imp.gamma], dtype=np.float64)
    lam, W, th = np.empty(N), np.empty(N * N), np.empty(N * N * B)
    ctx.check(ctx.lib.nhp_disc_resample_params(ctx.h, d.h, int(seed), int(counter), _ptr(Mn), _ptr(hy), hy.size, _ptr(lam), _ptr(W), _ptr(th)))
    b.lam = lam
    w.W = W.reshape(N, N).T.copy()                             # W[p + N c] -> [p, c]
    imp.theta = th.reshape(B, N, N).transpose(2, 1, 0).copy()  # theta[p + N (c + N b)] -> [p, c, b]
    if process.adjacency_matrix is not None:
        resample_adjacency_matrix_(process, d, seed=seed, counter=counter + (1 << 40))
        process.network.resample_(process.adjacency_matrix, rng if rng is not None else np.random.default_rng(seed + counter))
    return process.params()


def vb_row_sums(process, d):
    if not hasattr(d, "_rowsum"):
        N, B = d.N, process.impulses.nbasis()
        st = vb_statistics(process, d, np.ones(N), np.ones((N, N, B)))
        d._rowsum = st["nu_sum"][:, 0].copy()
    return d._rowsum


def mcmc_(process, data, nsteps=1000, seed=0):
    rng = np.random.default_rng(seed)
    d = process._ready(data)
    t0 = time.time()
    samples = [resample_(process, d, rng, seed=seed, counter=s) for s in range(nsteps)]
    from .continuous import MarkovChainMonteCarlo
    return MarkovChainMonteCarlo(samples, time.time() - t0)


def vb_(process, data, max_steps=100):
    """`vb!` (inference.jl:153-181; the reference's convergence test is commented out, so it runs max_steps)."""
    d = process._ready(data)
    trace = [update_(process, d) for _ in range(max_steps)]
    return trace
