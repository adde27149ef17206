"""Known-answer vectors for the oracle, derived INDEPENDENTLY of oracle/ from the formulas cited
in SURVEY.md section 8c (reference file:line in each block) using only numpy/scipy.

The reference itself cannot run here (no Julia), so these are formula-derived, not
reference-run, vectors: parity stays "unpinned" in the strict sense (see oracle/nhp_oracle.h).
Run:  python tests/golden/make_kat.py   ->  tests/golden/kat.json
"""
import json
import os

import numpy as np
from scipy import stats
from scipy.special import gammaln, logit

out = {}


def exp_pdf(theta, dt):  # impulses.jl:106-108
    return stats.expon(scale=1.0 / theta).pdf(dt)


def ln_pdf(mu, tau, x):  # impulses.jl:174-178 (no 1/dtmax Jacobian)
    if not (0.0 < x < 1.0):
        return 0.0
    return stats.norm(mu, tau ** -0.5).pdf(logit(x)) / (x * (1.0 - x))


def window_intensities(events, nodes, lam0, W, A, pdf, dtmax):  # continuous.jl:286-300
    lam = []
    for i, (t, c) in enumerate(zip(events, nodes)):
        v = lam0[c - 1]
        for j in range(i - 1, -1, -1):
            if not events[j] > t - dtmax:
                break
            p = nodes[j]
            a = 1.0 if A is None else A[p - 1][c - 1]
            v += a * W[p - 1][c - 1] * pdf(p - 1, c - 1, t - events[j])
        lam.append(v)
    return lam


def loglik(events, nodes, T, lam0, W, A, lam):  # continuous.jl:210-239 / 360-389
    ll = -sum(l * T for l in lam0)
    for p in nodes:
        ll -= sum((1.0 if A is None else A[p - 1][c]) * W[p - 1][c] for c in range(len(lam0)))
    return ll + sum(np.log(lam))


# KAT-A: README Exponential example (README.md:27-38), dtmax = Inf
ev = [0.5, 1.0, 1.5, 3.0]
nd = [1, 2, 1, 2]
lam0 = [1.0, 1.0]
W = [[0.1, 0.1], [0.1, 0.1]]
lamA = window_intensities(ev, nd, lam0, W, None, lambda p, c, dt: exp_pdf(1.0, dt), np.inf)
out["A"] = dict(events=ev, nodes=nd, duration=4.0, lambda0=lam0, W=W, theta=[[1.0, 1.0], [1.0, 1.0]],
                intensities=lamA, ll=loglik(ev, nd, 4.0, lam0, W, None, lamA))

# KAT-B: LogitNormal tutorial parameters, dtmax = 1, Standard and Network
ev = [0.1, 0.4, 0.9, 1.3, 2.5, 2.6]
nd = [1, 2, 1, 2, 1, 1]
lam0 = [1.0, 2.0]
W = [[0.1, 0.2], [0.2, 0.1]]
A = [[1.0, 0.0], [1.0, 1.0]]
pdfB = lambda p, c, dt: ln_pdf(1.0, 1.0, dt / 1.0)
lamB = window_intensities(ev, nd, lam0, W, None, pdfB, 1.0)
lamBn = window_intensities(ev, nd, lam0, W, A, pdfB, 1.0)
out["B"] = dict(events=ev, nodes=nd, duration=3.0, lambda0=lam0, W=W, A=A, mu=1.0, tau=1.0, dtmax=1.0,
                intensities=lamB, ll=loglik(ev, nd, 3.0, lam0, W, None, lamB),
                intensities_network=lamBn, ll_network=loglik(ev, nd, 3.0, lam0, W, A, lamBn))

# KAT-C: parent of event 4 in KAT-B (parents.jl:25-46): most recent first, baseline last
t4, c4 = ev[3], nd[3]
ws = []
idx = []
for j in (2, 1, 0):
    if ev[j] > t4 - 1.0:
        ws.append(W[nd[j] - 1][c4 - 1] * pdfB(0, 0, t4 - ev[j]))
        idx.append(j + 1)
ws.append(lam0[c4 - 1])
idx.append(0)
p = np.array(ws) / np.sum(ws)


def categorical(p, u):  # Distributions rand(Categorical)
    cp, i = p[0], 0
    while cp <= u and i < len(p) - 1:
        i += 1
        cp += p[i]
    return i


out["C"] = dict(weights=ws, indices=idx, p=p.tolist(),
                draws=[dict(u=u, parent=idx[categorical(p, u)]) for u in (0.01, 0.1, 0.2, 0.999)])

# KAT-D: sufficient statistics for parents [0,1,2,3,0,5] (impulses.jl:84-96, 228-252; parents.jl:70-79; baselines.jl:87-96)
par = [0, 1, 2, 3, 0, 5]
K = 2
Mnm = np.zeros((K, K)); Xs = np.zeros((K, K)); D = np.zeros((K, K)); M0 = np.zeros(K)
for i, q in enumerate(par):
    if q > 0:
        a, b = nd[q - 1] - 1, nd[i] - 1
        d = ev[i] - ev[q - 1]
        Mnm[a, b] += 1; D[a, b] += d; Xs[a, b] += np.log(d / (1.0 - d))
    else:
        M0[nd[i] - 1] += 1
with np.errstate(invalid="ignore", divide="ignore"):
    Xbar = Xs / Mnm
    dm = np.nan_to_num(D / Mnm, nan=0.0)
V = np.zeros((K, K))
for i, q in enumerate(par):
    if q > 0:
        a, b = nd[q - 1] - 1, nd[i] - 1
        d = ev[i] - ev[q - 1]
        V[a, b] += (np.log(d / (1.0 - d)) - Xbar[a, b]) ** 2
out["D"] = dict(parents=par, Mnm=Mnm.tolist(), Xsum=Xs.tolist(), V=V.tolist(), duration_mean=dm.tolist(), M0=M0.tolist(),
                Mn=[4.0, 2.0])

# KAT-E: discrete (impulses.jl:321-335, discrete.jl:115-129, 146-151, 91-102); data of test/baselines.jl:77
L, B, dt = 4, 3, 1.0
sigma = L / (B - 1)
mus = np.linspace(1, L, B + 2)[1:-1]
lags = np.arange(1, L + 1)
phi = np.exp(-((lags[:, None] - mus[None, :]) ** 2) / (4.0 * sigma))
phi = phi / (phi.sum(axis=0, keepdims=True) * dt)
data = np.array([[0, 0, 0, 1, 0, 1, 0, 0, 0, 1], [2, 0, 0, 0, 0, 0, 0, 0, 0, 0]])
N, T = data.shape
conv = np.zeros((T, N, B))
for b in range(B):
    for n in range(N):
        full = np.convolve(data[n].astype(float), np.concatenate([[0.0], phi[:, b]]))
        conv[:, n, b] = np.maximum(full[:T], 0.0)
lam0 = np.array([1.0, 2.0]); Wd = np.array([[0.1, 0.2], [0.2, 0.1]]); theta = np.full((N, N, B), 1.0 / 3.0)
lam = np.tile(lam0 * dt, (T, 1))
for c in range(N):
    for p_ in range(N):
        for b in range(B):
            lam[:, c] += conv[:, p_, b] * Wd[p_, c] * theta[p_, c, b] * dt
ll = 0.0
for t in range(T):
    for n in range(N):
        s = data[n, t]
        ll += (s * np.log(lam[t, n]) if s > 0 else 0.0) - lam[t, n] - gammaln(s + 1.0)
out["E"] = dict(L=L, B=B, dt=dt, phi=phi.tolist(), data=data.tolist(), lambda0=lam0.tolist(), W=Wd.tolist(),
                conv=conv.tolist(), lam=lam.tolist(), ll=float(ll))

# peripheral fixtures the reference's own tests hold (test/baselines.jl:10-23, 77-78)
out["ref_tests"] = dict(
    node_counts=[dict(nodes=[1, 1, 2, 2], parentnodes=[0, 1, 0, 2], K=2, expect=[1.0, 1.0]),
                 dict(nodes=[], parentnodes=[], K=2, expect=[0.0, 0.0]),
                 dict(nodes=[1, 1, 2, 2], parentnodes=[1, 2, 1, 2], K=2, expect=[0.0, 0.0])],
    disc_suffstats=dict(data=data.tolist(), expect_Mn=[3, 2], expect_T=10))

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
    json.dump(out, f, indent=1)
print({k: (v.get("ll") if isinstance(v, dict) else None) for k, v in out.items()})
