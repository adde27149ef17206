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

# ------------------------------------------------------------------------------------------------------------------------
# Round-2 vectors for the rows that had none (a5, a9, Q3, a12, a13, a14): literal numpy restatements of the cited Julia,
# written without looking at oracle/.
# ------------------------------------------------------------------------------------------------------------------------
# KAT-F: intensity(process, data, times) (continuous.jl:76-96): window time - dtmax < t_j < time, STRICT on both sides, so
# a query at an event's own time excludes that event, and an event exactly dtmax back is excluded as well
evF, ndF = out["B"]["events"], out["B"]["nodes"]
lam0F, WF, AF = out["B"]["lambda0"], out["B"]["W"], out["B"]["A"]
timesF = [0.0, 0.05, 0.4, 0.9, 1.35, 1.4, 2.55, 2.6, 3.0]  # 0.4 / 0.9 / 2.6 are event times; 1.4 - 1.0 = 0.4 sits on the window edge


def intensity_at(time, A):
    lam = [0.0, 0.0]
    for c in range(2):
        for tj, p in zip(evF, ndF):
            if time - 1.0 < tj < time:
                a = 1.0 if A is None else A[p - 1][c]
                lam[c] += a * WF[p - 1][c] * ln_pdf(1.0, 1.0, (time - tj) / 1.0)
    return [lam0F[c] + lam[c] for c in range(2)]


out["F"] = dict(times=timesF, standard=[intensity_at(t, None) for t in timesF], network=[intensity_at(t, AF) for t in timesF])

# KAT-G: resample_adjacency_matrix! on KAT-B (continuous.jl:444-519), uniforms given: A[p,c] = (u[p,c] <= exp(ll1 - logsumexp(ll0, ll1))),
# columns independent, p sequential, sum_log_intensity skipping the log term of event 1 (quirk Q4)
from scipy.special import logsumexp


def sum_log_intensity(node, A):  # continuous.jl:500-519
    S = 0.0
    for index in range(1, len(evF) + 1):
        if ndF[index - 1] != node:
            continue
        lam = lam0F[node - 1]
        if index == 1:
            continue
        pi = index - 1
        while evF[pi - 1] > evF[index - 1] - 1.0:
            pn = ndF[pi - 1]
            lam += A[pn - 1][node - 1] * WF[pn - 1][node - 1] * ln_pdf(1.0, 1.0, evF[index - 1] - evF[pi - 1])
            pi -= 1
            if pi == 0:
                break
        S += np.log(lam)
    return S


def integrated(node, A, counts, T):  # continuous.jl:489-498
    return lam0F[node - 1] * T + sum(A[p][node - 1] * WF[p][node - 1] * counts[p] for p in range(2))


def adjacency_sweep(A0, rho, u, T=3.0):
    A = [row[:] for row in A0]
    counts = [sum(1 for x in ndF if x == k + 1) for k in range(2)]
    probs = [[0.0, 0.0], [0.0, 0.0]]
    for node in (1, 2):
        for p in (1, 2):
            A[p - 1][node - 1] = 0.0
            ll0 = -integrated(node, A, counts, T) + sum_log_intensity(node, A) + np.log(1.0 - rho)
            A[p - 1][node - 1] = 1.0
            ll1 = -integrated(node, A, counts, T) + sum_log_intensity(node, A) + np.log(rho)
            pr = float(np.exp(ll1 - logsumexp([ll0, ll1])))
            probs[p - 1][node - 1] = pr
            A[p - 1][node - 1] = 1.0 if u[p - 1][node - 1] <= pr else 0.0
    return A, probs


casesG = []
for rho, u in ((0.5, [[0.5, 0.5], [0.5, 0.5]]), (0.3, [[0.1, 0.9], [0.35, 0.2]]), (0.7, [[0.95, 0.05], [0.6, 0.65]])):
    Aout, probs = adjacency_sweep(AF, rho, u)
    casesG.append(dict(rho=rho, u=u, A=Aout, p1=probs))
out["G"] = dict(A0=AF, cases=casesG)

# KAT-H: recursive_loglikelihood of the NETWORK process (continuous.jl:407-442): Exponential impulse, full history; the integral
# term sums W WITHOUT the adjacency matrix (quirk Q3), the intensities use effective weights A .* W
evH, ndH, TH = [0.5, 1.0, 1.5, 3.0, 3.2], [1, 2, 1, 2, 2], 4.0
lam0H, WH, AH, thH = [1.0, 0.5], [[0.1, 0.3], [0.2, 0.4]], [[1.0, 0.0], [1.0, 1.0]], [[1.0, 2.0], [0.5, 1.5]]
llH = -sum(l * TH for l in lam0H)
for p in ndH:
    llH -= sum(WH[p - 1])
lamH = []
for i, (t, c) in enumerate(zip(evH, ndH)):
    v = lam0H[c - 1]
    for j in range(i):
        p = ndH[j]
        v += AH[p - 1][c - 1] * WH[p - 1][c - 1] * thH[p - 1][c - 1] * np.exp(-thH[p - 1][c - 1] * (t - evH[j]))
    lamH.append(v)
llH += sum(np.log(lamH))
llH_windowed = -sum(l * TH for l in lam0H) - sum(sum(AH[p - 1][c] * WH[p - 1][c] for c in range(2)) for p in ndH) + sum(np.log(lamH))
out["H"] = dict(events=evH, nodes=ndH, duration=TH, lambda0=lam0H, W=WH, A=AH, theta=thH, intensities=lamH, ll_recursive=float(llH),
                ll_windowed=float(llH_windowed))

# KAT-I: discrete Gibbs parents on KAT-E (parents.jl:82-117): per (t, c) Multinomial(data[c,t], mu), mu = [lambda0_c dt; vec(lambda[b,p])] / sum,
# realised as data[c,t] categorical inverse-cdf draws (Distributions' rand(Categorical): first k with cp > u) consuming one
# uniform each in (t outer, c inner, draw) order; only sum_t parents[t,c,:] is used downstream (parents.jl:124-134)
data1, conv1, T1 = data, conv, T
data = np.array([[1, 0, 2, 1, 0, 0, 3, 1, 0, 1, 2, 0], [0, 1, 1, 0, 2, 0, 1, 0, 1, 1, 0, 2]])  # a busier 2 x 12 matrix for KAT-I/J/K
N, T = data.shape
conv = np.zeros((T, N, B))
for b in range(B):
    for n in range(N):
        full = np.convolve(data[n].astype(float), np.concatenate([[0.0], phi[:, b]]))
        conv[:, n, b] = np.maximum(full[:T], 0.0)
thetaI = np.array([[[0.2, 0.3, 0.5], [0.6, 0.3, 0.1]], [[1 / 3, 1 / 3, 1 / 3], [0.1, 0.1, 0.8]]])  # [p, c, b]
WI = np.array([[0.3, 0.6], [0.9, 0.2]])
lam0I = np.array([0.4, 0.7])
uI = [0.05, 0.62, 0.97, 0.33, 0.81, 0.18, 0.44, 0.71, 0.09, 0.56, 0.93, 0.27, 0.38, 0.66, 0.02, 0.85, 0.49, 0.74, 0.13, 0.91, 0.58, 0.30]
countsI = np.zeros((N, 1 + N * B))
muI = []
iu = 0
for t in range(T):
    for c in range(N):
        s = int(data[c, t])
        if s == 0:
            continue
        mu = [lam0I[c] * dt] + [conv[t, p_, b] * WI[p_, c] * thetaI[p_, c, b] * dt for p_ in range(N) for b in range(B)]
        mu = np.array(mu) / np.sum(mu)
        muI.append(dict(t=t + 1, c=c + 1, mu=mu.tolist()))
        for _ in range(s):
            countsI[c, categorical(mu, uI[iu])] += 1
            iu += 1
assert iu == int(data.sum())
out["I"] = dict(data=data.tolist(), conv=conv.tolist(), lambda0=lam0I.tolist(), W=WI.tolist(), theta=thetaI.tolist(), u=uI[:iu], counts=countsI.tolist(), mu=muI)

# KAT-J: VB statistics on KAT-E (parents.jl:136-177 update_parents; baselines.jl:444-452, weights.jl:70-91, impulses.jl:355-371):
# u[t,c,:] = [e0_c; conv[t,p,b] E[p,c,b]] / Z;  alpha_sum_c = sum_t u[t,c,1] data[c,t];  gamma_sum[p,c,b] = sum_t data[c,t] u[t,c,1+(p-1)B+b];
# kappa_sum[p,c] = sum_b gamma_sum[p,c,b];  nu_sum[p,c] = sum_t data[p,t]
e0J = np.array([0.8, 1.7])
EJ = np.array([[[0.11, 0.23, 0.05], [0.31, 0.07, 0.19]], [[0.13, 0.29, 0.17], [0.02, 0.37, 0.41]]])  # [p, c, b]
aJ, gJ = np.zeros(N), np.zeros((N, N, B))
for t in range(T):
    for c in range(N):
        uu = np.array([e0J[c]] + [conv[t, p_, b] * EJ[p_, c, b] for p_ in range(N) for b in range(B)])
        uu = uu / uu.sum()
        aJ[c] += uu[0] * data[c, t]
        for p_ in range(N):
            for b in range(B):
                gJ[p_, c, b] += data[c, t] * uu[1 + p_ * B + b]
out["J"] = dict(e0=e0J.tolist(), E=EJ.tolist(), alpha_sum=aJ.tolist(), gamma_sum=gJ.tolist(), kappa_sum=gJ.sum(axis=2).tolist(),
                nu_sum=np.tile(data.sum(axis=1)[:, None].astype(float), (1, N)).tolist())

# KAT-K: discrete adjacency sweep on KAT-E (discrete.jl:426-480), uniforms given; conditional_loglikelihood recomputes the whole column
AK0 = np.array([[1.0, 0.0], [1.0, 1.0]])


def cond_ll(A, value, pidx, cidx):  # discrete.jl:462-480
    ll = 0.0
    for t in range(T):
        lam_ = lam0I[cidx] * dt
        for p_ in range(N):
            a = value if p_ == pidx else A[p_, cidx]
            for b in range(B):
                lam_ += conv[t, p_, b] * a * WI[p_, cidx] * thetaI[p_, cidx, b] * dt
        ll += stats.poisson(lam_).logpmf(data[cidx, t])
    return ll


casesK = []
for rho, u in ((0.5, [[0.5, 0.5], [0.5, 0.5]]), (0.2, [[0.05, 0.4], [0.9, 0.1]])):
    A = AK0.copy()
    pr = np.zeros((N, N))
    for c in range(N):
        for p_ in range(N):
            l0 = cond_ll(A, 0.0, p_, c) + np.log(1 - rho)
            l1 = cond_ll(A, 1.0, p_, c) + np.log(rho)
            pr[p_, c] = np.exp(l1 - logsumexp([l0, l1]))
            A[p_, c] = 1.0 if u[p_][c] <= pr[p_, c] else 0.0
    casesK.append(dict(rho=rho, u=u, A=A.tolist(), p1=pr.tolist()))
out["K"] = dict(A0=AK0.tolist(), cases=casesK)
data, conv, T = data1, conv1, T1

# peripheral fixtures the reference's own tests hold (test/baselines.jl:10-23, 77-78)
out["ref_tests"] = dict(
    node_counts=[dict(nodes=[1, 1, 2, 2], parentnodes=[0, 1, 0, 2], K=2, expect=[1.0, 1.0]),
                 dict(nodes=[], parentnodes=[], K=2, expect=[0.0, 0.0]),
                 dict(nodes=[1, 1, 2, 2], parentnodes=[1, 2, 1, 2], K=2, expect=[0.0, 0.0])],
    disc_suffstats=dict(data=data.tolist(), expect_Mn=[3, 2], expect_T=10))

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
    json.dump(out, f, indent=1)
print({k: (v.get("ll") if isinstance(v, dict) else None) for k, v in out.items()})
