"""Developer micro-benchmark (not the driver's bench.py): times single entry points on synthetic streams.

modes:  main | all | g                      dense / sparse / exponential sweeps at the config-2 / config-4 shapes
        one K n rate density [kind] [reps]  one shape
        cfg5 n | cfg1                       config-5 shape log-likelihood; README example latency (vs the oracle)
        grad kind K n rate [density]        log-likelihood + analytic gradient
        mcmc K n rate density nsweeps       full Gibbs sweeps, host draws vs device draws
        adj K n rate density                adjacency sampler
        disc N T B L rate                   discrete path (convolve, contraction, Gibbs, VB, adjacency)
        query K n rate nq | upload | peaks  intensity at query times; upload; roofline micro-benchmarks"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "networkhawkesprocesses.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nhp_b200 as nhp  # noqa: E402
import synth  # noqa: E402


def run(name, K, n, rate, density, kind="ln", reps=5, G=None):
    if G is not None:
        os.environ["NHP_G"] = str(G)
    else:
        os.environ.pop("NHP_G", None)
    t, nodes, T = synth.poisson_stream(n, K, rate, 1)
    if kind == "ln":
        lam0, W, mu, tau, A = synth.ln_params(K, 2, density=density)
        imp = nhp.LogitNormalImpulseResponse(mu, tau, 1.0)
    else:
        lam0, W, theta, A = synth.exp_params(K, 2, density=density)
        imp = nhp.ExponentialImpulseResponse(theta, dtmax=1.0)
    base, wts = nhp.HomogeneousProcess(lam0), nhp.DenseWeightModel(W)
    proc = nhp.ContinuousStandardHawkesProcess(base, imp, wts) if A is None else nhp.ContinuousNetworkHawkesProcess(base, imp, wts, A, nhp.BernoulliNetworkModel(density, K))
    ctx = proc._ctx()
    d = proc.upload((t, nodes, T))
    out = {}
    for op in ("loglik", "parents"):
        ms = []
        for r in range(reps + 2):
            if op == "loglik":
                nhp.loglikelihood(proc, d)
            else:
                nhp.resample_parents(proc, d, seed=1, counter=r, export=False)
            ms.append(ctx.last_kernel_ms)
        out[op] = float(np.median(ms[2:]))
    w = rate * 1.0
    print(f"{name:34s} K={K:5d} n={n:.1e} w~{w:5.0f} dens={density} G={G}  loglik {out['loglik']:8.3f} ms = {n / out['loglik'] / 1e3:9.1f} Mev/s "
          f"({n * w / out['loglik'] / 1e6:7.1f} Gpair/s) | parents {out['parents']:8.3f} ms = {n / out['parents'] / 1e3:9.1f} Mev/s", flush=True)
    d.free()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "upload":
        import ctypes
        from nhp_b200.core import _ptr
        n, K = int(float(sys.argv[2])), 1000
        t, nodes, T = synth.poisson_stream(n, K, 64.0, 1)
        ctx = nhp.default_context()
        import torch
        pt = torch.empty(n, dtype=torch.float64, pin_memory=True); pt.numpy()[:] = t
        pc = torch.empty(n, dtype=torch.int64, pin_memory=True); pc.numpy()[:] = nodes
        for name, a, b in (("pageable", _ptr(t), _ptr(nodes)), ("pinned", ctypes.c_void_p(pt.data_ptr()), ctypes.c_void_p(pc.data_ptr()))):
            for rep in range(3):
                h = ctypes.c_void_p()
                t0 = time.perf_counter()
                ctx.check(ctx.lib.nhp_events_upload(ctx.h, a, b, n, T, K, 0, 0, 1, ctypes.byref(h)))
                t1 = time.perf_counter()
                ctx.lib.nhp_events_free(ctx.h, h)
                t2 = time.perf_counter()
                print(f"upload {name} n={n:.1e}: upload {1e3*(t1-t0):.1f} ms ({n*16/(t1-t0)/1e9:.1f} GB/s), free {1e3*(t2-t1):.1f} ms", flush=True)
        lam0, W, mu, tau, A = synth.ln_params(K, 2, density=0.05)
        from nhp_b200.core import _fmat
        pl0, pW, pA, pmu, ptau = (np.ascontiguousarray(lam0), _fmat(W), _fmat(A), _fmat(mu), _fmat(tau))
        for rep in range(3):
            t0 = time.perf_counter()
            ctx.check(ctx.lib.nhp_cont_params_set(ctx.h, 1, K, _ptr(pl0), _ptr(pW), _ptr(pA), _ptr(pmu), _ptr(ptau), 1.0))
            print(f"params_set K={K}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
        sys.exit(0)
    if which == "adj":  # adj K n rate density
        K, n, rate, dens = int(sys.argv[2]), int(float(sys.argv[3])), float(sys.argv[4]), float(sys.argv[5])
        t, nodes, T = synth.poisson_stream(n, K, rate, 1)
        lam0, W, mu, tau, A = synth.ln_params(K, 2, wmax=0.5 / (K * dens), density=dens)
        proc = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W), A,
                                                  nhp.BernoulliNetworkModel(dens, K))
        d = proc.upload((t, nodes, T))
        ctx = proc._ctx()
        reset = len(sys.argv) > 6 and sys.argv[6] == "reset"  # every sweep starts from the initial matrix (all its links flip on surrogate data)
        A0 = proc.adjacency_matrix.copy()
        for rep in range(4):
            if reset:
                proc.adjacency_matrix = A0.copy()
            t0 = time.perf_counter()
            nhp.resample_adjacency_matrix_(proc, d, seed=1, counter=rep)
            info = nhp.adjacency_info(ctx)
            print(f"adjacency K={K} n={n:.1e}: kernel {ctx.last_kernel_ms:.2f} ms, call {1e3*(time.perf_counter()-t0):.1f} ms, links {int(proc.adjacency_matrix.sum())}, "
                  f"pairs {info['pairs']:.3g} ({info['pairs'] / max(ctx.last_kernel_ms, 1e-9) / 1e6:.1f} Gpair/s), batches {info['batches']:.0f}, flips {info['flips']:.0f}, "
                  f"recomputed {info['recomputed_steps']:.0f}, vcols {info['virtual_columns']:.0f}, build {info['build_ms']:.0f} ms", flush=True)
        sys.exit(0)
    if which == "disc":  # disc N T B L rate
        from nhp_b200 import discrete as D
        N, T, B, L, rate = int(sys.argv[2]), int(float(sys.argv[3])), int(sys.argv[4]), int(sys.argv[5]), float(sys.argv[6])
        rng = np.random.default_rng(3)
        lam0 = np.full(N, 0.02)
        A = (rng.random((N, N)) < 0.1).astype(np.float64)
        W = rng.uniform(0.0, 0.5, (N, N)) * A
        W *= 0.5 / max(1e-9, np.max(np.abs(np.linalg.eigvals(W))))
        theta = rng.dirichlet(np.ones(B), (N, N))
        data = rng.poisson(rate, (N, T)).astype(np.int64)
        proc = D.DiscreteNetworkHawkesProcess(D.DiscreteHomogeneousProcess(lam0), D.DiscreteGaussianImpulseResponse(theta, L), nhp.DenseWeightModel(W), A,
                                              nhp.BernoulliNetworkModel(0.1, N))
        ctx = proc._ctx()
        t0 = time.perf_counter(); d = proc.upload(data); print(f"upload {1e3*(time.perf_counter()-t0):.0f} ms, events {int(data.sum())}, nonzero bins {int((data>0).sum())}", flush=True)
        for rep in range(4):
            D.convolve(proc, d, export=False); print(f"convolve kernel {ctx.last_kernel_ms:.2f} ms ({(4*N*T + 8*T*N*B)/ctx.last_kernel_ms/1e6:.0f} GB/s)", flush=True)
        for rep in range(3):
            ll = D.loglikelihood(proc, d); ms = ctx.last_kernel_ms
            print(f"loglik (GEMM+Poisson) {ms:.2f} ms = {2*T*N*N*B/ms/1e9:.2f} TFLOP/s, {T/ms/1e3:.1f} Mbins/s  ll={ll:.6e}", flush=True)
        for rep in range(2):
            C = D.resample_parents(proc, d, seed=1, counter=rep); print(f"gibbs counts {ctx.last_kernel_ms:.2f} ms  (sum {C.sum():.0f})", flush=True)
        e0 = np.ones(N); E = rng.uniform(0.01, 0.2, (N, N, B))
        for rep in range(2):
            st = D.vb_statistics(proc, d, e0, E); print(f"vb stats {ctx.last_kernel_ms:.2f} ms", flush=True)
        for rep in range(2):
            D.resample_adjacency_matrix_(proc, d, seed=2, counter=rep); print(f"disc adjacency {ctx.last_kernel_ms:.2f} ms links {int(proc.adjacency_matrix.sum())}", flush=True)
        for rep in range(2):
            llg, _ = D.loglikelihood_gradient(proc, d); print(f"loglik + analytic gradient {ctx.last_kernel_ms:.2f} ms  ll={llg:.6e}", flush=True)
        sys.exit(0)
    if which == "cfg5":  # cfg5 n
        n, K = int(float(sys.argv[2])), 5000
        t, nodes, T = synth.poisson_stream(n, K, 3.2, 1)
        lam0, W, theta, _ = synth.exp_params(K, 2)
        proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.ExponentialImpulseResponse(theta), nhp.DenseWeightModel(W))
        ctx = proc._ctx()
        d = proc.upload((t, nodes, T))
        for rep in range(3):
            t0 = time.perf_counter(); ll = nhp.loglikelihood(proc, d, recursive=True)
            print(f"cfg5-like exp K=5000 n={n:.1e}: loglik kernel {ctx.last_kernel_ms:.2f} ms = {n/ctx.last_kernel_ms/1e3:.1f} Mev/s, call {1e3*(time.perf_counter()-t0):.0f} ms, ll={ll:.6e}", flush=True)
        sys.exit(0)
    if which == "grad":  # grad kind K n rate [density]
        kind, K, n, rate = sys.argv[2], int(sys.argv[3]), int(float(sys.argv[4])), float(sys.argv[5])
        dens = float(sys.argv[6]) if len(sys.argv) > 6 and sys.argv[6] != "none" else None
        t, nodes, T = synth.poisson_stream(n, K, rate, 1)
        if kind == "exp":
            lam0, W, theta, A = synth.exp_params(K, 2, density=dens)
            imp = nhp.ExponentialImpulseResponse(theta)
        else:
            lam0, W, mu, tau, A = synth.ln_params(K, 2, density=dens)
            imp = nhp.LogitNormalImpulseResponse(mu, tau, 1.0)
        if A is None:
            proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), imp, nhp.DenseWeightModel(W))
        else:
            proc = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(lam0), imp, nhp.DenseWeightModel(W), A, nhp.BernoulliNetworkModel(dens, K))
        ctx = proc._ctx()
        d = proc.upload((t, nodes, T))
        proc._push(ctx)
        import ctypes
        for rep in range(3):
            ll = nhp.loglikelihood(proc, d); k_ll = ctx.last_kernel_ms
            ctx.check(ctx.lib.nhp_cont_loglik_grad_dev(ctx.h, d.h, 1)); ctx.lib.nhp_cont_loglik_grad_read(ctx.h, d.h, None, None, None, None, None); k_g = ctx.last_kernel_ms
            print(f"grad {kind} K={K} n={n:.1e} dens={dens}: loglik {k_ll:.2f} ms, loglik+gradient {k_g:.2f} ms = {n/k_g/1e3:.1f} Mev/s (x{k_g/k_ll:.1f} of a loglik; finite differences: x{2*(K+(2 if kind=='exp' else 3)*K*K)} )", flush=True)
        sys.exit(0)
    if which == "mcmc":  # mcmc K n rate density nsweeps : full Gibbs sweeps, host draws vs device draws
        K, n, rate, dens, ns = int(sys.argv[2]), int(float(sys.argv[3])), float(sys.argv[4]), (None if sys.argv[5] == "none" else float(sys.argv[5])), int(sys.argv[6])
        t, nodes, T = synth.poisson_stream(n, K, rate, 1)
        lam0, W, mu, tau, A = synth.ln_params(K, 2, density=dens)
        for dev in (False, True):
            if A is None:
                proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W))
            else:
                proc = nhp.ContinuousNetworkHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W), A.copy(), nhp.BernoulliNetworkModel(dens, K))
            d = proc.upload((t, nodes, T))
            nhp.mcmc_(proc, d, nsteps=2, seed=1, device_draws=dev, store_every=10**9)
            t0 = time.perf_counter()
            nhp.mcmc_(proc, d, nsteps=ns, seed=2, device_draws=dev, store_every=10**9)
            dt = (time.perf_counter() - t0) / ns
            print(f"mcmc K={K} n={n:.1e} dens={dens} device_draws={dev}: {1e3*dt:.1f} ms per full Gibbs sweep = {n/dt/1e6:.1f} Mev/s", flush=True)
        sys.exit(0)
    if which == "cfg1":  # README example: K=2 exponential, T=1000: per-call latency of loglikelihood (params pushed each call, as mle! does)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_ffi as orc
        K = 2
        proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(np.ones(K)), nhp.ExponentialImpulseResponse(np.ones((K, K))), nhp.DenseWeightModel(0.1 * np.ones((K, K))))
        t, nodes, T = nhp.rand(proc, 1000.0, np.random.default_rng(0))
        d = proc.upload((t, nodes, T))
        for rec in (True, False):
            for _ in range(20):
                nhp.loglikelihood(proc, d, recursive=rec)
            t0 = time.perf_counter()
            for _ in range(500):
                ll = nhp.loglikelihood(proc, d, recursive=rec)
            dt = (time.perf_counter() - t0) / 500
            print(f"cfg1 K=2 n={len(t)} recursive={rec}: {1e6*dt:.0f} us per loglikelihood call (kernel {1e3*proc._ctx().last_kernel_ms:.0f} us), ll={ll:.6f}", flush=True)
        om = orc.Cont(0, np.ones(K), 0.1 * np.ones((K, K)), np.ones((K, K)))
        for rec in (True, False):
            t0 = time.perf_counter()
            for _ in range(50):
                llo = om.loglik(t, nodes, T, recursive=rec)
            print(f"oracle (1 thread) recursive={rec}: {1e6*(time.perf_counter()-t0)/50:.0f} us per call, ll={llo:.6f}", flush=True)
        sys.exit(0)
    if which == "query":  # query K n rate nq
        K, n, rate, nq = int(sys.argv[2]), int(float(sys.argv[3])), float(sys.argv[4]), int(sys.argv[5])
        t, nodes, T = synth.poisson_stream(n, K, rate, 1)
        lam0, W, mu, tau, A = synth.ln_params(K, 2)
        proc = nhp.ContinuousStandardHawkesProcess(nhp.HomogeneousProcess(lam0), nhp.LogitNormalImpulseResponse(mu, tau, 1.0), nhp.DenseWeightModel(W))
        d = proc.upload((t, nodes, T))
        tq = np.linspace(0.0, T, nq)
        for rep in range(3):
            t0 = time.perf_counter(); lam = nhp.intensity(proc, d, tq)
            print(f"intensity(process, data, times) K={K} nq={nq} w~{rate:.0f}: kernel {proc._ctx().last_kernel_ms:.2f} ms, call {1e3*(time.perf_counter()-t0):.1f} ms, {nq*rate*K/proc._ctx().last_kernel_ms/1e6:.1f} Gpair/s", flush=True)
        sys.exit(0)
    if which == "peaks":
        import ctypes
        ctx = nhp.default_context()
        for w, name in ((0, "DFMA TFLOP/s"), (3, "DMMA m8n8k4 TFLOP/s"), (1, "LN pairs/s"), (2, "EXP pairs/s")):
            r = ctypes.c_double(); ctx.check(ctx.lib.nhp_bench_fp64(ctx.h, w, ctypes.byref(r))); print(name, r.value, flush=True)
        sys.exit(0)
    if which == "one":  # one K n rate density [kind] [reps]
        K, n, rate = int(sys.argv[2]), int(float(sys.argv[3])), float(sys.argv[4])
        dens = None if sys.argv[5] == "none" else float(sys.argv[5])
        run("one", K, n, rate, dens, kind=sys.argv[6] if len(sys.argv) > 6 else "ln", reps=int(sys.argv[7]) if len(sys.argv) > 7 else 2)
    if which in ("all", "g"):
        for G in (1, 2, 4, 8, 16, 32):
            run("cfg2 LN std K=50", 50, 1_000_000, 100.0, None, G=G)
        for G in (1, 4, 8, 16, 32):
            run("cfg4-like LN std K=1000", 1000, 4_000_000, 64.0, None, G=G)
    if which in ("all", "main"):
        run("cfg2 LN std K=50", 50, 1_000_000, 100.0, None)
        run("cfg4-like LN std K=1000 dense", 1000, 10_000_000, 64.0, None)
        run("cfg4-like LN net K=1000 rho=.05", 1000, 10_000_000, 64.0, 0.05)
        run("exp std K=50 dtmax=1", 50, 1_000_000, 100.0, None, kind="exp")
