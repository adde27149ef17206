"""`python bench.py --config {2,3,5}`: the other BASELINE.json configurations as their own JSON line, for 1..N GPUs
(one process per GPU under torchrun).  The collectives are the library's (nhp_comm_*: NCCL over NVLink inside libnhp).

  --config 2   continuous LogitNormal standard, K = 50, 1e6 events: log-likelihood + Gibbs parent sweep with fused statistics
               (time shards + dtmax halo; all-reduce of the statistics)
  --config 3   discrete Gaussian network, N = 200, T = 1e6 bins, B = 6, L = 12: convolve (once), log-likelihood contraction,
               Gibbs counts, VB statistics (time shards + L-bin halo; all-reduce of the counts / statistics)
  --config 5   continuous Exponential standard, K = 5000, recursive semantics: log-likelihood + analytic gradient
               (time shards + cut-off halo; all-reduce of the gradient planes); 1.25e8 events per GPU unless --events is given
Also `--query`-style sharding of intensity(process, data, times) is timed inside --config 2 (query times are
partitioned over the ranks, no collective).
"""
import ctypes
import time

import numpy as np

RATES = {2: 100.0, 5: 3.2}


def run(args, rank, world, local_rank, emit, setup_distributed, make_context, ClockSampler, hbm_peak):
    torch, dist, dev = setup_distributed(local_rank, world)
    ctx, stream = make_context(torch, dev, local_rank, rank, world, dist)
    if args.config == 3:
        line = run_cfg3(args, rank, world, ctx, torch, dist, dev, stream)
    else:
        line = run_cont(args, rank, world, ctx, torch, dist, dev, stream)
    if rank == 0:
        line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"})
        emit(line)
    if world > 1:
        ctx.lib.nhp_comm_destroy(ctx.h)
        dist.destroy_process_group()


def timed_steps(args, torch, dist, dev, stream, world, step):
    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    tot = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    return float(tot.item()) / args.steps


def run_cont(args, rank, world, ctx, torch, dist, dev, stream):
    import synth
    from nhp_b200.core import _fmat, _ptr
    lib = ctx.lib
    cfg = args.config
    K = 50 if cfg == 2 else 5000
    rate = RATES[cfg]
    if cfg == 2:
        n_total = int(args.events) if args.events_given else 1_000_000
        scaling = "strong"
    else:
        n_total = int(args.events) if args.events_given else 125_000_000 * world
        scaling = "strong" if args.events_given else "weak"
    n = n_total // world
    if cfg == 2:
        lam0, W, mu, tau, _ = synth.ln_params(K, 2)
        ctx.check(lib.nhp_cont_params_set(ctx.h, 1, K, _ptr(np.ascontiguousarray(lam0)), _ptr(_fmat(W)), None, _ptr(_fmat(mu)), _ptr(_fmat(tau)), 1.0))
        recursive = 0
    else:
        lam0, W, theta, _ = synth.exp_params(K, 2)
        ctx.check(lib.nhp_cont_params_set(ctx.h, 0, K, _ptr(np.ascontiguousarray(lam0)), _ptr(_fmat(W)), None, _ptr(_fmat(theta)), None, float("inf")))
        recursive = 1
    hz = ctypes.c_double()
    ctx.check(lib.nhp_cont_horizon(ctx.h, n_total, recursive, ctypes.byref(hz)))
    # this rank's shard of one global Poisson-surrogate stream: gaps seeded per rank, absolute times fixed up through an all-gather
    rng = np.random.Generator(np.random.Philox(key=0x4E485030 + 16 * cfg + rank))
    gaps = rng.exponential(1.0 / rate, n)
    nodes = rng.integers(1, K + 1, n, dtype=np.int64)
    span = float(gaps.sum())
    halo_n = int(rate * hz.value * 1.5) + 64
    if world > 1:
        spans = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(spans, torch.tensor([span], dtype=torch.float64, device=dev))
        start = float(sum(s.item() for s in spans[:rank]))
        t_local = start + np.cumsum(gaps)
        tails_t = [torch.zeros(halo_n, dtype=torch.float64, device=dev) for _ in range(world)]
        tails_c = [torch.zeros(halo_n, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(tails_t, torch.from_numpy(t_local[-halo_n:]).to(dev))
        dist.all_gather(tails_c, torch.from_numpy(nodes[-halo_n:]).to(dev))
        duration = float(sum(s.item() for s in spans)) * (1 + 1e-9)
    else:
        t_local, duration = np.cumsum(gaps), span * (1 + 1e-9)
    del gaps
    if rank > 0:
        ht, hc = tails_t[rank - 1].cpu().numpy(), tails_c[rank - 1].cpu().numpy()
        keep = ht > t_local[0] - hz.value
        ht, hc = ht[keep], hc[keep]
    else:
        ht, hc = np.zeros(0), np.zeros(0, np.int64)
    n_halo = ht.size
    t_all, c_all = np.concatenate([ht, t_local]), np.concatenate([hc, nodes])
    h = ctypes.c_void_p()
    ctx.check(lib.nhp_events_upload(ctx.h, _ptr(t_all), _ptr(c_all), n_halo + n, duration, K, n_halo, max(rank * n - n_halo, 0), 1 if rank == 0 else 0, ctypes.byref(h)))
    ms = {}
    cnt = [0]
    ll_tot = [0.0]

    def rec(key):
        ms.setdefault(key, []).append(ctx.last_kernel_ms)

    if cfg == 2:
        def step():
            cnt[0] += 1
            ll = ctypes.c_double()
            ctx.check(lib.nhp_cont_loglik(ctx.h, h, 0, ctypes.byref(ll))); rec("loglik")
            v = np.array([ll.value]); ctx.check(lib.nhp_comm_allreduce_host(ctx.h, _ptr(v), 1)); ll_tot[0] = float(v[0])
            ctx.check(lib.nhp_cont_resample_parents(ctx.h, h, 7, cnt[0], None, None, None)); rec("parents")
            ctx.check(lib.nhp_comm_allreduce_stats(ctx.h, 0))
            ctx.check(lib.nhp_cont_suffstats_second_pass(ctx.h, h))
            ctx.check(lib.nhp_comm_allreduce_stats(ctx.h, 1))
        what = "loglikelihood + Gibbs parent sweep with fused statistics (+ second pass, + the two all-reduces)"
        metric = "loglik+parent_sweep_events_per_s"
    else:
        def step():
            ctx.check(lib.nhp_cont_loglik_dev(ctx.h, h, 1))
            ctx.check(lib.nhp_cont_loglik_grad_dev(ctx.h, h, 1))
            ctx.check(lib.nhp_comm_allreduce_stats(ctx.h, 0))  # [ll terms, dlambda0, Mn, dW, dtheta]
        what = "loglikelihood + loglikelihood-with-analytic-gradient + all-reduce of the gradient planes (2 + 2K + 2K^2 doubles)"
        metric = "loglik+gradient_sweep_events_per_s"
    ms_step = timed_steps(args, torch, dist, dev, stream, world, step)
    line = {"metric": metric, "value": n_total / (ms_step * 1e-3), "unit": "events/s", "ms_per_step": ms_step, "scaling": scaling,
            "config": {"workload": "cfg%d: K=%d, %.3g events total (%d per GPU), rate %.1f/s, look-back horizon %.2f (~%d pairs/event); step = %s"
                                   % (cfg, K, n_total, n, rate, hz.value, int(hz.value * rate), what), "halo_events": int(n_halo)},
            "kernel_ms_rank0": {k: float(np.median(v)) for k, v in ms.items()}}
    if cfg == 2:
        # intensity(process, data, times) with the query times partitioned over the ranks (events of the shard + halo
        # only serve the queries that fall inside the shard); no collective, the host gathers the rows
        nq = 10001
        tq_all = np.linspace(0.0, duration / (1 + 1e-9), nq)
        lo_t = t_all[n_halo] if rank > 0 else 0.0
        hi_t = t_all[-1] if rank < world - 1 else np.inf
        mine = tq_all[(tq_all >= lo_t) & (tq_all < hi_t)] if world > 1 else tq_all
        out = np.empty(mine.size * K)
        tq0 = time.perf_counter()
        if mine.size:
            ctx.check(lib.nhp_cont_intensity(ctx.h, h, _ptr(np.ascontiguousarray(mine)), mine.size, _ptr(out)))
        qs = torch.tensor([time.perf_counter() - tq0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(qs, op=dist.ReduceOp.MAX)
        line["query_sharding"] = {"queries_total": nq, "queries_rank0": int(mine.size), "seconds_max_over_ranks": float(qs.item()),
                                  "queries_per_s": nq / float(qs.item())}
        line["log_likelihood"] = ll_tot[0]
    lib.nhp_events_free(ctx.h, h)
    return line


def run_cfg3(args, rank, world, ctx, torch, dist, dev, stream):
    import nhp_b200 as nhp
    from nhp_b200 import discrete as D
    from nhp_b200.core import _ptr
    lib = ctx.lib
    N, B, L = 200, 6, 12
    T = int(args.events) if args.events_given else 1_000_000
    rng = np.random.default_rng(3)
    lam0 = np.full(N, 0.02)
    A = (rng.random((N, N)) < 0.1).astype(np.float64)
    W = rng.uniform(0.0, 0.5, (N, N)) * A
    W *= 0.5 / max(1e-9, np.max(np.abs(np.linalg.eigvals(W))))
    theta = rng.dirichlet(np.ones(B), (N, N))
    t0, t1 = rank * T // world, (rank + 1) * T // world
    halo = min(L, t0)
    rs = np.random.default_rng(100)  # the same global count matrix on every rank (cheap at this size), sliced per shard
    data = rs.poisson(0.04, (N, T)).astype(np.int64)[:, t0 - halo:t1]
    proc = D.DiscreteNetworkHawkesProcess(D.DiscreteHomogeneousProcess(lam0), D.DiscreteGaussianImpulseResponse(theta, L), nhp.DenseWeightModel(W), A,
                                          nhp.BernoulliNetworkModel(0.1, N))
    proc.ctx = ctx
    d = D.DiscreteData(ctx, data, t_halo=halo)
    D.convolve(proc, d, export=False)
    conv_ms = ctx.last_kernel_ms
    e0, E = np.ones(N), rng.uniform(0.01, 0.2, (N, N, B))
    ms = {}
    cnt = [0]
    res = {}

    def step():
        cnt[0] += 1
        ll = np.array([D.loglikelihood(proc, d)]); ms.setdefault("loglik_contraction", []).append(ctx.last_kernel_ms)
        ctx.check(lib.nhp_comm_allreduce_host(ctx.h, _ptr(ll), 1))
        C = np.ascontiguousarray(D.resample_parents(proc, d, seed=1, counter=cnt[0])); ms.setdefault("gibbs_counts", []).append(ctx.last_kernel_ms)
        ctx.check(lib.nhp_comm_allreduce_host(ctx.h, _ptr(C), C.size))
        st = D.vb_statistics(proc, d, e0, E); ms.setdefault("vb_stats", []).append(ctx.last_kernel_ms)
        g = np.ascontiguousarray(st["gamma_sum"])
        ctx.check(lib.nhp_comm_allreduce_host(ctx.h, _ptr(g), g.size))
        res["ll"], res["events"] = float(ll[0]), float(C.sum())
    ms_step = timed_steps(args, torch, dist, dev, stream, world, step)
    return {"metric": "loglik+gibbs+vb_bins_per_s", "value": T / (ms_step * 1e-3), "unit": "bins/s", "ms_per_step": ms_step, "scaling": "strong",
            "config": {"workload": "cfg3: discrete Gaussian network Hawkes, N=200, T=%d bins (%d per GPU + %d halo bins), B=6, L=12; step = loglikelihood contraction + Gibbs "
                                   "counts + VB statistics, each all-reduced" % (T, t1 - t0, halo)},
            "kernel_ms_rank0": {k: float(np.median(v)) for k, v in ms.items()}, "convolve_ms_rank0": conv_ms,
            "log_likelihood": res.get("ll"), "events_attributed": res.get("events")}
