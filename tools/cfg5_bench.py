#!/usr/bin/env python3
"""BASELINE.json configs[4]: continuous Exponential standard Hawkes, K = 5000 nodes, 1e9 events, log-likelihood + `mle!`
gradient sweep, time-sharded across the GPUs of one box (one process per GPU, Δt_cut halo, NCCL all-reduce of the
gradient planes).  Not the driver's bench line (bench.py measures config 4); run it as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/cfg5_bench.py --events 1e9 --steps 3

`--events` is the GLOBAL event count (each rank takes events / world).  Prints one JSON line on rank 0.  The stream is the
Poisson surrogate of SURVEY.md section 8c (inter-arrivals Exp(rate), nodes uniform), parameters W ~ U(0, 1/K),
theta ~ U(0.5, 2), lambda0 ~ U(0.5, 1.5); `recursive=true` semantics (full history through the cut-off horizon)."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "networkhawkesprocesses.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events", type=float, default=1e9)
    ap.add_argument("--nodes", type=int, default=5000)
    ap.add_argument("--rate", type=float, default=3.2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import nhp_b200 as nhp
    import synth
    from nhp_b200.core import _fmat, _ptr

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = nhp.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    lib = ctx.lib
    ctx.check(lib.nhp_set_stream(ctx.h, ctypes.c_void_p(stream.cuda_stream)))
    K, n = args.nodes, int(args.events) // world
    lam0, W, theta, _ = synth.exp_params(K, 2)
    pl0, pW, pth = np.ascontiguousarray(lam0), _fmat(W), _fmat(theta)
    ctx.check(lib.nhp_cont_params_set(ctx.h, 0, K, _ptr(pl0), _ptr(pW), None, _ptr(pth), None, float("inf")))
    hz = ctypes.c_double()
    ctx.check(lib.nhp_cont_horizon(ctx.h, int(args.events), 1, ctypes.byref(hz)))

    rng = np.random.Generator(np.random.Philox(key=0x4E485035 + rank))
    gaps = rng.exponential(1.0 / args.rate, n)
    nodes = rng.integers(1, K + 1, n, dtype=np.int64)
    span = float(gaps.sum())
    halo_n = int(args.rate * hz.value * 1.5) + 64
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
    del t_local, nodes
    h = ctypes.c_void_p()
    t0 = time.perf_counter()
    ctx.check(lib.nhp_events_upload(ctx.h, _ptr(t_all), _ptr(c_all), n_halo + n, duration, K, n_halo, max(rank * n - n_halo, 0), 1 if rank == 0 else 0, ctypes.byref(h)))
    upload_s = time.perf_counter() - t0
    del t_all, c_all

    def stats_tensor(phase):
        p, cnt = ctypes.c_void_p(), ctypes.c_int64()
        ctx.check(lib.nhp_cont_stats_dev(ctx.h, phase, ctypes.byref(p), ctypes.byref(cnt)))

        class _Raw:
            __cuda_array_interface__ = {"shape": (cnt.value,), "typestr": "<f8", "data": (p.value, False), "version": 3}
        return torch.as_tensor(_Raw(), device=dev)

    st0 = stats_tensor(0)
    ms = {"loglik": [], "grad": []}

    def step(record):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record(stream)
        ctx.check(lib.nhp_cont_loglik_dev(ctx.h, h, 1))
        ev[1].record(stream)
        ctx.check(lib.nhp_cont_loglik_grad_dev(ctx.h, h, 1))
        ev[2].record(stream)
        if world > 1:
            dist.all_reduce(st0)  # [ll terms, dlambda0, Mn, dW, dtheta]: 2 + 2K + 2K^2 doubles
        if record:
            ev[2].synchronize()
            ms["loglik"].append(ev[0].elapsed_time(ev[1]))
            ms["grad"].append(ev[1].elapsed_time(ev[2]))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    e1.record(stream)
    sync_all()
    tot = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ll = ctypes.c_double()
    g0 = np.empty(K)
    ctx.check(lib.nhp_cont_loglik_grad_read(ctx.h, h, ctypes.byref(ll), _ptr(g0), None, None, None))
    if rank == 0:
        ms_step = float(tot.item()) / args.steps
        print(json.dumps({"metric": "loglik+gradient_sweep_events_per_s", "value": world * n / (ms_step * 1e-3), "unit": "events/s", "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "scaling": "strong (global event count fixed by --events)",
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": "cfg5: continuous Exponential standard Hawkes, K=%d, %.3g events total (%d per GPU), rate %.1f/s, recursive "
                                                 "(cut-off horizon %.1f s ~ %d pairs/event); step = loglikelihood + loglikelihood-with-analytic-gradient + "
                                                 "NCCL all-reduce of the %d-double gradient buffer" % (K, world * n, n, args.rate, hz.value, int(hz.value * args.rate), st0.numel()),
                                     "halo_events": int(n_halo), "upload_s": upload_s},
                          "kernel_ms_rank0": {k: float(np.median(v)) for k, v in ms.items()},
                          "ll_share_rank0": ll.value, "dlambda0_head": g0[:3].tolist()}), flush=True)
    lib.nhp_events_free(ctx.h, h)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
