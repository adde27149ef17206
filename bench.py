#!/usr/bin/env python
"""bench.py -- headline benchmark of the event-history hot path (BASELINE.json metric:
"loglik & Gibbs-sweep events/s at 1/2/4/8 B200, % roofline, vs CPU Threads").

Workload (BASELINE.json configs[3], the configuration the north_star target is quoted on; it fits one
GPU): continuous LogitNormal *network* Hawkes process, Bernoulli(rho=0.05) adjacency, K=1000 nodes,
1e8 events PER GPU (weak scaling: contiguous time shards of one global stream, each with its dtmax
halo), total rate 64 events/s, dtmax = 1 (mean predecessor window 64).  Synthetic Poisson-surrogate
stream.

A step = one log-likelihood evaluation + one Gibbs sweep (parent resampling fused with the
sufficient statistics, the two-pass variance statistic, and -- N > 1 -- the NCCL allreduce of the
statistics) over every event of the shard.

  python bench.py [--gpus N] [--steps K] [--warmup W]           (torchrun launches N > 1)
  python bench.py --impl reference ...                          (CPU oracle arm, all host threads)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "networkhawkesprocesses.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "loglik+gibbs_sweep_events_per_s"
UNIT = "events/s"
RATE, DTMAX, RHO = 64.0, 1.0, 0.05
FLOPS_PER_LN_PAIR, FLOPS_PER_EVENT = 96.0, 40.0  # SURVEY.md section 8d (nominal FP64 flops, libdevice-class accuracy)
BYTES_LOGLIK, BYTES_PARENTS = 14.0, 18.0        # per event: 8 B time + 4 B node + 2 B cached window length (+ 4 B parent offset written)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--events", type=float, default=1e8, help="events per GPU")
    ap.add_argument("--nodes", type=int, default=1000)
    ap.add_argument("--cpu-sample", type=float, default=2e6, help="events of the CPU-baseline sample")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --events per GPU (default, the driver's contract); strong: --events in total, split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--adjacency", action="store_true", help="also time one adjacency Gibbs sweep (reported in detail, not in value)")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "cfg4: continuous LogitNormal network Hawkes, Bernoulli(rho=0.05) adjacency, K=%d, %.0e events per GPU, "
                        "rate 64/s, dtmax=1 (mean window 64), step = loglikelihood + Gibbs sweep (parents + fused statistics)" % (args.nodes, args.events),
            "K": args.nodes, "events_per_gpu": int(args.events) // (world if args.scaling == "strong" else 1),
            "global_events": int(args.events) * (1 if args.scaling == "strong" else world), "mean_window": RATE * DTMAX,
            "rho": RHO, "sharding": "contiguous time shards + dtmax halo" if world > 1 else "single shard",
            "l2_policy": "inputs (1.2 GB of events per GPU) are larger than the 126 MB L2; no explicit flush"}


def make_params(K, seed=2):
    import synth
    lam0, W, mu, tau, A = synth.ln_params(K, seed, wmax=0.5 / (K * RHO), density=RHO)
    return lam0, W, mu, tau, A


def make_shard(n, K, rank):
    """Inter-arrival gaps of this rank's shard (seeded by rank); absolute times are fixed up by the caller."""
    rng = np.random.default_rng(1000 + rank)
    gaps = rng.exponential(1.0 / RATE, n)
    nodes = rng.integers(1, K + 1, n, dtype=np.int64)
    return gaps, nodes


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons DURING the timed region (NVML; nvidia-smi as fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.max_mhz = index, [], False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop_flag:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((mhz, [k for k, b in bits.items() if r & b]))
                time.sleep(0.02)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.max_mhz = float(f[1])
                self.rows.append((float(f[0]), [n for i, n in enumerate(names) if f[2 + i].lower().startswith("active")]))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        reasons = sorted({r for _, rs in self.rows for r in rs})
        return {"sm_mhz": float(np.median([m for m, _ in self.rows])), "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference, C + OpenMP at the reference's Threads.@threads sites)
# ------------------------------------------------------------------------------------------
def cpu_step_time(args, n_sample, reps):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_ffi as orc
    K = args.nodes
    lam0, W, mu, tau, A = make_params(K)
    gaps, nodes = make_shard(n_sample, K, 0)
    t = np.cumsum(gaps)
    T = float(t[-1])
    cores = orc.max_threads()
    orc.set_threads(cores)
    om = orc.Cont(1, lam0, W, mu, tau, A=A, dtmax=DTMAX)
    u = np.random.default_rng(5).random(n_sample)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        om.loglik(t, nodes, T)
        par, pn = om.resample_parents(t, nodes, u)
        orc.suffstats(1, t, nodes, par, pn, K, DTMAX)
        times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_sample = int(args.cpu_sample)
    times, cores = cpu_step_time(args, n_sample, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = float(np.sum(timed))
    value = n_sample * len(timed) / total
    sample = "first %d events of the rank-0 shard per step (oracle/liboracle.so: loglik + resample_parents + sufficient statistics, OpenMP over %d threads at the reference's Threads.@threads sites)" % (n_sample, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import ctypes
    import torch
    import torch.distributed as dist
    import nhp_b200 as nhp
    from nhp_b200.core import _fmat, _ptr

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = nhp.Context(local_rank)
    # one explicit (non-default) stream shared by libnhp's kernels, the NCCL collectives and the timing events
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.check(ctx.lib.nhp_set_stream(ctx.h, ctypes.c_void_p(stream.cuda_stream)))
    lib = ctx.lib
    K, n = args.nodes, int(args.events) // (world if args.scaling == "strong" else 1)

    # ---- synthetic shard (pinned host buffers so the e2e leg copies at PCIe speed)
    lam0, W, mu, tau, A = make_params(K)
    gaps, nodes_np = make_shard(n, K, rank)
    span = float(gaps.sum())
    halo_n = int(RATE * DTMAX * 4) + 64  # generous: the halo only needs the events within dtmax of the shard start
    if world > 1:
        spans = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(spans, torch.tensor([span], dtype=torch.float64, device=dev))
        start = float(sum(s.item() for s in spans[:rank]))
        tails_t = [torch.zeros(halo_n, dtype=torch.float64, device=dev) for _ in range(world)]
        tails_c = [torch.zeros(halo_n, dtype=torch.int64, device=dev) for _ in range(world)]
        t_local = start + np.cumsum(gaps)
        dist.all_gather(tails_t, torch.from_numpy(t_local[-halo_n:]).to(dev))
        dist.all_gather(tails_c, torch.from_numpy(nodes_np[-halo_n:]).to(dev))
        duration = float(sum(s.item() for s in spans)) * (1 + 1e-9)
    else:
        start, duration = 0.0, span * (1 + 1e-9)
        t_local = np.cumsum(gaps)
    del gaps
    if rank > 0:
        ht, hc = tails_t[rank - 1].cpu().numpy(), tails_c[rank - 1].cpu().numpy()
        keep = ht > t_local[0] - DTMAX
        ht, hc = ht[keep], hc[keep]
    else:
        ht, hc = np.zeros(0), np.zeros(0, np.int64)
    n_halo = ht.size
    h_t = torch.empty(n_halo + n, dtype=torch.float64, pin_memory=True)
    h_c = torch.empty(n_halo + n, dtype=torch.int64, pin_memory=True)
    h_t.numpy()[:n_halo], h_t.numpy()[n_halo:] = ht, t_local
    h_c.numpy()[:n_halo], h_c.numpy()[n_halo:] = hc, nodes_np
    del t_local, nodes_np
    index_base = rank * n - n_halo
    flags = 1 if rank == 0 else 0

    pl0, pW, pA, pmu, ptau = (np.ascontiguousarray(lam0), _fmat(W), _fmat(A), _fmat(mu), _fmat(tau))

    def set_params():
        ctx.check(lib.nhp_cont_params_set(ctx.h, 1, K, _ptr(pl0), _ptr(pW), _ptr(pA), _ptr(pmu), _ptr(ptau), DTMAX))

    def upload():
        h = ctypes.c_void_p()
        ctx.check(lib.nhp_events_upload(ctx.h, ctypes.c_void_p(h_t.data_ptr()), ctypes.c_void_p(h_c.data_ptr()), n_halo + n, duration, K, n_halo,
                                        max(index_base, 0), flags, ctypes.byref(h)))
        return h

    # stats buffers as torch tensors (zero-copy) for the NCCL allreduce
    def stats_tensor(phase):
        p, cnt = ctypes.c_void_p(), ctypes.c_int64()
        ctx.check(lib.nhp_cont_stats_dev(ctx.h, phase, ctypes.byref(p), ctypes.byref(cnt)))

        class _Raw:
            __cuda_array_interface__ = {"shape": (cnt.value,), "typestr": "<f8", "data": (p.value, False), "version": 3}
        return torch.as_tensor(_Raw(), device=dev)

    set_params()
    ev = upload()
    st0, st1 = stats_tensor(0), stats_tensor(1)
    op_ms = {"loglik": [], "parents": [], "second_pass": []}

    def step(counter, record=False):
        ll = ctypes.c_double()
        ctx.check(lib.nhp_cont_loglik(ctx.h, ev, 0, ctypes.byref(ll)))
        if record:
            op_ms["loglik"].append(ctx.last_kernel_ms)
        ctx.check(lib.nhp_cont_resample_parents(ctx.h, ev, 20261018, counter, None, None, None))
        if record:
            op_ms["parents"].append(ctx.last_kernel_ms)
        if world > 1:
            dist.all_reduce(st0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.check(lib.nhp_cont_suffstats_second_pass(ctx.h, ev))
        e1.record(stream)
        if world > 1:
            dist.all_reduce(st1)
        if record:
            e1.synchronize()
            op_ms["second_pass"].append(e0.elapsed_time(e1))
        return ll.value

    # ---- roofline denominators measured on this device
    peaks = {}
    if rank == 0:
        for which, key in ((0, "fp64_fma_tflops"), (1, "ln_pairs_per_s"), (2, "exp_pairs_per_s")):
            r = ctypes.c_double()
            ctx.check(lib.nhp_bench_fp64(ctx.h, which, ctypes.byref(r)))
            peaks[key] = r.value

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(w)
    sync_all()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches
    t_begin = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_begin.record(stream)
    for k in range(args.steps):
        step(args.warmup + k, record=True)
    t_end.record(stream)
    sync_all()
    total_ms = t_begin.elapsed_time(t_end)
    launches = ctx.launches - launches0
    sampler.stop_flag = True
    sampler.join(timeout=3)
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- e2e: the public-API call sequence with HOST buffers; H2D of the events + parameters and D2H of the result inside the timed region
    e2e_steps = max(1, min(args.steps, 3))
    lib.nhp_events_free(ctx.h, ev)
    K2 = K * K
    M0, Mn, Mnm, S1, S2 = np.empty(K), np.empty(K), np.empty(K2), np.empty(K2), np.empty(K2)
    sync_all()
    tb = torch.cuda.Event(enable_timing=True); te = torch.cuda.Event(enable_timing=True)
    tb.record(stream)
    e2e_parts = {"params_set": 0.0, "upload": 0.0, "loglik": 0.0, "gibbs": 0.0, "readback+free": 0.0}
    for k in range(e2e_steps):
        w0 = time.perf_counter()
        set_params()
        w1 = time.perf_counter()
        ev = upload()
        w2 = time.perf_counter()
        ll = ctypes.c_double()
        ctx.check(lib.nhp_cont_loglik(ctx.h, ev, 0, ctypes.byref(ll)))
        w3 = time.perf_counter()
        ctx.check(lib.nhp_cont_resample_parents(ctx.h, ev, 20261018, 100 + k, None, None, None))
        if world > 1:
            dist.all_reduce(st0)
        ctx.check(lib.nhp_cont_suffstats_second_pass(ctx.h, ev))
        if world > 1:
            dist.all_reduce(st1)
        stream.synchronize()
        w4 = time.perf_counter()
        ctx.check(lib.nhp_cont_suffstats_read(ctx.h, _ptr(M0), _ptr(Mn), _ptr(Mnm), _ptr(S1), _ptr(S2)))
        lib.nhp_events_free(ctx.h, ev)
        w5 = time.perf_counter()
        for key, dtv in zip(e2e_parts, (w1 - w0, w2 - w1, w3 - w2, w4 - w3, w5 - w4)):
            e2e_parts[key] += 1e3 * dtv / e2e_steps
    te.record(stream)
    sync_all()
    e2e_ms = torch.tensor([tb.elapsed_time(te)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    h2d = (n_halo + n) * 16 + (K + 4 * K2) * 8
    d2h = 8 + (2 * K + 3 * K2) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch device time measured with CUDA events on the launching stream)
    med = {k: float(np.median(v)) for k, v in op_ms.items() if v}
    dom = max(("loglik", "parents"), key=lambda k: med[k])
    secs = med[dom] * 1e-3
    density = float(np.count_nonzero(A * W)) / A.size
    probes = n * RATE * DTMAX                       # adjacency-bit probes (every window pair)
    pairs = probes * density                        # pairs with a non-zero effective weight: the FP64 impulse evaluations
    flops = pairs * FLOPS_PER_LN_PAIR + n * FLOPS_PER_EVENT
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    if os.path.exists(peaks_file):
        with open(peaks_file) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    bytes_alg = n * (BYTES_LOGLIK if dom == "loglik" else BYTES_PARENTS)
    gbs = bytes_alg / secs / 1e9
    sm_hz = (sampler.summary().get("sm_mhz") or 1965.0) * 1e6
    tf = flops / secs / 1e12
    traffic, issue_view = None, None
    prof = os.path.join(ROOT, "profiles", "r01_ncu_dominant_kernel.json")
    if os.path.exists(prof):
        with open(prof) as f:
            pj = json.load(f)
        if n == pj.get("events"):
            traffic = pj.get("dram_bytes_per_launch_at_1e8_events")
        ipe = pj.get("warp_instructions_per_event")
        if ipe:
            issue_peak = 148 * 4 * sm_hz  # one warp instruction per SM sub-partition per cycle
            issue_view = {"bound": "instruction issue", "achieved": ipe * n / secs, "peak": issue_peak, "unit": "warp-instructions/s",
                          "frac": ipe * n / secs / issue_peak, "warp_instructions_per_event": ipe,
                          "source": "instruction count per event from the committed ncu capture (profiles/r01_ncu_dominant_kernel.json), rate measured live"}
    # L1/shared-memory data pipe: utilisation seen by ncu for the same launch, rescaled by the live duration (same work)
    lsu_view = {"bound": "L1/shared-memory data pipe (LSU wavefronts)", "unit": "fraction of wavefront peak", "peak": 1.0, "achieved": None, "frac": None,
                "algorithmic_probes_per_launch": probes, "probes_per_s": probes / secs}
    if os.path.exists(prof) and pj.get("l1_lsu_wavefront_pct_of_peak") and n == pj.get("events"):
        lsu_frac = pj["l1_lsu_wavefront_pct_of_peak"] / 100.0 * (pj["duration_ms_under_ncu"] * 1e-3) / secs
        lsu_view.update({"achieved": lsu_frac, "frac": lsu_frac,
                         "source": "l1tex__data_pipe_lsu_wavefronts % of peak from profiles/r01_ncu_dominant_kernel.json x (ncu duration / live duration)"})
    roofline = {"kernel": "k_sweep_sparse<LOGITNORMAL, %s>" % ("LOGLIK" if dom == "loglik" else "PARENTS"),
                "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": traffic,
                "peak_source": hbm_src, "algorithmic_bytes_per_launch": bytes_alg,
                "binding_resource": "L1/shared-memory data pipe (~70 % of its wavefront peak) and instruction issue (~55 % of the slots), not HBM: 12 B of "
                                    "event data carry ~64 adjacency probes and ~3 FP64 impulse evaluations per event (ncu: profiles/r01_*.md); the HBM "
                                    "fraction is therefore small by construction",
                "issue_view": issue_view,
                "lsu_view": lsu_view,
                "fp64_view": {"bound": "fp64", "achieved": tf, "peak": peaks["fp64_fma_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["fp64_fma_tflops"],
                              "algorithmic_flops_per_launch": flops, "active_pairs_per_launch": pairs,
                              "peak_source": "FP64 FMA peak measured in this run by nhp_bench_fp64 (MEASURED_PEAKS.json carries no FP64 figure)",
                              "ln_pairs_per_s_register_ceiling": peaks["ln_pairs_per_s"]},
                "kernel_ms": med}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world), "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "what": "nhp_cont_params_set + nhp_events_upload (pinned host buffers) + nhp_cont_loglik + nhp_cont_resample_parents + statistics read back",
                    "host_ms_per_step": e2e_parts},
            "gpu_launches": int(launches), "roofline": roofline,
            "detail": {"loglik_events_per_s": world * n / (med["loglik"] * 1e-3), "gibbs_sweep_events_per_s": world * n / ((med["parents"] + med["second_pass"]) * 1e-3)}}

    # one-pass variant: the parent sweep also yields the log-likelihood (nhp_cont_sweep_loglik); reported, not the headline
    if world == 1:
        ev3 = upload()
        ctx.check(lib.nhp_set_option(ctx.h, 1, 1))  # NHP_OPT_SWEEP_LOGLIK
        fused = []
        for r in range(5):
            f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            ctx.check(lib.nhp_cont_resample_parents(ctx.h, ev3, 20261018, 500 + r, None, None, None))
            ctx.check(lib.nhp_cont_suffstats_second_pass(ctx.h, ev3))
            llf = ctypes.c_double()
            ctx.check(lib.nhp_cont_sweep_loglik(ctx.h, ev3, ctypes.byref(llf)))
            f1.record(stream); f1.synchronize()
            fused.append(f0.elapsed_time(f1))
        ctx.check(lib.nhp_set_option(ctx.h, 1, 0))
        line["detail"]["fused_loglik_and_gibbs_sweep_ms"] = float(np.median(fused[1:]))
        line["detail"]["fused_loglik_and_gibbs_sweep_events_per_s"] = n / (float(np.median(fused[1:])) * 1e-3)
        # whole Gibbs sweep on the device: parent sweep + statistics + second pass + conjugate draws of lambda0 / W / (mu, tau) +
        # table rebuild (nhp_cont_resample_params); the chain continues from the drawn parameters, so this goes last
        hyper = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
        full = []
        for r in range(4):
            f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            ctx.check(lib.nhp_cont_resample_parents(ctx.h, ev3, 20261018, 700 + r, None, None, None))
            ctx.check(lib.nhp_cont_resample_params(ctx.h, ev3, 20261018, 700 + r, float(duration), _ptr(hyper), hyper.size, 1))
            f1.record(stream); f1.synchronize()
            full.append(f0.elapsed_time(f1))
        lib.nhp_events_free(ctx.h, ev3)
        line["detail"]["gibbs_sweep_with_device_conjugate_draws_ms"] = float(np.median(full[1:]))
        set_params()  # restore the workload's parameters for whatever follows

    if world == 1 and args.adjacency:
        # continuous.jl:444-519 on the same resident data (not part of the timed step; reported for completeness)
        ev2 = upload()
        rho = np.full(K2, RHO)
        Acur = pA.copy()
        ms = []
        for r in range(2):
            ctx.check(lib.nhp_cont_resample_adjacency(ctx.h, ev2, _ptr(rho), 20261018, 900 + r, None, _ptr(Acur)))
            ms.append(ctx.last_kernel_ms)
        lib.nhp_events_free(ctx.h, ev2)
        line["detail"]["adjacency_sweep_ms"] = float(min(ms))
        line["detail"]["full_gibbs_sweep_incl_adjacency_events_per_s"] = n / ((med["parents"] + med["second_pass"] + min(ms)) * 1e-3)

    if world == 1 and not args.no_cpu_baseline:
        n_sample = int(min(args.cpu_sample, n))
        times, cores = cpu_step_time(args, n_sample, 2)
        cpu_v = n_sample / min(times)
        line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "first %d events of the workload (oracle loglik + resample_parents + statistics, OpenMP %d threads), best of 2" % (n_sample, cores)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # NCCL / torch banners must not pollute the single JSON line
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
