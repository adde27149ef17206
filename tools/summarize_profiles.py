"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEYS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1/LSU data-pipe wavefronts % of peak'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'FP64 pipe %'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % (occupancy)'),
        ('smsp__inst_executed.sum', 'warp instructions'), ('smsp__thread_inst_executed_per_inst_executed.ratio', 'threads per instruction'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long scoreboard / issue'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short scoreboard / issue'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math pipe / issue'),
        ('launch__registers_per_thread', 'registers / thread'), ('launch__grid_size', 'grid'),
        ('launch__shared_mem_per_block_dynamic', 'dynamic smem / CTA'), ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %')]


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    hdr, units = rr[0], rr[1]
    return [{h: (row[i], units[i]) for i, h in enumerate(hdr)} for row in rr[2:]]


def table(title, cmd, rep, note):
    rows = raw(rep)
    md = ["# " + title, "", cmd, ""]
    md.append("| metric | " + " | ".join("`%s`" % r['Kernel Name'][0][:70] for r in rows) + " |")
    md.append("|---|" + "---:|" * len(rows))
    for k, label in KEYS:
        if k in rows[0]:
            md.append("| %s (%s) | " % (label, rows[0][k][1]) + " | ".join(r[k][0] for r in rows) + " |")
    md += ["", note, ""]
    return "\n".join(md), rows


# ---- launch list
lines = [r for r in csv.reader(open(os.path.join(G, 'r01_launches.csv'))) if r]
hi = [i for i, r in enumerate(lines) if r[0] == 'ID'][0]
hdr = lines[hi]
iname, ival, iunit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in lines[hi + 1:]:
    if len(r) <= ival:
        continue
    v = float(r[ival].replace(',', ''))
    u = r[iunit]
    ms = v / 1e6 if u in ('ns', 'nsecond') else (v / 1e3 if u in ('us', 'usecond') else v)
    agg.setdefault(r[iname], []).append(ms)
tot = sum(sum(v) for v in agg.values())
md = ["# r01 ncu launch list -- `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (N=1, 1e8 events, K=1000, rho=0.05)", "",
      "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` after the same command exited 0 without ncu. Per-launch times are cold-cache and serialised: compare SHARES.",
      "The timed step consists of `k_sweep_sparse<1,0>` (log-likelihood), `k_reduce_partials`, `k_sweep_sparse<1,2>` (parent sweep + statistics), `k_xbar`, `k_second_pass`;",
      "the remaining launches are set-up (upload, parameter tables, window prepass) and the roofline micro-benchmarks.", "",
      "| kernel | launches | total ms | share | mean ms |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    md.append("| `%s` | %d | %.3f | %.1f%% | %.3f |" % (k[:110], len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
open(os.path.join(P, 'r01_launches.md'), 'w').write("\n".join(md) + "\n")
import shutil
shutil.copy(os.path.join(G, 'r01_launches.csv'), os.path.join(P, 'r01_launches.csv'))

# ---- sweep kernels of the bench step
txt, rows = table("r01 ncu full capture -- the two sweep kernels of the bench step (N=1, 1e8 events, K=1000, rho=0.05, w=64)",
                  "`ncu --set full --clock-control none --import-source on -k regex:k_sweep_sparse -s 6 -c 2` under `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`.",
                  os.path.join(G, 'r01_prof_sparse.ncu-rep'),
                  "Reading: DRAM traffic equals the algorithmic bytes (1.2 GB of (time, node) records read once + 0.2 GB of cached window lengths; the parent sweep adds the 0.4 GB "
                  "parent-offset write): nothing is re-read from HBM. DRAM throughput is ~2 % of peak because the kernels are bound on-chip: ~65-70 warp instructions per event "
                  "(64 adjacency-bit probes, bit-row staging, compaction, ~3 FP64 impulse evaluations) at ~50-55 % of the issue slots, and the L1/shared-memory data pipe at "
                  "~65-75 % of its wavefront peak. History of this kernel within the round (1e7-event probe, log-likelihood / parent sweep): 1.79 / 1.99 ms -> 1.10 / 1.26 ms by "
                  "(i) keeping shared-memory accesses in the shared address space (no generic LD/ST), (ii) 128-event tiles with two filter threads per event, "
                  "(iii) a lane-per-event mapping with bank-private 'vertical' adjacency bit rows (bank conflicts 138 M -> 26 M per 1e7 events), "
                  "(iv) uniform 8-probe filter trips, (v) 256-bit bit-row gathers (half the L1 tag look-ups), (vi) the compensator from per-node counts. "
                  "Remaining stalls: block barriers between the four phases of a tile, long scoreboard on the bit-row / parameter-table gathers (L2), short scoreboard on the probes.")
open(os.path.join(P, 'r01_ncu_sweep_kernels.md'), 'w').write(txt)
f = lambda d, k: float(d[k][0].replace(',', ''))
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
import re
par = [d for d in rows if re.search(r'<\(?\w*\)?\d+, \(?\w*\)?2, ', d['Kernel Name'][0])][0]  # MODE == 2: the parent sweep
dom = {"kernel": par['Kernel Name'][0], "events": 100000000,
       "dram_bytes_read": f(par, 'dram__bytes_read.sum') * scale[par['dram__bytes_read.sum'][1]],
       "dram_bytes_write": f(par, 'dram__bytes_write.sum') * scale[par['dram__bytes_write.sum'][1]],
       "duration_ms_under_ncu": f(par, 'gpu__time_duration.sum'),
       "issue_active_pct": f(par, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
       "warp_instructions": f(par, 'smsp__inst_executed.sum'),
       "l1_lsu_wavefront_pct_of_peak": f(par, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
       "smem_bank_conflicts": f(par, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
       "smem_wavefronts": f(par, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
       "source": "gpurun_out/r01_prof_sparse.ncu-rep (ncu --set full --clock-control none, bench.py --steps 2 --warmup 3)"}
dom["dram_bytes_per_launch_at_1e8_events"] = dom["dram_bytes_read"] + dom["dram_bytes_write"]
dom["warp_instructions_per_event"] = dom["warp_instructions"] / dom["events"]
json.dump(dom, open(os.path.join(P, 'r01_ncu_dominant_kernel.json'), 'w'), indent=1)

txt, _ = table("r01 ncu full capture -- dense sweeps (config 2: LogitNormal standard, K=50, 1e6 events, w=100)",
               "`ncu --set full --clock-control none --import-source on -k regex:\"k_sweep|k_parents\" -s 2 -c 4` under `python tools/devbench.py one 50 1e6 100 none ln 1`.",
               os.path.join(G, 'r01_prof_dense.ncu-rep'),
               "Reading: every pair costs ~107 instructions (42 FP64); the kernel is bound by the L1/shared-memory data pipe (log/exp table look-ups with bank conflicts, "
               "the divergent parameter-table gather) and by instruction issue; the FP64 pipe is ~35-45 % busy. DRAM traffic is the 12 MB of event records.")
open(os.path.join(P, 'r01_ncu_dense_kernels.md'), 'w').write(txt)

txt, _ = table("r01 ncu full capture -- discrete contraction on FP64 tensor cores (N=200, B=6, 2e5 bins)",
               "`ncu --set full --clock-control none --import-source on -k regex:k_disc_dmma -c 1` under `python tools/devbench.py disc 200 2e5 6 12 0.04`.",
               os.path.join(G, 'r01_prof_dmma.ncu-rep'),
               "Reading: `DMMA.8x8x4` (mma.sync.m8n8k4.f64) register-pipelined GEMM with the Poisson log-likelihood fused into the epilogue: 24.7 TFLOP/s at config 3 "
               "(1e6 bins) = 67 % of the measured DMMA peak (36.7 TFLOP/s).")
open(os.path.join(P, 'r01_ncu_dmma_kernel.md'), 'w').write(txt)
print(open(os.path.join(P, 'r01_launches.md')).read()[:1500])
print(json.dumps(dom, indent=1))
