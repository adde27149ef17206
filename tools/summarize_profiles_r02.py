"""Round-2 summaries under profiles/ from the ncu artefacts brought back in gpurun_out/ (see profiles/README.md).

  python tools/summarize_profiles_r02.py [step_report] [build_report] [launch_csv] [bench_json]
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
KEYS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots active %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1/LSU data-pipe wavefronts % of peak'),
        ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'FP64 pipe %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active % (occupancy)'),
        ('smsp__inst_executed.sum', 'warp instructions'), ('smsp__thread_inst_executed_per_inst_executed.ratio', 'threads per instruction'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
        ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long scoreboard / issue'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short scoreboard / issue'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not selected / issue'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math pipe / issue'),
        ('launch__registers_per_thread', 'registers / thread'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__cluster_size', 'cluster size'),
        ('launch__shared_mem_per_block_dynamic', 'dynamic smem / CTA'), ('lts__t_sector_hit_rate.pct', 'L2 hit rate %')]
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(out)))
    hdr, units = rr[0], rr[1]
    return [{h: (row[i], units[i]) for i, h in enumerate(hdr)} for row in rr[2:]]


def table(title, cmd, rows, note):
    md = ["# " + title, "", cmd, ""]
    md.append("| metric | " + " | ".join("`%s`" % r['Kernel Name'][0][:70] for r in rows) + " |")
    md.append("|---|" + "---:|" * len(rows))
    for k, label in KEYS:
        if k in rows[0]:
            md.append("| %s (%s) | " % (label, rows[0][k][1]) + " | ".join(r[k][0] for r in rows) + " |")
    md += ["", note, ""]
    return "\n".join(md)


def regions(rep, kernel_index=None, min_share=0.01):  # kernel_index: substring of the kernel name
    """Per-region accounting from the SASS page: consecutive instructions with a similar execution count form a region (a loop body,
    a phase); reports each region's share of the stall samples and of the executed warp instructions."""
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # one table (sometimes repeated) per kernel of the report: "Kernel Name" line, header line ("Address", ...), instruction rows
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    pick = heads[0]
    if kernel_index is not None:
        named = [h for h in heads if h > 0 and rows[h - 1] and rows[h - 1][0] == "Kernel Name" and str(kernel_index) in rows[h - 1][1]]
        pick = named[0] if named else heads[0]
    hi = pick
    end = min([h for h in heads if h > hi] + [len(rows)])
    hdr = rows[hi]
    iS, iI, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    body = [r for r in rows[hi + 1:end] if len(r) > iP and r[0].startswith("0x")]
    tot_s = sum(int(r[iP]) for r in body) or 1
    tot_i = sum(int(r[iI]) for r in body) or 1
    segs, k0, cur, acc = [], 0, None, []
    for k, r in enumerate(body):
        c = int(r[iI])
        if cur is None or not ((0.5 * cur <= c <= 2 * cur) if cur > 0 else c == 0):
            if acc:
                segs.append((k0, k - 1, cur, acc))
            k0, cur, acc = k, c, []
        acc.append((int(r[iP]), c, r[iS].strip()))
    segs.append((k0, len(body) - 1, cur, acc))
    out_rows = []
    for a, b, c, acc in segs:
        s, i = sum(x[0] for x in acc), sum(x[1] for x in acc)
        if s / tot_s >= min_share or i / tot_i >= min_share:
            top = max(acc, key=lambda x: x[0])
            out_rows.append((a, b, c, 100.0 * s / tot_s, 100.0 * i / tot_i, 100.0 * top[0] / tot_s, top[2][:60]))
    return out_rows, tot_i


def f(d, k):
    return float(d[k][0].replace(',', ''))


def main():
    step_rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(G, 'r6a_prof_step.ncu-rep')
    build_rep = sys.argv[2] if len(sys.argv) > 2 else os.path.join(G, 'r3p_prof_build.ncu-rep')
    launch_csv = sys.argv[3] if len(sys.argv) > 3 else os.path.join(G, 'r6a_launches.csv')
    bench_json = sys.argv[4] if len(sys.argv) > 4 else os.path.join(G, 'r6a_bench.json')
    os.makedirs(P, exist_ok=True)

    # ---- launch list
    lines = [r for r in csv.reader(open(launch_csv)) if r]
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
    md = ["# r02 ncu launch list -- `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (N=1, cfg4: 1e8 events, K=1000, rho=0.05, hawkes data)", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600` after the same command exited 0 without ncu.  Per-launch times are cold-cache and",
          "serialised: compare SHARES.  One step = `k_sweep_sparse<1,0>` (log-likelihood) + `k_sweep_sparse<1,2>` (parent sweep + statistics) + `k_xbar`, `k_second_pass`,",
          "`k_conjugate`, the table rebuild, `k_adj_prep` + `k_adj_sweep` (adjacency sweep) + `k_beta_draw`; from the second step on the log-likelihood is `k_adj_loglik` (active buckets of the cached structure); `k_adj_build` / `k_adj_count` / `k_adj_links` and the CUB sort run once",
          "per events handle (structure build); `k_rand_*` is the device simulator that generates the stream; the rest is set-up and the roofline micro-benchmarks.", "",
          "| kernel | launches | total ms | share | mean ms |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        md.append("| `%s` | %d | %.3f | %.1f%% | %.3f |" % (k[:110], len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v)))
    open(os.path.join(P, 'r02_launches.md'), 'w').write("\n".join(md) + "\n")
    shutil.copy(launch_csv, os.path.join(P, 'r02_launches.csv'))

    # ---- the three kernels of the step
    rows = raw(step_rep)
    adj = [d for d in rows if 'k_adj_sweep' in d['Kernel Name'][0]][0]
    pairs = json.load(open(bench_json))["roofline"]["adjacency_sweep"]["pairs"]
    reg, tot_i = regions(step_rep, 'k_adj_sweep')
    note = ["Reading (adjacency sweep, `k_adj_sweep<LOGITNORMAL, payload, cluster>`): %.1f GB read for %.1f GB of cached pairs (18 B x %.3g pairs; the rest is the" %
            (f(adj, 'dram__bytes_read.sum') * SCALE[adj['dram__bytes_read.sum'][1]] / 1e9, 18 * pairs / 1e9, pairs),
            "3 %% section padding, the re-read of the buckets whose link flipped and of the links that are on at the start of a column); %.1f warp instructions per 32 pairs" %
            (f(adj, 'smsp__inst_executed.sum') / (pairs / 32.0)),
            "(round 1: 211).  The kernel is bound by instruction issue inside the streaming phase and by the barriers between the phases of a batch",
            "(stream 32 buckets -> exchange of the sums through distributed shared memory + one cluster barrier -> decisions, taken by every warp alike -> flips).  Regions of the SASS page (consecutive instructions with similar execution counts):", "",
            "| SASS instructions | executions | stall samples | warp instructions | hottest instruction (share of samples) |", "|---|---:|---:|---:|---|"]
    for a, b, c, s, i, t, src in reg:
        note.append("| %d-%d | %d | %.1f %% | %.1f %% | `%s` (%.1f %%) |" % (a, b, c, s, i, src, t))
    note += ["", "The region that executes once per 64-entry block of singles (~1e8 executions) is the streaming loop (140 instructions per 64 pairs: one 32-bit + one 256-bit load, two table-driven",
             "exps, two shared-memory intensity look-ups, product accumulation, three min/max trackers, an L2 prefetch); the regions with ~5e6 executions are the",
             "per-batch phases (32 warps x 130 CTAs x batches), those with ~8e6 the application of a link (a column's initial intensities, flips); `UCGABAR_WAIT` is the cluster barrier",
             "that ends a batch (waiting for the slowest warp of the slowest CTA), the `BRA` behind `BAR.SYNC` the barrier of a link's application."]
    obj = os.path.join(ROOT, "networkhawkesprocesses.jl_b200", "csrc", "build", "cont_adjacency.o")
    if os.path.exists(obj):  # the same samples by line of the kernel body (valid when the object is the build the report was captured from)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_lines.py"), step_rep, obj, "_Z11k_adj_sweepILi1ELi1ELb1EEv12AdjSweepArgs", "0"],
                           capture_output=True, text=True)
        if r.returncode == 0:
            note += ["", "The same samples by line of the kernel body (`python tools/ncu_source_lines.py`; `cont_adjacency.cu` as of the captured build (name its commit here): `adj_singles` /",
                     "`adj_runs` calls = the streaming phase, `cluster_sync_all` = the batch's barrier, the `adj_apply` call in the column prologue = the links that are on,",
                     "the `adj_apply` call behind the decisions = flips):", ""] + r.stdout.strip().split("\n")
    txt = table("r02 ncu full capture -- the three kernels of the bench step (N=1, cfg4, hawkes data): log-likelihood through the cached structure (`k_adj_loglik`), parent sweep (`k_sweep_sparse<1,2>`), adjacency sweep (`k_adj_sweep`)",
                "`ncu --set full --clock-control none --import-source on --kernel-name regex:\"k_adj_sweep|k_sweep_sparse|k_adj_loglik\" --launch-skip 5 --launch-count 3` under "
                "`python bench.py --steps 2 --warmup 1 --no-cpu-baseline`.", rows, "\n".join(note))
    open(os.path.join(P, 'r02_ncu_step_kernels.md'), 'w').write(txt)
    dom = {"kernel": adj['Kernel Name'][0], "pairs": pairs,
           "dram_bytes_read": f(adj, 'dram__bytes_read.sum') * SCALE[adj['dram__bytes_read.sum'][1]],
           "dram_bytes_write": f(adj, 'dram__bytes_write.sum') * SCALE[adj['dram__bytes_write.sum'][1]],
           "duration_ms_under_ncu": f(adj, 'gpu__time_duration.sum'),
           "issue_active_pct": f(adj, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
           "warp_instructions": f(adj, 'smsp__inst_executed.sum'), "grid": f(adj, 'launch__grid_size'),
           "source": "gpurun_out/%s (ncu --set full --clock-control none, bench.py --steps 2 --warmup 1 --no-cpu-baseline)" % os.path.basename(step_rep)}
    dom["dram_bytes_per_launch"] = dom["dram_bytes_read"] + dom["dram_bytes_write"]
    dom["warp_instructions_per_32_pairs"] = dom["warp_instructions"] / (pairs / 32.0)
    json.dump(dom, open(os.path.join(P, 'r02_ncu_dominant_kernel.json'), 'w'), indent=1)

    # ---- structure build
    if os.path.exists(build_rep):
        brows = raw(build_rep)
        txt = table("r02 ncu full capture -- adjacency structure build `k_adj_build` (1e8 events, K=1000, uniform nodes; once per events handle)",
                    "`ncu --set full --clock-control none --import-source on --kernel-name regex:k_adj_build --launch-count 1` under `python tools/devbench.py adj 1000 1e8 64 0.05 reset`.",
                    brows,
                    "Reading: the kernel scatters 6.4e9 entries (18 B each: 116 GB) from event-major to parent-major order.  History of the round (DRAM read / write, duration):\n"
                    "per-warp cursors 730 / 572 GB, 778 ms (every 8-byte store its own partial sector) -> CTA-wide cursors for the singles 401 / 349 GB, 425 ms -> "
                    "16-byte (logit, Jacobian) records, one 32-warp CTA per SM, L2 evict-last stores, bounded look-ahead prefetch: 373 / 161 GB, 184 ms (this capture; 177 ms after the\n"
                    "prefetch was bounded to 4 events per warp).  Written bytes are now 1.4x the payload; the reads are the two passes over the packed (node, same-node link) records\n"
                    "and the times (156 GB) plus what L2 could not keep; the kernel is latency bound (long scoreboard on the window loads) at 43 % of the issue slots.")
        open(os.path.join(P, 'r02_ncu_adj_build.md'), 'w').write(txt)
    shutil.copy(bench_json, os.path.join(P, 'r02_bench_line_n1.json'))
    print(open(os.path.join(P, 'r02_ncu_step_kernels.md')).read())


if __name__ == "__main__":
    main()
