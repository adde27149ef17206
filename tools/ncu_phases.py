#!/usr/bin/env python3
"""Per-phase instruction / stall accounting of one kernel in an .ncu-rep captured with --import-source on.

usage: ncu_phases.py report.ncu-rep [kernel-instance-index] [units]
Splits the SASS at barrier-like markers (BAR.SYNC, mbarrier try-wait, bulk copies, EXIT) in address order and
prints executed warp instructions per phase (divided by `units`, e.g. the number of events) plus the stall
sampling totals; also the overall stall-reason mix.  Development aid for the profiles/ summaries."""
import csv, subprocess, sys, io

rep = sys.argv[1]
inst_idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
k = -1; hdr = None; ins = []; name = None; stalls = {}
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        k += 1
        if k == inst_idx: name = r[1]
        continue
    if r[0] == "Address": hdr = r; continue
    if k != inst_idx: continue
    try:
        ins.append((r[1].strip(), int(r[5]), int(r[4])))
    except Exception:
        continue
    for n, v in zip(hdr, r):
        if n.startswith("stall_") and "Not Issued" not in n:
            try: stalls[n] = stalls.get(n, 0) + int(v)
            except ValueError: pass
tot = sum(x[1] for x in ins); smp = sum(x[2] for x in ins)
print(f"{name}: {len(ins)} SASS instructions, {tot:.4g} executed warp instructions ({tot / units:.2f} per unit), {smp} samples")
acc = sm = 0; first = 0
for idx, (s, n, sp) in enumerate(ins):
    acc += n; sm += sp
    if s.startswith("BAR.") or "SYNCS.PHASECHK" in s or s.startswith("EXIT") or "UBLKCP" in s:
        if acc: print(f"  #{first:5d}-{idx:5d} ..{s[:42]:42s} inst {100 * acc / tot:5.1f}% ({acc / units:7.2f}/unit)  samples {100 * sm / max(smp, 1):5.1f}%")
        acc = sm = 0; first = idx + 1
ts = sum(stalls.values())
print("stall mix: " + ", ".join(f"{n[6:]} {100 * v / ts:.1f}%" for n, v in sorted(stalls.items(), key=lambda x: -x[1])[:10]))
