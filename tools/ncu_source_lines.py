"""Stall samples and executed instructions of one kernel of an ncu report, grouped by the line of the KERNEL BODY they belong to.

  python tools/ncu_source_lines.py <report.ncu-rep> <object.o | cubin> <mangled kernel name> [table index]

The SASS page of `ncu --import-source on` lists the samples per instruction; `nvdisasm -gi` of the same binary gives every instruction's
inline chain (-lineinfo), whose outermost entry is the line of the kernel body the instruction was inlined into.  The two listings are matched
by instruction order, so the object file must be the one the report was captured from.  (ncu prints the table of a kernel twice; `table index`
picks the table: 0, 2, 4, ... for the first, second, third kernel of the capture.)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def line_map(binary, kernel):
    with tempfile.TemporaryDirectory() as tmp:
        cubin = binary
        if binary.endswith(".o") or binary.endswith(".so"):
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(binary)], cwd=tmp, capture_output=True)
            cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
            cubin = os.path.join(tmp, cubins[0])
        text = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(text) if l.startswith(".text." + kernel + ":")][0]
    ctx, fresh, out = [], True, []
    for l in text[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if not fresh:
                ctx, fresh = [], True
            ctx.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        if re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
            fresh = False
            out.append(ctx[-1] if ctx else ("?", 0))
    return out


def main():
    rep, binary, kernel = sys.argv[1:4]
    table = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    lines = line_map(binary, kernel)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hi, end = heads[table], (heads[table + 1] if table + 1 < len(heads) else len(rows))
    hdr = rows[hi]
    i_inst, i_smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[hi + 1:end] if len(r) > i_smp and r[0].startswith("0x")]
    if len(body) != len(lines):
        sys.exit("instruction counts differ (%d in the report, %d in the binary): not the binary the report was captured from" % (len(body), len(lines)))
    tot_s, tot_i = sum(int(r[i_smp]) for r in body) or 1, sum(int(r[i_inst]) for r in body) or 1
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for key, r in zip(lines, body):
        a = agg[key]
        a[0] += int(r[i_smp]); a[1] += int(r[i_inst])
        for j in stall:
            try:
                a[2][hdr[j][6:]] += int(r[j])
            except ValueError:
                pass
    print("| kernel-body line | stall samples | warp instructions | largest stall reasons |")
    print("|---|---:|---:|---|")
    for key, a in sorted(agg.items()):
        if a[0] / tot_s >= 0.006 or a[1] / tot_i >= 0.006:
            top = ", ".join("%s %d %%" % (k, round(100 * v / max(a[0], 1))) for k, v in a[2].most_common(3))
            print("| `%s:%d` | %.1f %% | %.1f %% | %s |" % (key[0], key[1], 100 * a[0] / tot_s, 100 * a[1] / tot_i, top))


if __name__ == "__main__":
    main()
