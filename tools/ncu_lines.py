#!/usr/bin/env python3
"""Per-source-line hot spots of one profiled kernel: joins `ncu --page source --csv` (SASS view: samples and executed
instructions per address) with `nvdisasm -g` line info of the same cubin.
usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [launch-id]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, lib, pat = sys.argv[1:4]
lid = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-id", f":::{lid}"] if lid != "0" else []), capture_output=True, text=True).stdout.splitlines()
hdr_i = next(i for i, l in enumerate(out) if l.startswith('"Address"'))
print(out[0][:120])
end_i = next((i for i in range(hdr_i + 1, len(out)) if out[i].startswith('"Kernel Name"')), len(out))   # first kernel of the report only
rows = list(csv.DictReader(out[hdr_i:end_i]))
base = int(rows[0]["Address"], 16)
per_off = {int(r["Address"], 16) - base: (int(r["# Samples"] or 0), int(r["Instructions Executed"] or 0), r["Source"]) for r in rows}
# line info
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cubin = os.path.join(d, [f for f in os.listdir(d) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
cur_fn, cur_line, in_fn = None, "?", False
line_of = {}
for l in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        in_fn = pat in m.group(1); continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = f"{os.path.basename(m.group(1))}:{m.group(2)}"; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur_line
agg = collections.defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
for off, (smp, ins, src) in per_off.items():
    k = line_of.get(off, "?")
    agg[k][0] += smp; agg[k][1] += ins; tot_s += smp; tot_i += ins
print(f"total samples {tot_s}, warp-instructions {tot_i}, matched offsets {sum(1 for o in per_off if o in line_of)}/{len(per_off)}")
ranges = [a for a in sys.argv[5:] if ":" in a]   # file.cuh:name:lo-hi -> aggregate by source range
if ranges:
    tot = collections.defaultdict(lambda: [0, 0])
    for k, (smp, ins) in agg.items():
        f, _, ln = k.partition(":")
        name = f
        for r in ranges:
            rf, rn, rr = r.split(":"); lo, hi = map(int, rr.split("-"))
            if f == rf and ln.isdigit() and lo <= int(ln) <= hi:
                name = rn
        tot[name][0] += smp; tot[name][1] += ins
    for k, (smp, ins) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print(f"[{k:26s}] samples {smp:7d} ({100*smp/max(tot_s,1):5.1f}%)  inst {ins:9d} ({100*ins/max(tot_i,1):5.1f}%)")
    sys.exit(0)
for k, (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{k:28s} samples {smp:7d} ({100*smp/max(tot_s,1):5.1f}%)  inst {ins:9d} ({100*ins/max(tot_i,1):5.1f}%)")
