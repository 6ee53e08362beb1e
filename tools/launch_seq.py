#!/usr/bin/env python3
"""Print the kernel sequence of an ncu launch list (gpu__time_duration csv): tag + microseconds, and per-kernel totals."""
import collections, csv, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
tag = {'k_heavy_solve2': 'Z', 'k_heavy_solve': 'Y', 'k_heavy_rows': 'W', 'k_pipe_heavy_all': 'HA', 'k_pipe_heavy': 'H', 'k_pipe_setup': 'S', 'k_pipe_light': 'L', 'k_pipe_reset': 'R', 'k_pipe_action': 'A',
       'k_pipe_finish': 'F', 'k_pipe_begin': 'B', 'k_pipe_split': 'P'}
out, agg = [], collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k, t = r['Kernel Name'], float(r['Metric Value']) / 1e3
    c = [v for kk, v in tag.items() if kk in k]
    out.append(f"{c[0] if c else '?'}{t:.0f}")
    agg[k[:44]][0] += 1; agg[k[:44]][1] += t
if '-q' not in sys.argv:
    print(' '.join(out))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} n={v[0]:4d} total={v[1]:10.1f} us  avg={v[1]/v[0]:8.1f} us  share={v[1]/tot:.3f}")
