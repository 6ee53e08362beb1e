#!/usr/bin/env python3
"""Development: %globaltimer timeline of one graph-launched xarm_step (XARM_TIMELINE=1).  usage: timeline.py [steps_before] [n]"""
import sys, os, ctypes as C, collections
os.environ["XARM_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_xarm_b200 import XarmVecEnv, _native
import bench
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
env = XarmVecEnv("pick_and_place", n, config=bench.bench_config("pick_and_place"), device="cuda:0", seed=0, auto_reset=True, stagger_phases=bool(os.environ.get("STAGGER")))
env.reset()
if not os.environ.get("NO_GRAPH"):
    env.capture_graph()
g = torch.Generator(device="cuda").manual_seed(1234)
L = _native.load()
L.xarm_debug_timeline.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
buf = C.create_string_buffer(1 << 20)
for t in range(steps):
    env.step(torch.rand(n, 4, generator=g, device="cuda") * 2 - 1)
    if t >= steps - int(os.environ.get("LAST", 2)):
        torch.cuda.synchronize()
        L.xarm_debug_timeline(env._h, buf, len(buf))
        rows = [l.split() for l in buf.value.decode().splitlines()]
        rows = [(r[0], r[1], float(r[2]), float(r[3]), int(r[4]) if len(r) > 4 else 0) for r in rows]
        print(f"== step {t}: {len(rows)} launches with work")
        for br in "EML":
            rb = [r for r in rows if r[0] == br]
            if not rb: continue
            t0, t1 = min(r[2] for r in rb), max(r[3] for r in rb)
            acc = collections.defaultdict(lambda: [0, 0.0])
            for r in rb:
                acc[r[1]][0] += 1; acc[r[1]][1] += r[3] - r[2]
            print(f" branch {br}: {t0/1e3:.2f} .. {t1/1e3:.2f} ms | " + " | ".join(f"{k}: {v[0]} x {v[1]/v[0]:.0f} us = {v[1]/1e3:.2f} ms" for k, v in acc.items()))
        if os.environ.get("PASSES"):
            for name in ("setup", "heavy_rows", "heavy_solve", "heavy_fused", "light"):
                rb = sorted([r for r in rows if r[0] == "E" and r[1] == name], key=lambda r: r[2])
                per = [rb[k:k + 15] for k in range(0, len(rb), 15)]
                print(f"  E {name:12s} per pass avg us: " + " ".join(f"{sum(r[3]-r[2] for r in p_)/len(p_):.0f}" for p_ in per) + "   max: " + " ".join(f"{max(r[3]-r[2] for r in p_):.0f}" for p_ in per))
            for br in "EM":
                rb = sorted([r for r in rows if r[0] == br and r[1] in ("heavy_rows", "heavy_fused")], key=lambda r: r[2])
                per = [rb[k:k + 15] for k in range(0, len(rb), 15)]
                print(f"  {br} heavy env count per pass (avg): " + " ".join(f"{sum(r[4] for r in p_)/len(p_):.0f}" for p_ in per))
            rb = sorted([r for r in rows if r[0] == "E"], key=lambda r: r[2])
            st = [r for r in rb if r[1] == "setup"]
            print("  E pass durations ms: " + " ".join(f"{(st[min(k+15, len(st)-1)][2]-st[k][2])/1e3:.2f}" for k in range(0, len(st), 15)))
        if os.environ.get("SEQ"):
            rb = sorted([r for r in rows if r[0] == os.environ["SEQ"]], key=lambda r: r[2])
            print(" ".join(f"{r[1][:3] if r[1] != 'heavy_solve' else 'SOL'}@{r[2]:.0f}+{r[3]-r[2]:.0f}" for r in rb[:int(os.environ.get("SEQN", 80))]))
