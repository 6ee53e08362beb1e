#!/usr/bin/env python3
"""DRAM traffic of ONE env step, measured with ncu (bench.py's roofline.traffic reads the file this writes).

  # on the GPU box (one GPU; ncu replays every kernel, so this is never a bench number):
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none --csv --log-file gpurun_out/traffic_raw.csv python tools/ncu_traffic.py run [task] [envs]
  python tools/ncu_traffic.py parse gpurun_out/traffic_raw.csv [task] [envs]     # -> profiles/traffic_<task>.json + a launch summary

`run`: the bench workload (staggered phases, pre-rolled) stepped with plain launches; the profiler is switched on around ONE
step, so the capture holds every kernel launch of that step (main branch, early branch with its auto-reset passes, late tail).
`parse`: sums dram__bytes_read.sum + dram__bytes_write.sum over the launches, per kernel and in total, and stamps the result
with the sha256 of the kernel sources - bench.py refuses a capture taken from other sources or at another size.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(task, n):
    import torch
    import bench
    from gym_xarm_b200 import XarmVecEnv
    from gym_xarm_b200.specs import SPECS
    env = XarmVecEnv(task, n, config=bench.bench_config(task), device="cuda:0", seed=0, auto_reset=True, stagger_phases=True, use_graph=False)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1234)
    pre = int(1.2 * SPECS[task].max_episode_steps) + 5
    for _ in range(pre):
        env.step(torch.rand(n, env.act_dim, generator=g, device="cuda") * 2 - 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    env.step(torch.rand(n, env.act_dim, generator=g, device="cuda") * 2 - 1)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one step of", n, task, "envs after", pre, "pre-roll steps;", env.episode_stats())


def parse(path, task, n, write=True):
    import bench
    import gzip
    rows = []
    with (gzip.open(path, "rt", newline="") if path.endswith(".gz") else open(path, newline="")) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    per = collections.OrderedDict()
    launches = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"<.*", "", r["Kernel Name"]).split("(")[0]
        if not name.startswith(("k_", "void k_")):
            continue
        name = name.replace("void ", "")
        metric, val = r["Metric Name"], float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "")
        if metric.startswith("dram__bytes"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d = per.setdefault(name, {"launches": set(), "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
            d["dram_read" if "read" in metric else "dram_write"] += val * mult
            d["launches"].add(r["ID"])
        elif metric.startswith("gpu__time_duration"):
            mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            d = per.setdefault(name, {"launches": set(), "dram_read": 0.0, "dram_write": 0.0, "us": 0.0})
            d["us"] += val * mult
            d["launches"].add(r["ID"])
    total = sum(d["dram_read"] + d["dram_write"] for d in per.values())
    tot_us = sum(d["us"] for d in per.values()) or 1.0
    kernels = [{"kernel": k, "launches": len(d["launches"]), "dram_read_bytes": d["dram_read"], "dram_write_bytes": d["dram_write"],
                "sum_us_under_ncu": d["us"], "share_of_time_under_ncu": d["us"] / tot_us} for k, d in sorted(per.items(), key=lambda kv: -(kv[1]["dram_read"] + kv[1]["dram_write"]))]
    out = {"task": task, "envs": n, "kernel_source_hash": bench.kernel_source_hash(), "dram_bytes_per_step": total,
           "algorithmic_bytes_per_step": bench.ALGO_BYTES[task] * n, "kernels": kernels,
           "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every kernel launch of one staggered, pre-rolled step "
                   "(plain launches, --clock-control none; per-launch times under ncu are serialised and cold-cache: shares only)"}
    dst = os.path.join(ROOT, "profiles", f"traffic_{task}.json")
    if not write:
        return out
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{dst}: {total / 1e9:.3f} GB of DRAM traffic per step of {n} envs ({total / (bench.ALGO_BYTES[task] * n):.1f} x the algorithmic {bench.ALGO_BYTES[task] * n / 1e6:.1f} MB)")
    for k in kernels[:10]:
        print(f"  {k['kernel']:28s} {k['launches']:5d} launches | read {k['dram_read_bytes'] / 1e6:9.1f} MB | write {k['dram_write_bytes'] / 1e6:9.1f} MB | {k['sum_us_under_ncu'] / 1e3:8.2f} ms ({100 * k['share_of_time_under_ncu']:.1f} %)")


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "run"
    if mode == "run":
        run(sys.argv[2] if len(sys.argv) > 2 else "pick_and_place", int(sys.argv[3]) if len(sys.argv) > 3 else 131072)
    else:
        parse(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "pick_and_place", int(sys.argv[4]) if len(sys.argv) > 4 else 131072)
