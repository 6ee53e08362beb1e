#!/usr/bin/env python3
"""Development: a SMALL staggered PickAndPlace batch stepped with plain launches (no graph), so that `ncu` sees launches of
the size the auto-reset tail of a full-size step has (one or a few warps per SM).  usage: small_run.py [envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_xarm_b200 import XarmVecEnv
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3456
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
task = os.environ.get("TASK", "pick_and_place")
env = XarmVecEnv(task, n, config=bench.bench_config(task), device="cuda:0", seed=0, auto_reset=True, stagger_phases=True, use_graph=False)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
for t in range(steps):
    env.step(torch.rand(n, env.act_dim, generator=g, device="cuda") * 2 - 1)
torch.cuda.synchronize()
print("ok", env.episode_stats())
