#!/usr/bin/env python3
"""Development: which finishers does the finish predictor of k_pipe_split miss (they cost their step a serial late tail)?
Steps the bench workload, watches the late branch's env-substep counter and prints, for the steps where it moved, the features the
predictor saw for the success finishers of that step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_xarm_b200 import XarmVecEnv
import bench
n = 131072
env = XarmVecEnv("pick_and_place", n, config=bench.bench_config("pick_and_place"), device="cuda:0", seed=0, auto_reset=True, stagger_phases=True)
env.reset()
env.set_profiling(True)
env.capture_graph()
g = torch.Generator(device="cuda").manual_seed(1234)
ring = [torch.rand(n, 4, generator=g, device="cuda") * 2 - 1 for _ in range(64)]
lastL = 0
for t in range(int(sys.argv[1]) if len(sys.argv) > 1 else 260):
    obs = env.obs_buf
    ag0, dg0 = obs["achieved_goal"].clone(), obs["desired_goal"].clone()
    o0 = obs["observation"].clone()
    st0 = torch.from_numpy(env.get_state()).cuda() if False else None
    o, r, d, info = env.step(ring[t % 64])
    torch.cuda.synchronize()
    kt = env.kernel_times()
    L = kt.get(("L", "#setup_envs"), (0, 0))[0]
    if L != lastL and t > 60:
        succ = d & (env.success_buf > 0)
        rel0 = o0[:, 21:24]; hd = rel0.norm(dim=1)
        v0 = (o0[:, 15:18] + o0[:, 3:6]).norm(dim=1)
        d0 = (ag0 - dg0).norm(dim=1)
        free_t = torch.clamp(3.0 * v0 / 60.0 + 9.8 / 3600.0 + 0.02, max=0.2075)
        far = torch.clamp(1.5 * v0 / 60.0 + 9.8 / 3600.0 + 0.003, max=0.11)
        mid = torch.clamp(free_t, max=0.11)
        tight = torch.where(hd < 0.13, torch.full_like(hd, 0.11), torch.where(hd > 0.24, far, mid))
        reach = 0.05 + tight
        cand = d0 < reach
        miss = succ & ~cand
        idx = torch.nonzero(d & ~cand).flatten()[:6].tolist()
        print(f"step {t}: late env-substeps +{L - lastL}; finishers {int(d.sum())} (success {int(succ.sum())}); emulated misses {int(miss.sum())}; done-but-not-candidate (incl. time limit) {int((d & ~cand).sum())}")
        for i in torch.nonzero(miss).flatten()[:4].tolist():
            print(f"   env {i}: d0 {d0[i]:.4f} reach {reach[i]:.4f} hd {hd[i]:.4f} v {v0[i]:.4f} lego z {ag0[i, 2]:.4f} goal z {dg0[i, 2]:.4f}")
    lastL = L
print("done; late env-substeps in total", lastL)
