#!/usr/bin/env python3
"""Development: how far does a lego travel within one env step, and how far from the goal were the envs that finished by success?
(input for the finish predictor of k_pipe_split: the early branch should hold every env that may finish, and few others)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_xarm_b200 import XarmVecEnv
import bench
n = 131072
env = XarmVecEnv("pick_and_place", n, config=bench.bench_config("pick_and_place"), device="cuda:0", seed=0, auto_reset=True, stagger_phases=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1234)
prev = None
D0s, trav, held_f = [], [], []
allt = []
rules = [(H, R) for H in (0.10, 0.13, 0.16, 0.20) for R in (0.10, 0.12, 0.14, 0.17)]
cand = {r: 0 for r in rules}
miss = {r: 0 for r in rules}
cur_cand = cur_miss = 0
rules3 = [(H1, R, H2, k) for H1 in (0.13, 0.15) for R in (0.10, 0.11, 0.12) for H2 in (0.24, 0.28) for k in (1.5, 3.0)]
cand3 = {r: 0 for r in rules3}
miss3 = {r: 0 for r in rules3}
heavy3 = {r: 0 for r in rules3}
for t in range(160):
    obs = env.obs_buf
    ag0, dg0 = obs["achieved_goal"].clone(), obs["desired_goal"].clone()
    rel0 = obs["observation"][:, 21:24].clone()          # lego pos - hand pos
    v0 = (obs["observation"][:, 15:18] + obs["observation"][:, 3:6]).norm(dim=1)   # lego velocity (obs holds it relative to the hand)
    o, r, d, info = env.step(torch.rand(n, 4, generator=g, device="cuda") * 2 - 1)
    if t < 60: continue
    term = env.terminal_buf
    succ = d & (env.success_buf > 0)
    ag1 = torch.where(d[:, None], term[:, 24:27], o["achieved_goal"])
    move = (ag1 - ag0).norm(dim=1)
    d0 = (ag0 - dg0).norm(dim=1)
    hd = rel0.norm(dim=1)
    free_r = torch.clamp(3.0 * v0 / 60.0 + 9.8 / 3600.0 + 0.02, max=0.2075)
    for (H, R) in rules:
        reach = 0.05 + torch.where(hd < H, torch.full_like(hd, R), torch.clamp(free_r, max=R))
        c = d0 < reach
        cand[(H, R)] += int(c.sum()); miss[(H, R)] += int((succ & ~c).sum())
    for r3 in rules3:
        H1, R, H2, k = r3
        mid = torch.clamp(free_r, max=R)
        far = torch.clamp(k * v0 / 60.0 + 9.8 / 3600.0 + 0.003, max=R)
        reach = 0.05 + torch.where(hd < H1, torch.full_like(hd, R), torch.where(hd > H2, far, mid))
        c = d0 < reach
        cand3[r3] += int(c.sum()); miss3[r3] += int((succ & ~c).sum()); heavy3[r3] += int((c & (hd < 0.10)).sum())
    D0s.append(d0[succ].cpu().numpy()); trav.append(move[succ].cpu().numpy()); held_f.append(rel0[succ].norm(dim=1).cpu().numpy())
    allt.append(move.cpu().numpy())
D0, TR, HF = np.concatenate(D0s), np.concatenate(trav), np.concatenate(held_f)
AT = np.concatenate(allt)
print("success finishers:", len(D0), "per step", len(D0) / 100)
for q in (50, 90, 99, 99.9, 100):
    print(f" q{q}: start distance to goal {np.percentile(D0, q):.4f}  lego travel in the step {np.percentile(TR, q):.4f}  |lego - hand| before {np.percentile(HF, q):.4f}")
print("all envs: lego travel per step quantiles", [round(float(np.percentile(AT, q)), 4) for q in (50, 90, 99, 99.9, 99.99, 100)])
for reach in (0.08, 0.10, 0.12, 0.15, 0.18, 0.21, 0.2575):
    print(f" reach {reach}: finishers outside {(D0 > reach).sum()} of {len(D0)}")

print("rule (hand-lego distance below H -> reach 0.05 + R, else the free-lego bound): candidates per step, missed finishers in 100 steps")
for r in rules:
    print(f"  H {r[0]:.2f} R {r[1]:.2f}: {cand[r] / 100:8.0f} candidates/step, missed {miss[r]}")

print("3-tier rule (hd < H1 -> 0.05 + R | hd > H2 -> free flight k |v| dt + g dt^2 + 0.003 | else min(3 |v| dt + g dt^2 + 0.02, R)): success+other candidates per step (time-limit ones come on top), of them with hd < 0.10, missed in 100 steps")
for r in rules3:
    print(f"  H1 {r[0]:.2f} R {r[1]:.2f} H2 {r[2]:.2f} k {r[3]:.1f}: {cand3[r] / 100:8.0f} candidates/step, near hand {heavy3[r] / 100:7.0f}, missed {miss3[r]}")
