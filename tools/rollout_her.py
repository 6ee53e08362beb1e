#!/usr/bin/env python3
"""Env step + HER buffer together on one GPU: XarmPDPickAndPlace-v0, random actions, every step followed by
XarmHerReplayBuffer.add() and one sample(batch) - the data path of the reference's sparse-reward training arm
[REF benchmark/train.py:81-97] without the learner.  Prints one JSON line (CUDA events around the timed steps).
usage: rollout_her.py [num_envs] [steps] [batch]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_xarm_b200 as gx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
env = gx.make_vec("XarmPDPickAndPlace-v0", n, device="cuda:0", seed=0)
buf = gx.XarmHerReplayBuffer(env, episodes_per_env=4, n_sampled_goal=4, seed=0)
env.reset(); buf.begin()
env.capture_graph()
g = torch.Generator(device="cuda").manual_seed(0)
ring = [torch.rand(n, env.act_dim, generator=g, device="cuda") * 2 - 1 for _ in range(16)]

def run(k, with_her):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(k):
        env.step(ring[i % 16])
        if with_her:
            buf.add()
            if buf_ready:
                buf.sample(batch)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k

buf_ready = False
run(60, True)            # warm-up: one full episode per env, so that every ring holds a finished episode
buf_ready = True
ms_her = run(steps, True)
ms_env = run(steps, False)
st = buf.stats()
r = buf.sample(batch).rewards
print(json.dumps({"workload": f"XarmPDPickAndPlace-v0, {n} envs, step + her.add + her.sample({batch}) per step", "steps": steps,
                  "ms_per_step_with_her": ms_her, "ms_per_step_env_only": ms_env, "env_steps_per_s_with_her": n / ms_her * 1e3,
                  "env_steps_per_s_env_only": n / ms_env * 1e3, "her_stats": st, "relabelled_success_rate": float((r == 1).float().mean())}))
