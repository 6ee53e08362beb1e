#!/usr/bin/env python3
"""Roofline of the on-device HER buffer on synthetic transitions of PickAndPlace shape.
  add:    algorithmic bytes per env = read + write of (next_obs O + next_ag G + action A + reward 1) words + done / truncated bytes
          (+ the first row of the next episode where an episode ends: 2 x (O + 2 G) words)
  sample: algorithmic bytes per sample = read + write of (2 O + 3 G + A + 1) words + 1 byte + the 16-byte index written and read
usage: bench_her.py [num_envs] [batch]   -> one JSON line per op"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_xarm_b200.her import XarmHerReplayBuffer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
O, G, A, K, T = 24, 3, 4, 4, 50
dev = torch.device("cuda", 0)
buf = XarmHerReplayBuffer(num_envs=n, obs_dim=O, goal_dim=G, action_dim=A, task=1, reward_type=0, num_obj=1, episodes_per_env=K,
                          max_episode_length=T, n_sampled_goal=4, seed=1, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
row_bytes = (2 * O + 3 * G + A + 1) * 4 + 2
nring = max(4, int((256 << 20) // (n * row_bytes)) + 2)      # inputs rotate through a ring larger than the 126 MB L2
def batch():
    done = (torch.rand(n, generator=g, device=dev) < 0.03).to(torch.uint8)
    return dict(obs={"observation": torch.randn(n, O, generator=g, device=dev), "achieved_goal": torch.rand(n, G, generator=g, device=dev) * 0.1,
                     "desired_goal": torch.rand(n, G, generator=g, device=dev) * 0.1},
                action=torch.rand(n, A, generator=g, device=dev), reward=torch.randn(n, generator=g, device=dev), done=done,
                truncated=done & (torch.rand(n, generator=g, device=dev) < 0.5).to(torch.uint8), terminal=torch.randn(n, O + 2 * G, generator=g, device=dev))
ring = [batch() for _ in range(nring)]
buf.begin(ring[0]["obs"])
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0

QUICK = bool(os.environ.get("HER_BENCH_QUICK"))    # a short run for ncu captures


def timed(fn, reps, inner):
    if QUICK:
        reps, inner = 2, 30 if reps == 25 else 1
    for i in range(3):
        fn(i)
    ms = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(inner):
            fn(i * inner + k)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / inner)
    ms.sort()
    return ms[len(ms) // 2] * 1e-3

t_add = timed(lambda i: buf.add(**ring[i % nring]), 25, 8)     # 203 adds: every ring slot holds finished episodes afterwards
st = buf.stats()
closing = st["episodes"] / max(st["transitions"], 1)
algo_add = n * (2 * 4 * (O + G + A + 1) + 3 + closing * 2 * 4 * (O + 2 * G))
print(json.dumps({"op": "xarm_her_add", "num_envs": n, "ms": t_add * 1e3, "algorithmic_bytes": algo_add, "transitions_per_s": n / t_add,
                  "l2": "inputs rotate through a ring of %d batches (%.0f MB), no flush kernel" % (nring, nring * n * row_bytes / 1e6),
                  "roofline": {"bound": "hbm", "achieved": algo_add / t_add / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_add / t_add / 1e9 / peak,
                               "note": "2 launches (store, advance); time-major rings: an env slab writes neighbouring rows"}}))
for b in sorted({256, 65536, B}):
    t_s = timed(lambda i: buf.sample(b), 20, 8)
    algo_s = b * (2 * 4 * (2 * O + 3 * G + A + 1) + 2 + 32)
    print(json.dumps({"op": "xarm_her_sample", "num_envs": n, "batch": b, "ms": t_s * 1e3, "algorithmic_bytes": algo_s, "samples_per_s": b / t_s,
                      "store_bytes": n * K * ((T + 1) * (O + G) + T * (A + 1)) * 4,
                      "roofline": {"bound": "hbm", "achieved": algo_s / t_s / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_s / t_s / 1e9 / peak,
                                   "note": "2 launches (index, gather); gathered rows are 96 / 12 / 16-byte segments at random addresses of a store far larger than L2: sector-granular (32 B) reads"}}))
print(json.dumps({"stats": buf.stats()}))
buf.close()
