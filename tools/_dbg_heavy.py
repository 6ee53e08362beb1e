import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_xarm_b200 import XarmVecEnv
n = int(os.environ.get("N", 256)); steps = int(os.environ.get("STEPS", 3))
cfg = {"init_grasp_rate": 1.0, "goal_shape": "air"}
envs = [XarmVecEnv("pick_and_place", n, device="cuda:0", seed=2, config=cfg) for _ in range(2)]
if os.environ.get("GRAPH"): envs[1].capture_graph()
for e in envs: e.reset()
rng = np.random.default_rng(0)
for t in range(steps):
    a = torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda()
    o0 = envs[0].step(a)[0]["observation"].clone(); o1 = envs[1].step(a)[0]["observation"].clone()
    d = (o0 - o1).abs().max(dim=1).values
    print(t, "mismatching envs", int((d > 0).sum()), "max", float(d.max()))
