import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_xarm_b200 import XarmVecEnv, _native
n = 256
cfg = {"init_grasp_rate": 1.0, "goal_shape": "air"}
env = XarmVecEnv("pick_and_place", n, device="cuda:0", seed=2, config=cfg)
env.reset()
L = _native.load()
rng = np.random.default_rng(0)
for t in range(1):
    a = torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda()
    env.step(a)
buf = np.zeros((48, 2048), np.float32)
cnt = L.xarm_debug_mismatch(buf.ctypes.data_as(C.c_void_p))
print("mismatch events", cnt)
np.save("gpurun_out/mm_dump.npy", buf[:min(cnt, 48)])
