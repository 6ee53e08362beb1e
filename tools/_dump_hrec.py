import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_xarm_b200 import XarmVecEnv, _native
n = 256
cfg = {"init_grasp_rate": 1.0, "goal_shape": "air"}
os.environ["XARM_NO_SPLIT"] = "1"
env = XarmVecEnv("pick_and_place", n, device="cuda:0", seed=2, config=cfg, auto_reset=False)
env.reset()
L = _native.load()
L.xarm_debug_hrec.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
rng = np.random.default_rng(0)
out = {}
for t in range(6):
    a = torch.from_numpy(rng.uniform(-1, 1, (n, 4)).astype(np.float32)).cuda()
    a[:, 3] = -1.0   # close the fingers
    env.step(a)
    buf = np.zeros(1812 * n, np.float32)
    w = L.xarm_debug_hrec(env._h, buf.ctypes.data_as(C.c_void_p), buf.size)
    out[f"step{t}"] = buf.reshape(n, w).copy()
np.savez_compressed("gpurun_out/hrec_dump.npz", **out)
print("saved", {k: v.shape for k, v in out.items()})
