#!/usr/bin/env python3
"""Roofline of the on-device VecNormalize step on synthetic batches: algorithmic bytes = one read and
one write of the observation batch + reward in / out + done + returns (read, written twice), against the measured HBM copy peak.
usage: bench_vecnorm.py [num_envs] [obs_dim]   -> one JSON line"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from gym_xarm_b200 import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
O = int(sys.argv[2]) if len(sys.argv) > 2 else 29
L = _native.load()
cfg = _native.XarmVecNormConfig(num_envs=n, obs_dim=O, device=0, gamma=0.99, clip_obs=10, clip_reward=10, epsilon=1e-8, norm_obs=1, norm_reward=1, training=1, reserved=0)
h = C.c_void_p()
_native.check(L.xarm_vecnorm_create(C.byref(cfg), C.byref(h)), "create")
dev = torch.device("cuda", 0)
ring = [(torch.randn(n, O, device=dev) * 3 + 1, torch.randn(n, device=dev), (torch.rand(n, device=dev) < 0.02).to(torch.uint8)) for _ in range(6)]
out_o, out_r = torch.empty(n, O, device=dev), torch.empty(n, device=dev)
# L2: no flush kernel (its 126 MB of dirty lines would be written back during the timed launches); instead the inputs rotate
# through a ring larger than L2 (ring bytes reported below), so every step reads its batch from HBM
nring = max(6, int((256 << 20) // (n * O * 4)) + 2)
ring = [(torch.randn(n, O, device=dev) * 3 + 1, torch.randn(n, device=dev), (torch.rand(n, device=dev) < 0.02).to(torch.uint8)) for _ in range(nring)]
s = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def step(i):
    o, r, d = ring[i % nring]
    _native.check(L.xarm_vecnorm_step(h, C.c_void_p(o.data_ptr()), C.c_void_p(r.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(out_o.data_ptr()), C.c_void_p(out_r.data_ptr()), s), "step")
for i in range(5):
    step(i)
ms = []
REP = 8   # steps per event pair: the stream stays busy, so host launch latency (3 launches per step) is not what is timed
for i in range(30):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(REP):
        step(i * REP + k)
    e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1) / REP)
ms.sort()
t = ms[len(ms) // 2] * 1e-3
algo = n * O * 4 * 2 + n * (4 + 4 + 1 + 4 * 3)
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
print(json.dumps({"op": "xarm_vecnorm_step", "num_envs": n, "obs_dim": O, "ms": t * 1e3, "algorithmic_bytes": algo, "l2": "inputs rotate through a ring of %d batches (%.0f MB > 126 MB L2), no flush kernel" % (nring, nring * n * O * 4 / 1e6),
                  "roofline": {"bound": "hbm", "achieved": algo / t / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / t / 1e9 / peak,
                               "note": "3 launches (moments, finalize, apply); the moments pass reads the batch from HBM, the apply pass re-reads it (L2 when it fits) and writes it"}}))
L.xarm_vecnorm_destroy(h)
