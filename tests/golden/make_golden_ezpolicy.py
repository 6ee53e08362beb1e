#!/usr/bin/env python3
"""Generate tests/golden/reference_ezpolicy.npz: the reference's scripted Handover policy [REF xarm_handover.py:404-446]
evaluated by the REFERENCE's own code (imported with the stub modules of make_golden.py) on random observations that
straddle its branch thresholds.  Only inputs/outputs are stored.  Usage: python tests/golden/make_golden_ezpolicy.py"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_ezpolicy.npz")


def main():
    mg.install_stubs()
    sys.path.insert(0, mg.REF)
    hand = importlib.import_module("gym_xarm.envs.xarm_handover")
    rng = np.random.default_rng(20261018)
    n = 768
    obs = np.zeros((n, 29))
    obs[:, 0:3] = rng.uniform([-0.3, -0.2, 0.0], [0.3, 0.2, 0.25], (n, 3))          # lego position
    obs[:, 3:13] = rng.normal(0, 0.3, (n, 10))
    near1 = rng.random(n) < 0.5                                                      # gripper 1 near the lego in half the cases
    obs[:, 13:16] = np.where(near1[:, None], obs[:, 0:3] + rng.normal(0, 0.04, (n, 3)), rng.uniform([-0.3, -0.2, 0.1], [0.0, 0.2, 0.22], (n, 3)))
    near2 = rng.random(n) < 0.5
    obs[:, 21:24] = np.where(near2[:, None], obs[:, 0:3] + rng.normal(0, 0.04, (n, 3)), rng.uniform([0.0, -0.2, 0.1], [0.3, 0.2, 0.22], (n, 3)))
    obs[:, 16:19] = rng.normal(0, 0.2, (n, 3)); obs[:, 24:27] = rng.normal(0, 0.2, (n, 3))
    obs[:, 19] = rng.uniform(0.0, 0.5, n); obs[:, 27] = rng.uniform(0.0, 0.5, n)      # finger joints (the policy compares with 0.25)
    obs[:, 20] = rng.normal(0, 0.1, n); obs[:, 28] = rng.normal(0, 0.1, n)
    act = np.zeros((n, 8))
    for i in range(n):
        a = hand.XarmHandover.ezpolicy(None, {"observation": obs[i]})
        act[i] = np.array([float(x) for x in a])
    np.savez_compressed(OUT, obs=obs, act=act)
    print("wrote", OUT, act.shape, "branches:", (act[:, 0:3] != 0).any(1).sum(), (act[:, 4:7] != 0).any(1).sum())


if __name__ == "__main__":
    main()
