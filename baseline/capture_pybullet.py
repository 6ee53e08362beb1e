#!/usr/bin/env python3
"""Dump golden traces from the REAL reference envs (PyBullet DIRECT) - the Bullet-level pin the oracle still lacks.

    python baseline/capture_pybullet.py [--tasks reach pick_and_place ...] [--episodes 4] [--out tests/golden]

Needs `pybullet`, `gym` and the reference package with its URDFs (/root/reference, or baseline/_ref plus the asset
directories).  None of that exists in the build image or on the GPU boxes (no wheel, no network), so this script is the
HOOK for the day a box has PyBullet: it writes tests/golden/pybullet_<task>.npz, and tests/test_pybullet_traces.py - which
skips while the files are absent - replays every trace through the oracle from the captured simulator state and checks the
recalled Bullet constants of include/xarm_constants.h (joint / object poses 1e-3 over the contact-free prefix, rewards and
flags exactly).  Per episode it stores:
  q0, qd0      joint positions / velocities of every arm right after env.reset() (PyBullet joint order)
  qt0          the last POSITION_CONTROL target of every joint (NaN = never commanded)
  obj0         per object: position(3) quaternion xyzw(4) linear velocity(3) angular velocity(3)
  door0        door joint position / velocity (PushWithDoor)
  goal         desired goal
  actions      [T, A] float32 U(-1, 1)
  obs, ag, dg  [T + 1, ...] observation dicts (row 0 = after reset), reward [T], done [T], success [T]
  q, qd, obj   [T, ...] simulator state after every step
  contacts     [T] number of contact points that involve a gripper link (finger1 / finger2 / hand) after every step
"""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline.run_cpu_baseline import REF_CLASSES, pybullet_available, reference_config  # noqa: E402

LIMIT = {"reach": 25, "pick_and_place": 50, "stack_tower": 50, "push_with_door": 50, "handover": 100}


def client_of(env, mod):
    return getattr(env, "_p", None) or mod.p


def arms_of(env):
    return [env.xarm] if hasattr(env, "xarm") else [env.xarm_1, env.xarm_2]


def capture(task, episodes, seed):
    mod_name, cls = REF_CLASSES[task]
    mod = importlib.import_module(mod_name)
    cfg = reference_config(task, {"reward_type": "sparse"})
    env = getattr(mod, cls)(cfg)
    p = client_of(env, mod)
    targets = {}
    real_motor = p.setJointMotorControl2

    def motor(body, joint, mode, targetPosition=None, *a, **k):   # remember the motor targets: they are simulator state
        if targetPosition is not None:
            targets[(body, joint)] = float(np.squeeze(targetPosition))
        return real_motor(body, joint, mode, targetPosition, *a, **k)
    p.setJointMotorControl2 = motor
    arms = arms_of(env)
    nj = p.getNumJoints(arms[0])
    legos = list(getattr(env, "legos", []))
    door = getattr(env, "door", None)

    def joints():
        js = [p.getJointStates(a, list(range(nj))) for a in arms]
        return np.array([[s[0] for s in j] for j in js]), np.array([[s[1] for s in j] for j in js])

    def objects():
        rows = []
        for b in legos:
            pos, orn = p.getBasePositionAndOrientation(b)
            v, w = p.getBaseVelocity(b)
            rows.append(list(pos) + list(orn) + list(v) + list(w))
        return np.array(rows).reshape(len(legos), 13)

    def gripper_contacts():
        n = 0
        for a in arms:
            for link in (9, 10, 11):
                n += len(p.getContactPoints(bodyA=a, linkIndexA=link))
        return n

    rng = np.random.default_rng(seed)
    out = {"task": np.array(task), "episodes": np.array(episodes), "num_joints": np.array(nj)}
    for ep in range(episodes):
        o = env.reset()
        T = LIMIT[task]
        q0, qd0 = joints()
        rec = {"q0": q0, "qd0": qd0, "obj0": objects(), "goal": np.asarray(env.goal, np.float64).reshape(-1),
               "qt0": np.array([[targets.get((a, j), np.nan) for j in range(nj)] for a in arms]),
               "door0": np.array(p.getJointState(door, 0)[:2]) if door is not None else np.zeros(2)}
        A = env.action_space.shape[0]
        acts = rng.uniform(-1, 1, (T, A)).astype(np.float32)
        obs, ag, dg, rew, done, suc, qs, qds, objs, nct = [o["observation"]], [o["achieved_goal"]], [o["desired_goal"]], [], [], [], [], [], [], []
        for t in range(T):
            o, r, d, info = env.step(acts[t])
            obs.append(o["observation"]); ag.append(o["achieved_goal"]); dg.append(o["desired_goal"])
            rew.append(float(np.squeeze(r))); done.append(bool(np.squeeze(d))); suc.append(float(np.squeeze(info["is_success"])))
            q, qd = joints()
            qs.append(q); qds.append(qd); objs.append(objects()); nct.append(gripper_contacts())
            if done[-1]:
                acts = acts[:t + 1]
                break
        rec.update(actions=acts, obs=np.array(obs, np.float64), ag=np.array(ag, np.float64).reshape(len(ag), -1),
                   dg=np.array(dg, np.float64).reshape(len(dg), -1), reward=np.array(rew), done=np.array(done), success=np.array(suc),
                   q=np.array(qs), qd=np.array(qds), obj=np.array(objs), contacts=np.array(nct))
        for k, v in rec.items():
            out[f"ep{ep}_{k}"] = v
    try:
        import pybullet
        out["pybullet_api_version"] = np.array(pybullet.getAPIVersion())
    except Exception:  # noqa: BLE001
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tasks", nargs="*", default=list(REF_CLASSES))
    ap.add_argument("--episodes", type=int, default=4)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    ok, why = pybullet_available()
    if not ok:
        print(f"capture_pybullet: cannot run here - {why}")
        return 2
    for task in a.tasks:
        path = os.path.join(a.out, f"pybullet_{task}.npz")
        np.savez_compressed(path, **capture(task, a.episodes, a.seed))
        print("wrote", path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
