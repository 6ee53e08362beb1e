#!/usr/bin/env python3
"""CPU baseline of the gym-xarm env step: the reference's own path when it can run, else the
oracle port.

    python baseline/run_cpu_baseline.py [--task pick_and_place] [--seconds 5] [--min-steps 2000]

1. `import pybullet, gym` (+ the reference package: baseline/_ref, installed there by
   `pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`; /root/reference itself when it exists).
   If all of it imports, the REAL env classes run in DIRECT mode, one process per host core (plain multiprocessing with
   SubprocVecEnv's step protocol: every worker steps its env, auto-resets on done), U(-1, 1) float32 actions -> kind "pybullet".
   In the build image and on the GPU boxes pybullet / gym are not installed and cannot be fetched, and the reference's
   setup.py ships no package data (its URDFs and meshes are not installed by pip), so this branch reports why it cannot run.
2. Otherwise the repo's CPU oracle (oracle/xarm_oracle.c: float64 restatement of the same pipeline), one env per host
   thread, all cores -> kind "port", labelled "oracle stand-in, not PyBullet".
Both: warm-up, then blocks of steps until >= --seconds have elapsed AND every worker has done >= --min-steps steps; the
result is total env-steps / total seconds with the core count.  Test infrastructure: only bench.py's CPU legs import it.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_CLASSES = {"reach": ("gym_xarm.envs.xarm_reach", "XarmReachEnv"), "pick_and_place": ("gym_xarm.envs.xarm_pick_and_place", "XarmPickAndPlace"),
               "stack_tower": ("gym_xarm.envs.xarm_stack_tower", "XarmStackTowerEnv"),
               "push_with_door": ("gym_xarm.envs.xarm_push_with_door", "XarmPushWithDoorEnv"), "handover": ("gym_xarm.envs.xarm_handover", "XarmHandover")}


def reference_config(task, cfg):
    """the `config` dict of the reference constructors for the bench workload [REF xarm_pick_and_place.py:353-360; xarm_handover.py:449-456]"""
    c = {"GUI": False, "reward_type": cfg.get("reward_type", "sparse")}
    if task == "pick_and_place":
        c.update(num_obj=cfg.get("num_obj", 1), goal_shape=cfg.get("goal_shape", "air"), init_grasp_rate=cfg.get("init_grasp_rate", 0.0),
                 goal_ground_rate=cfg.get("goal_ground_rate", 0.0))
    if task == "handover":
        c.update(num_obj=cfg.get("num_obj", 1), goal_shape=cfg.get("goal_shape", "ground"), same_side_rate=cfg.get("same_side_rate", 0.5), use_stand=False)
    return c


def pybullet_available():
    """(ok, why_not): pybullet + gym + the reference package importable"""
    try:
        import pybullet  # noqa: F401
        import gym  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return False, f"pybullet / gym not importable ({type(e).__name__}: {e})"
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "gym_xarm")) and p not in sys.path:
            sys.path.insert(0, p)
    try:
        import gym_xarm  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return False, f"reference package not importable ({type(e).__name__}: {e})"
    return True, ""


def _pb_worker(task, config, seed, seconds, min_steps, warmup, out):
    import importlib
    import numpy as np
    mod, cls = REF_CLASSES[task]
    env = getattr(importlib.import_module(mod), cls)(config)
    limit = {"reach": 25, "pick_and_place": 50, "stack_tower": 50, "push_with_door": 50, "handover": 100}[task]   # gym TimeLimit [REF gym_xarm/__init__.py:6-22]
    rng = np.random.default_rng(seed)
    env.reset()
    steps = t_ep = 0
    t0 = None
    while True:
        if steps == warmup:
            t0 = time.perf_counter()
        a = rng.uniform(-1, 1, env.action_space.shape).astype(np.float32)
        _, _, done, _ = env.step(a)
        steps += 1
        t_ep += 1
        if done or t_ep >= limit:
            env.reset()
            t_ep = 0
        if t0 is not None and steps - warmup >= min_steps and time.perf_counter() - t0 >= seconds:
            break
    out.put((steps - warmup, time.perf_counter() - t0))


def run_pybullet(task, cfg, cores, seconds, min_steps, warmup=200):
    q = mp.Queue()
    ps = [mp.Process(target=_pb_worker, args=(task, reference_config(task, cfg), 1000 + i, seconds, min_steps, warmup, q)) for i in range(cores)]
    for p in ps:
        p.start()
    res = [q.get() for _ in ps]
    for p in ps:
        p.join()
    total, sec = sum(r[0] for r in res), max(r[1] for r in res)
    return {"value": total / sec, "unit": "env-steps/s", "cores": cores, "kind": "pybullet", "env_steps": total, "seconds": sec,
            "sample": f"real reference env ({REF_CLASSES[task][1]}, PyBullet DIRECT), one process per core x {cores}, {total} env-steps in {sec:.1f} s after {warmup} warm-up steps per worker"}


def run_port(task, cfg, cores, seconds, min_steps, warmup=200, block=250):
    """the oracle on `cores` threads, one env per thread (the SubprocVecEnv shape), blocks of `block` steps until both bounds hold"""
    from oracle import oracle as orc
    kw = {k: v for k, v in cfg.items() if k != "use_stand"}
    orc.bench(task, cores, warmup, cores, seed=1, **kw)
    total = sec = 0.0
    blocks = 0
    while sec < seconds or blocks * block < min_steps:
        done, s = orc.bench(task, cores, block, cores, seed=blocks, **kw)
        total += done
        sec += s
        blocks += 1
    return {"value": total / sec, "unit": "env-steps/s", "cores": cores, "kind": "port", "env_steps": total, "seconds": sec,
            "sample": f"oracle stand-in, not PyBullet: {cores} envs (one per thread) x {blocks * block} steps in blocks of {block} "
                      f"(auto-reset, U(-1,1) actions), {total:.0f} env-steps in {sec:.1f} s after {warmup} warm-up steps per worker"}


def run(task, cfg, cores=None, seconds=5.0, min_steps=2000):
    cores = cores or os.cpu_count() or 1
    ok, why = pybullet_available()
    if ok:
        try:
            return run_pybullet(task, cfg, cores, seconds, min_steps)
        except Exception as e:  # noqa: BLE001
            why = f"the reference env failed to run ({type(e).__name__}: {e})"
    r = run_port(task, cfg, cores, seconds, min_steps)
    r["pybullet_unavailable"] = why
    return r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="pick_and_place", choices=list(REF_CLASSES))
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--min-steps", type=int, default=2000)
    ap.add_argument("--cores", type=int, default=0)
    a = ap.parse_args()
    base = {"reward_type": "sparse"}
    if a.task == "pick_and_place":
        base.update(num_obj=1, goal_shape="air", init_grasp_rate=0.0, goal_ground_rate=0.0)
    if a.task == "handover":
        base.update(num_obj=1, goal_shape="ground", same_side_rate=0.5)
    print(json.dumps(run(a.task, base, a.cores or None, a.seconds, a.min_steps)))
