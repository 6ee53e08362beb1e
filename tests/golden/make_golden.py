#!/usr/bin/env python3
"""Generate tests/golden/reference_logic.npz by running the REFERENCE's own Python methods.

Runs only in the build container (it imports /root/reference).  The reference env classes cannot be constructed here
(every constructor connects to PyBullet, which is not installable), so this script imports the modules with stub `gym`,
`pybullet`, `pybullet_data`, `pybullet_utils` modules and calls the reference methods UNBOUND on a small fake `self`:

* compute_reward / _is_success of all five classes on float32 goal batches (what SB3's HER feeds them) and on float64;
* _get_obs of all five classes with a fake PyBullet client that returns canned joint/link/base states, which pins the
  observation layout (ordering, relative terms, the eef2grip offset);
* _set_action of all five classes with the same fake client, capturing the IK target position, the finger targets and
  (Handover) the lego clamp / re-orientation, which pins the command scaling and clipping.

Only the numeric inputs/outputs are stored; no reference code is copied.  Usage: python tests/golden/make_golden.py
"""
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_logic.npz")


# ------------------------------------------------------------------ stubs
class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        if shape is not None:
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = np.asarray(low).astype(self.dtype)
        self.high = np.asarray(high).astype(self.dtype)
        self.shape = self.low.shape


def install_stubs():
    gym = types.ModuleType("gym")
    gym.GoalEnv = object
    gym.Env = object
    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    spaces.Dict = dict
    gym.spaces = spaces
    gym.error = types.ModuleType("gym.error")
    gym.utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    gym.utils.seeding = seeding
    gym.wrappers = types.ModuleType("gym.wrappers")
    mon = types.ModuleType("gym.wrappers.monitoring")
    mon.video_recorder = types.ModuleType("gym.wrappers.monitoring.video_recorder")
    gym.wrappers.monitoring = mon
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **kw: None
    envs.registration = registration
    gym.envs = envs
    sys.modules["gym.envs"] = envs
    sys.modules["gym.envs.registration"] = registration
    for name, mod in {"gym": gym, "gym.spaces": spaces, "gym.error": gym.error, "gym.utils": gym.utils,
                      "gym.utils.seeding": seeding, "gym.wrappers": gym.wrappers, "gym.wrappers.monitoring": mon,
                      "gym.wrappers.monitoring.video_recorder": mon.video_recorder}.items():
        sys.modules[name] = mod
    pb = types.ModuleType("pybullet")
    pb.POSITION_CONTROL = 2
    pb.getQuaternionFromEuler = lambda e: quat_from_euler(e)
    sys.modules["pybullet"] = pb
    sys.modules["pybullet_data"] = types.ModuleType("pybullet_data")
    pu = types.ModuleType("pybullet_utils")
    pu.bullet_client = types.ModuleType("pybullet_utils.bullet_client")
    sys.modules["pybullet_utils"] = pu
    sys.modules["pybullet_utils.bullet_client"] = pu.bullet_client
    return pb


def quat_from_euler(e):
    r, p, y = e
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)


def euler_from_quat(q):
    x, y, z, w = q
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = np.arcsin(np.clip(2 * (w * y - z * x), -1, 1))
    yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return (roll, pitch, yaw)


class FakeBullet:
    """Canned PyBullet getters; records the setters the reference calls."""
    POSITION_CONTROL = 2

    def __init__(self, rng, n_bodies_arm, legos):
        self.rng = rng
        self.arms = {b: {"q": rng.uniform(-1, 1, 17), "qd": rng.uniform(-1, 1, 17), "eef": rng.uniform(-0.3, 0.5, 3),
                         "hand": rng.uniform(-0.3, 0.5, 3), "handv": rng.uniform(-1, 1, 3)} for b in n_bodies_arm}
        for a in self.arms.values():
            a["q"][10] = rng.uniform(0.0, 0.04)
        self.legos = {}
        for b in legos:
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            self.legos[b] = {"pos": rng.uniform(-0.4, 0.4, 3), "quat": q, "v": rng.uniform(-1, 1, 3), "w": rng.uniform(-1, 1, 3)}
        self.ik_calls, self.motor_calls, self.dyn_calls, self.reset_calls = [], [], [], []
        self.contacts = {}

    def getJointStates(self, body, idx):
        a = self.arms[body]
        return [(a["q"][i], a["qd"][i], (0,) * 6, 0.0) for i in idx]

    def getJointState(self, body, i):
        a = self.arms[body]
        return (a["q"][i], a["qd"][i], (0,) * 6, 0.0)

    def getLinkState(self, body, link, computeLinkVelocity=0):
        a = self.arms[body]
        pos = tuple(a["eef"]) if link == 8 else tuple(a["hand"])
        return (pos, (0, 0, 0, 1), (0, 0, 0), (0, 0, 0, 1), pos, (0, 0, 0, 1), tuple(a["handv"]), (0, 0, 0))

    def getBasePositionAndOrientation(self, body):
        l = self.legos[body]
        return (tuple(l["pos"]), tuple(l["quat"]))

    def getBaseVelocity(self, body):
        l = self.legos[body]
        return (tuple(l["v"]), tuple(l["w"]))

    def calculateInverseKinematics(self, body, link, pos, orn, maxNumIterations=20):
        self.ik_calls.append((body, link, np.array(pos, dtype=np.float64), tuple(orn), maxNumIterations))
        return tuple(0.1 * (i + 1) for i in range(13))

    def setJointMotorControl2(self, body, joint, mode, target, force=None):
        self.motor_calls.append((body, joint, float(np.squeeze(target)), force))

    def getContactPoints(self, a, b, link):
        return self.contacts.get((a, b, link), [])

    def changeDynamics(self, body, link, lateralFriction=None):
        self.dyn_calls.append((body, link, lateralFriction))

    def resetBasePositionAndOrientation(self, body, pos, orn):
        self.reset_calls.append((body, np.array(pos, dtype=np.float64), np.array(orn, dtype=np.float64)))

    def getEulerFromQuaternion(self, q):
        return euler_from_quat(q)

    def getQuaternionFromEuler(self, e):
        return quat_from_euler(e)


def fake_self(**kw):
    return types.SimpleNamespace(**kw)


def main():
    sys.dont_write_bytecode = True  # /root/reference is read-only
    sys.path.insert(0, REF)
    pb = install_stubs()
    reach = importlib.import_module("gym_xarm.envs.xarm_reach")
    pap = importlib.import_module("gym_xarm.envs.xarm_pick_and_place")
    stack = importlib.import_module("gym_xarm.envs.xarm_stack_tower")
    push = importlib.import_module("gym_xarm.envs.xarm_push_with_door")
    hand = importlib.import_module("gym_xarm.envs.xarm_handover")
    rng = np.random.default_rng(20261018)
    out = {}

    # ---------------- rewards / success (float32 batches, as HER calls them; and the 1-D float64 in-step form)
    def goals(n, G, thr):
        dg = rng.uniform(-0.4, 0.4, (n, G)).astype(np.float32)
        ag = (dg + rng.normal(0, 0.6 * thr / np.sqrt(G), (n, G))).astype(np.float32)
        ag[:8] = dg[:8]
        return ag, dg

    n = 4096
    ag, dg = goals(n, 3, 0.05)
    for rt in ("sparse", "dense"):
        s = fake_self(reward_type=rt, distance_threshold=0.05)
        out[f"reach_{rt}_ag"], out[f"reach_{rt}_dg"] = ag, dg
        out[f"reach_{rt}_reward"] = np.asarray(reach.XarmReachEnv.compute_reward(s, ag, dg, {}))
    s = fake_self(goal=dg, distance_threshold=0.05)
    out["reach_success"] = np.asarray(reach.XarmReachEnv._is_success(s, ag, dg))
    for rt in ("sparse", "dense_o2g"):
        s = fake_self(config={"reward_type": rt, "num_obj": 1}, distance_threshold=0.05)
        s._subgoal_distances = lambda a, b, s=s: pap.XarmPickAndPlace._subgoal_distances(s, a, b)
        out[f"pap_{rt}_reward"] = np.asarray(pap.XarmPickAndPlace.compute_reward(s, ag, dg, {}))
    out["pap_ag"], out["pap_dg"] = ag, dg
    ag9, dg9 = goals(n, 9, 0.09)
    for rt in ("sparse", "dense"):
        s = fake_self(reward_type=rt, distance_threshold=0.03 * 3)
        out[f"stack_{rt}_reward"] = np.asarray(stack.XarmStackTowerEnv.compute_reward(s, ag9, dg9, {}))
    out["stack_ag"], out["stack_dg"] = ag9, dg9
    s = fake_self(goal=dg9, distance_threshold=0.03 * 3)
    out["stack_success"] = np.asarray(stack.XarmStackTowerEnv._is_success(s, ag9, dg9))
    ag3, dg3 = goals(n, 3, 0.03)
    for rt in ("sparse", "dense"):
        s = fake_self(reward_type=rt, distance_threshold=0.03 * 1)
        out[f"push_{rt}_reward"] = np.asarray(push.XarmPushWithDoorEnv.compute_reward(s, ag3, dg3, {}))
    out["push_ag"], out["push_dg"] = ag3, dg3
    s = fake_self(reward_type="sparse", config={"num_obj": 1}, distance_threshold=0.05)
    out["handover_sparse_reward"] = np.asarray(hand.XarmHandover.compute_reward(s, ag, dg, {}))
    out["handover_ag"], out["handover_dg"] = ag, dg
    out["handover_success"] = np.array([hand.XarmHandover._is_success(s, ag[i], dg[i]) for i in range(256)])
    # 1-D float64 (the form step() uses): the -0.0 of the Stack/Push sparse reward, the +1/0 of Reach
    s = fake_self(reward_type="sparse", distance_threshold=0.09)
    out["stack_sparse_zero"] = np.asarray(stack.XarmStackTowerEnv.compute_reward(s, dg9[0].astype(np.float64), dg9[0].astype(np.float64), {}))

    # ---------------- _get_obs layouts
    def dump_bullet(prefix, fb):
        for b, a in fb.arms.items():
            for k, v in a.items():
                out[f"{prefix}_arm{b}_{k}"] = np.asarray(v)
        for b, l in fb.legos.items():
            for k, v in l.items():
                out[f"{prefix}_lego{b}_{k}"] = np.asarray(v)

    fb = FakeBullet(rng, [1], [])
    reach.p = fb
    s = fake_self(xarm=1, num_joints=17, gripper_driver_index=10, gripper_base_index=9, goal=rng.uniform(0.3, 0.4, 3).astype(np.float32))
    o = reach.XarmReachEnv._get_obs(s)
    dump_bullet("obs_reach", fb)
    out["obs_reach_goal"] = s.goal
    for k, v in o.items():
        out[f"obs_reach_{k}"] = np.asarray(v, dtype=np.float64)

    fb = FakeBullet(rng, [1], [5])
    pap.p = fb
    s = fake_self(xarm=1, num_joints=13, finger1_index=10, gripper_base_index=9, legos=[5], goal=rng.uniform(0.3, 0.4, (1, 3)).astype(np.float32))
    o = pap.XarmPickAndPlace._get_obs(s)
    dump_bullet("obs_pap", fb)
    out["obs_pap_goal"] = s.goal
    for k, v in o.items():
        out[f"obs_pap_{k}"] = np.asarray(v, dtype=np.float64)

    for name, mod, cls, nobj in (("stack", stack, "XarmStackTowerEnv", 3), ("push", push, "XarmPushWithDoorEnv", 1)):
        fb = FakeBullet(rng, [1, 2], list(range(5, 5 + nobj)))
        mod.p = fb
        s = fake_self(xarm_1=1, xarm_2=2, num_joints=13, finger1_index=10, gripper_base_index=9, legos=list(range(5, 5 + nobj)),
                      num_obj=nobj, goal=rng.uniform(-0.3, 0.3, 3 * nobj))
        o = getattr(mod, cls)._get_obs(s)
        dump_bullet(f"obs_{name}", fb)
        out[f"obs_{name}_goal"] = s.goal
        for k, v in o.items():
            out[f"obs_{name}_{k}"] = np.asarray(v, dtype=np.float64)

    fb = FakeBullet(rng, [1, 2], [5])
    s = fake_self(_p=fb, xarm_1=1, xarm_2=2, num_joints=13, finger1_index=10, gripper_base_index=9, legos=[5], config={"num_obj": 1},
                  eef2grip_offset=[0, 0, 0.088 - 0.021], goal=rng.uniform(-0.3, 0.3, 3))
    o = hand.XarmHandover._get_obs(s)
    dump_bullet("obs_handover", fb)
    out["obs_handover_goal"] = s.goal
    for k, v in o.items():
        out[f"obs_handover_{k}"] = np.asarray(v, dtype=np.float64)

    # ---------------- _set_action: IK targets, finger targets, friction switch, lego clamp
    def record(prefix, fb, action):
        out[f"{prefix}_action"] = action
        out[f"{prefix}_ik_targets"] = np.stack([c[2] for c in fb.ik_calls])
        out[f"{prefix}_ik_iters"] = np.array([c[4] for c in fb.ik_calls])
        out[f"{prefix}_motors"] = np.array([(c[0], c[1], c[2], -1 if c[3] is None else c[3]) for c in fb.motor_calls], dtype=np.float64)
        if fb.dyn_calls:
            out[f"{prefix}_friction"] = np.array([c[2] for c in fb.dyn_calls], dtype=np.float64)

    K = 16
    for trial in range(K):
        action = rng.uniform(-1.5, 1.5, 4).astype(np.float32)  # beyond [-1,1]: exercises the clip
        fb = FakeBullet(rng, [1], [])
        fb.arms[1]["eef"] = rng.uniform([0.15, -0.45, 0.15], [0.85, 0.45, 0.65])
        reach.p = fb
        s = fake_self(xarm=1, arm_eef_index=8, gripper_driver_index=10, num_joints=17, max_vel=1, max_gripper_vel=20, dt=20 / 240.,
                      n_substeps=20, pos_space=Box(low=np.array([0.2, -0.4, 0.2]), high=np.array([0.8, 0.4, 0.6])))
        a = np.clip(action, -1, 1)  # Reach clips in step() [REF xarm_reach.py:83]
        reach.XarmReachEnv._set_action(s, a)
        dump_bullet(f"act_reach{trial}", fb)
        record(f"act_reach{trial}", fb, action)

        fb = FakeBullet(rng, [1], [5])
        fb.arms[1]["eef"] = rng.uniform([0.25, -0.35, 0.1], [0.55, 0.35, 0.45])
        if trial % 2:
            fb.contacts[(1, 5, 10)] = [1]
            fb.contacts[(1, 5, 11)] = [1]
        pap.p = fb
        s = fake_self(xarm=1, arm_eef_index=8, finger1_index=10, finger2_index=11, legos=[5], max_vel=0.25, max_gripper_vel=0.08, dt=0.25,
                      n_substeps=15, action_space=Box(-1., 1., shape=(4,), dtype="float32"),
                      pos_space=Box(low=np.array([0.3, -0.3, 0.15]), high=np.array([0.5, 0.3, 0.4])), gripper_space=Box(low=0.01, high=0.04, shape=[1]))
        pap.XarmPickAndPlace._set_action(s, action)
        dump_bullet(f"act_pap{trial}", fb)
        record(f"act_pap{trial}", fb, action)

        action8 = rng.uniform(-1.0, 1.0, 8).astype(np.float32)
        fb = FakeBullet(rng, [1, 2], [5, 6, 7])
        for b in (1, 2):
            fb.arms[b]["eef"] = rng.uniform([-0.45, -0.35, 0.1], [0.45, 0.35, 0.45])
        stack.p = fb
        s = fake_self(xarm_1=1, xarm_2=2, arm_eef_index=8, finger1_index=10, finger2_index=11, max_vel=0.25, max_gripper_vel=1, dt=0.25,
                      n_substeps=15, pos_space_1=Box(low=np.array([-0.4, -0.3, 0.125]), high=np.array([0.3, 0.3, 0.4])),
                      pos_space_2=Box(low=np.array([-0.3, -0.3, 0.125]), high=np.array([0.4, 0.3, 0.4])), gripper_space=Box(low=0.021, high=0.04, shape=[1]))
        stack.XarmStackTowerEnv._set_action(s, action8)
        dump_bullet(f"act_stack{trial}", fb)
        record(f"act_stack{trial}", fb, action8)

        fb = FakeBullet(rng, [1, 2], [5])
        for b in (1, 2):
            fb.arms[b]["eef"] = rng.uniform([-0.35, -0.25, 0.05], [0.35, 0.25, 0.3])
        fb.legos[5]["pos"] = rng.uniform([-0.4, -0.3, 0.0], [0.4, 0.3, 0.2])
        if trial % 3 == 0:
            fb.contacts[(1, 5, 10)] = [1]
            fb.contacts[(1, 5, 11)] = [1]
        s = fake_self(_p=fb, xarm_1=1, xarm_2=2, arm_eef_index=8, finger1_index=10, finger2_index=11, legos=[5], config={"num_obj": 1},
                      max_vel=1.8, max_gripper_vel=1, dt=15 / 240., n_substeps=15,
                      pos_space_1=Box(low=np.array([-0.3, -0.2, 0.1]), high=np.array([0.0, 0.2, 0.22])),
                      pos_space_2=Box(low=np.array([0.0, -0.2, 0.1]), high=np.array([0.3, 0.2, 0.22])),
                      gripper_space=Box(low=0.020, high=0.04, shape=[1]), obj_space=Box(low=np.array([0.11, -0.18]), high=np.array([0.28, 0.2])))
        hand.XarmHandover._set_action(s, action8)
        dump_bullet(f"act_handover{trial}", fb)
        record(f"act_handover{trial}", fb, action8)
        out[f"act_handover{trial}_lego_pos"] = fb.reset_calls[0][1]
        out[f"act_handover{trial}_lego_quat"] = fb.reset_calls[0][2]
        out[f"act_handover{trial}_grasp"] = np.array([s.if_xarm1_grasp, s.if_xarm2_grasp])

    # ---------------- staged dense rewards (the reference's training path: benchmark/train.py sets reward_type='dense' on
    # XarmPDHandoverNoGoal-v1) [REF xarm_pick_and_place.py:166-175; xarm_handover.py:185-199].  Every stage is visited:
    # grasp flags x lego height around the 0.05 lift test x random hand / goal positions.  Timing of the contact query:
    # PickAndPlace asks getContactPoints INSIDE compute_reward (contacts as they are after the step); Handover reads
    # self.if_xarm*_grasp, which _set_action stored BEFORE the step's stepSimulation calls - the fake client's contacts are
    # switched between the two calls to pin exactly that.
    def set_contacts(fb, arm_body, lego, on):
        fb.contacts.pop((arm_body, lego, 10), None); fb.contacts.pop((arm_body, lego, 11), None)
        if on:
            fb.contacts[(arm_body, lego, 10)] = [1]; fb.contacts[(arm_body, lego, 11)] = [1]

    KD = 96
    rec = {k: [] for k in ("hand", "ag", "dg", "g_set", "g_rew", "reward")}
    for trial in range(KD):
        fb = FakeBullet(rng, [1], [5])
        fb.arms[1]["eef"] = rng.uniform([0.3, -0.3, 0.15], [0.5, 0.3, 0.4])
        fb.arms[1]["hand"] = rng.uniform([0.3, -0.3, 0.1], [0.5, 0.3, 0.45])
        pap.p = fb
        s = fake_self(xarm=1, arm_eef_index=8, finger1_index=10, finger2_index=11, gripper_base_index=9, legos=[5], max_vel=0.25,
                      max_gripper_vel=0.08, dt=0.25, n_substeps=15, action_space=Box(-1., 1., shape=(4,), dtype="float32"),
                      pos_space=Box(low=np.array([0.3, -0.3, 0.15]), high=np.array([0.5, 0.3, 0.4])), gripper_space=Box(low=0.01, high=0.04, shape=[1]),
                      config={"reward_type": "dense", "num_obj": 1}, distance_threshold=0.05, eef2grip_offset=[0, 0, 0.088 - 0.021])
        s._subgoal_distances = lambda a, b, s=s: pap.XarmPickAndPlace._subgoal_distances(s, a, b)
        g_set, g_rew = bool(trial & 1), bool(trial & 2)
        set_contacts(fb, 1, 5, g_set)
        pap.XarmPickAndPlace._set_action(s, rng.uniform(-1, 1, 4).astype(np.float32))
        set_contacts(fb, 1, 5, g_rew)   # ... the step happens here ...
        ag = rng.uniform([0.3, -0.3, 0.02], [0.5, 0.3, 0.045 if trial & 4 else 0.3])   # lego position as _get_obs returns it (float64)
        if trial < 8:
            ag = fb.arms[1]["hand"] - np.array([0, 0, 0.088 - 0.021]) + np.array([0.06, 0, 0]) + (0 if trial < 4 else 1e-4)  # d_ao = 0 / tiny
        dg = (ag + rng.normal(0, 0.05, 3)).astype(np.float32)
        r = pap.XarmPickAndPlace.compute_reward(s, ag, dg, {})
        for k, v in (("hand", fb.arms[1]["hand"]), ("ag", ag), ("dg", dg), ("g_set", g_set), ("g_rew", g_rew), ("reward", float(r))):
            rec[k].append(np.asarray(v))
    for k, v in rec.items():
        out[f"dense_pap_{k}"] = np.stack(v)

    rec = {k: [] for k in ("hand1", "hand2", "ag", "dg", "g_set", "g_rew", "reward", "raised")}
    for trial in range(KD):
        fb = FakeBullet(rng, [1, 2], [5])
        for b in (1, 2):
            fb.arms[b]["eef"] = rng.uniform([-0.35, -0.25, 0.05], [0.35, 0.25, 0.3])
            fb.arms[b]["hand"] = rng.uniform([-0.35, -0.25, 0.05], [0.35, 0.25, 0.3])
        fb.legos[5]["pos"] = rng.uniform([-0.3, -0.2, 0.0], [0.3, 0.2, 0.2])
        s = fake_self(_p=fb, xarm_1=1, xarm_2=2, arm_eef_index=8, finger1_index=10, finger2_index=11, gripper_base_index=9, legos=[5],
                      config={"num_obj": 1}, max_vel=1.8, max_gripper_vel=1, dt=15 / 240., n_substeps=15, reward_type="dense",
                      distance_threshold=0.05, eef2grip_offset=[0, 0, 0.088 - 0.021],
                      pos_space_1=Box(low=np.array([-0.3, -0.2, 0.1]), high=np.array([0.0, 0.2, 0.22])),
                      pos_space_2=Box(low=np.array([0.0, -0.2, 0.1]), high=np.array([0.3, 0.2, 0.22])),
                      gripper_space=Box(low=0.020, high=0.04, shape=[1]), obj_space=Box(low=np.array([0.11, -0.18]), high=np.array([0.28, 0.2])))
        g_set = (bool(trial & 1), bool(trial & 2))
        g_rew = (bool(trial & 4), bool(trial & 8))
        set_contacts(fb, 1, 5, g_set[0]); set_contacts(fb, 2, 5, g_set[1])
        hand.XarmHandover._set_action(s, rng.uniform(-1, 1, 8).astype(np.float32))
        assert (s.if_xarm1_grasp, s.if_xarm2_grasp) == g_set
        set_contacts(fb, 1, 5, g_rew[0]); set_contacts(fb, 2, 5, g_rew[1])   # ... the step happens here ...
        ag = rng.uniform([-0.3, -0.2, 0.02], [0.3, 0.2, 0.045 if trial & 16 else 0.25])
        dg = (ag + rng.normal(0, 0.05, 3)).astype(np.float32)
        raised = 0
        try:
            r = float(hand.XarmHandover.compute_reward(s, ag, dg, {}))
        except NameError:   # D2: the (not grasp 1, grasp 2) branch reads an undefined `d` [REF xarm_handover.py:199]
            r, raised = np.nan, 1
        for k, v in (("hand1", fb.arms[1]["hand"]), ("hand2", fb.arms[2]["hand"]), ("ag", ag), ("dg", dg), ("g_set", g_set), ("g_rew", g_rew),
                     ("reward", r), ("raised", raised)):
            rec[k].append(np.asarray(v))
    for k, v in rec.items():
        out[f"dense_handover_{k}"] = np.stack(v)
    out["n_trials"] = np.array(K)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
