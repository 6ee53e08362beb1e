"""ctypes binding of the CPU oracle (oracle/xarm_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs - never by
gym_xarm_b200 (the product has no CPU fallback).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libxarm_oracle.so")

TASKS = {"reach": 0, "pick_and_place": 1, "stack_tower": 2, "push_with_door": 3, "handover": 4}
REWARDS = {"sparse": 0, "dense": 1, "dense_o2g": 2, "dense_diff": 3}
GOALS = {"air": 0, "ground": 1}


class XarmConfig(C.Structure):
    _fields_ = [
        ("task", C.c_int32), ("reward_type", C.c_int32), ("num_obj", C.c_int32), ("goal_shape", C.c_int32),
        ("init_grasp_rate", C.c_float), ("goal_ground_rate", C.c_float), ("same_side_rate", C.c_float),
        ("use_stand", C.c_int32), ("max_episode_steps", C.c_int32), ("auto_reset", C.c_int32),
        ("device", C.c_int32), ("stagger_phases", C.c_int32),
        ("num_envs", C.c_int64), ("env_index_base", C.c_int64), ("seed", C.c_uint64),
    ]


def build(force=False):
    src = os.path.join(_HERE, "xarm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, dp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.or_create.restype = C.c_void_p
        L.or_create.argtypes = [C.POINTER(XarmConfig), C.c_int64]
        L.or_destroy.argtypes = [C.c_void_p]
        L.or_dims.argtypes = [C.c_int32, C.c_int32] + [C.POINTER(C.c_int32)] * 4
        L.or_reset.argtypes = [C.c_void_p, fp, fp, fp]
        L.or_get_obs.argtypes = [C.c_void_p, fp, fp, fp]
        L.or_step.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, u8p, fp, u8p]
        L.or_get_state.argtypes = [C.c_void_p, fp]
        L.or_set_state.argtypes = [C.c_void_p, fp]
        L.or_get_state_d.argtypes = [C.c_void_p, dp]
        L.or_set_state_d.argtypes = [C.c_void_p, dp]
        L.or_arm_contacts.restype = C.c_long
        L.or_arm_contacts.argtypes = [C.c_void_p, C.c_int]
        L.or_flops.restype = C.c_double
        L.or_flops.argtypes = [C.c_void_p, C.c_int]
        L.or_compute_reward.argtypes = [C.c_int32, C.c_int32, C.c_int32, fp, fp, C.c_int64, fp]
        L.or_fk.argtypes = [C.c_int32, dp, dp, dp, dp]
        L.or_dynamics.argtypes = [C.c_int32, dp, dp, dp, dp, dp]
        L.or_ik.argtypes = [C.c_int32, C.c_int, dp, dp, dp]
        L.or_box_box.argtypes = [dp, dp, dp]
        L.or_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.or_debug_assemble_obs.argtypes = [C.c_int32, C.c_int32, dp, dp, dp, dp, dp, fp, fp, fp, fp]
        L.or_debug_command.argtypes = [C.c_int32, C.c_int, fp, dp, C.c_double, dp, C.POINTER(C.c_double)]
        L.or_debug_lego_clamp.argtypes = [dp, dp]
        L.or_debug_dense_reward.restype = C.c_float
        L.or_debug_dense_reward.argtypes = [C.c_int32, C.c_int32, dp, fp, fp, C.POINTER(C.c_int32)]
        L.or_batch.argtypes = [C.POINTER(C.c_void_p), C.c_int64, C.c_int, fp, fp, fp, fp, fp, u8p, fp, u8p, C.c_int]
        L.or_bench.restype = C.c_double
        L.or_bench.argtypes = [C.POINTER(XarmConfig), C.c_int64, C.c_int, C.c_int, dp]
        _lib = L
    return _lib


def make_config(task, reward_type="sparse", num_obj=1, goal_shape="air", init_grasp_rate=0.0, goal_ground_rate=0.0,
                same_side_rate=0.5, use_stand=False, max_episode_steps=0, auto_reset=1, device=0, num_envs=1,
                env_index_base=0, seed=0):
    t = TASKS[task] if isinstance(task, str) else int(task)
    if t == 2:
        num_obj = 3
    if t == 3:
        num_obj = 1
    if t == 0:
        num_obj = 0
    return XarmConfig(t, REWARDS[reward_type] if isinstance(reward_type, str) else reward_type, num_obj,
                      GOALS[goal_shape] if isinstance(goal_shape, str) else goal_shape, init_grasp_rate,
                      goal_ground_rate, same_side_rate, int(use_stand), max_episode_steps, auto_reset, device, 0,
                      num_envs, env_index_base, seed)


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def dims(task, num_obj=1):
    a, o, g, s = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    t = TASKS[task] if isinstance(task, str) else int(task)
    if lib().or_dims(t, num_obj, a, o, g, s):
        raise ValueError("bad task")
    return a.value, o.value, g.value, s.value


class OracleEnv:
    """One env of the oracle; mirrors the reference env API (reset/step/compute_reward) on float32 arrays."""

    def __init__(self, task, env_index=0, **kw):
        self.cfg = make_config(task, **kw)
        self.L = lib()
        self.h = self.L.or_create(C.byref(self.cfg), env_index)
        if not self.h:
            raise ValueError("or_create failed")
        self.act_dim, self.obs_dim, self.goal_dim, self.state_words = dims(self.cfg.task, max(self.cfg.num_obj, 0))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.or_destroy(self.h)
            self.h = None

    def _bufs(self):
        return (np.zeros(self.obs_dim, np.float32), np.zeros(self.goal_dim, np.float32), np.zeros(self.goal_dim, np.float32))

    def reset(self):
        o, a, d = self._bufs()
        self.L.or_reset(self.h, _f(o), _f(a), _f(d))
        return {"observation": o, "achieved_goal": a, "desired_goal": d}

    def get_obs(self):
        o, a, d = self._bufs()
        self.L.or_get_obs(self.h, _f(o), _f(a), _f(d))
        return {"observation": o, "achieved_goal": a, "desired_goal": d}

    def step(self, action):
        action = np.ascontiguousarray(action, np.float32)
        assert action.shape == (self.act_dim,), "action shape error"
        o, a, d = self._bufs()
        r, s = C.c_float(), C.c_float()
        done, trunc = C.c_uint8(), C.c_uint8()
        self.L.or_step(self.h, _f(action), _f(o), _f(a), _f(d), C.byref(r), C.byref(done), C.byref(s), C.byref(trunc))
        return ({"observation": o, "achieved_goal": a, "desired_goal": d}, r.value, bool(done.value),
                {"is_success": s.value, "TimeLimit.truncated": bool(trunc.value)})

    def get_state(self):
        s = np.zeros(self.state_words, np.float32)
        self.L.or_get_state(self.h, _f(s))
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float32)
        assert s.shape == (self.state_words,)
        self.L.or_set_state(self.h, _f(s))

    def get_state_d(self):
        s = np.zeros(self.state_words, np.float64)
        self.L.or_get_state_d(self.h, _d(s))
        return s

    def set_state_d(self, s):
        s = np.ascontiguousarray(s, np.float64)
        self.L.or_set_state_d(self.h, _d(s))

    def arm_contacts(self, reset=True):
        """gripper-link contact points since the last call (parity tests: contact-free steps get the strict tolerance)"""
        return self.L.or_arm_contacts(self.h, int(reset))

    def flops(self, reset=True):
        return self.L.or_flops(self.h, int(reset))


class OracleBatch:
    """n independent oracle envs (global env indices env_index_base .. +n) stepped by all host threads; same call shapes as
    the CUDA library wrapper.  For the statistics tests at >= 32 768 envs."""

    def __init__(self, task, n, env_index_base=0, threads=None, **kw):
        self.envs = [OracleEnv(task, env_index=env_index_base + i, **kw) for i in range(n)]
        self.n, self.L = n, lib()
        self.threads = threads or os.cpu_count() or 1
        e0 = self.envs[0]
        self.act_dim, self.obs_dim, self.goal_dim, self.state_words = e0.act_dim, e0.obs_dim, e0.goal_dim, e0.state_words
        self._h = (C.c_void_p * n)(*[e.h for e in self.envs])

    def _run(self, op, actions=None):
        n = self.n
        o, a, d = np.zeros((n, self.obs_dim), np.float32), np.zeros((n, self.goal_dim), np.float32), np.zeros((n, self.goal_dim), np.float32)
        r, s = np.zeros(n, np.float32), np.zeros(n, np.float32)
        dn, tr = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        u8 = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint8))
        act = _f(actions) if actions is not None else None
        self.L.or_batch(self._h, n, op, act, _f(o), _f(a), _f(d), _f(r), u8(dn), _f(s), u8(tr), self.threads)
        return {"observation": o, "achieved_goal": a, "desired_goal": d}, r, dn.astype(bool), s, tr.astype(bool)

    def reset(self):
        return self._run(0)[0]

    def get_obs(self):
        return self._run(2)[0]

    def step(self, actions):
        actions = np.ascontiguousarray(actions, np.float32)
        assert actions.shape == (self.n, self.act_dim), "action shape error"
        return self._run(1, actions)

    def get_state(self):
        return np.stack([e.get_state() for e in self.envs])

    def set_state(self, st):
        for e, s in zip(self.envs, st):
            e.set_state(s)

    def arm_contacts(self):
        return np.array([e.arm_contacts() for e in self.envs])


def compute_reward(task, reward_type, num_obj, ag, dg):
    t = TASKS[task] if isinstance(task, str) else int(task)
    ag = np.ascontiguousarray(ag, np.float32)
    dg = np.ascontiguousarray(dg, np.float32)
    _, _, g, _ = dims(t, num_obj)
    n = ag.size // g
    out = np.zeros(n, np.float32)
    lib().or_compute_reward(t, REWARDS[reward_type] if isinstance(reward_type, str) else reward_type, num_obj, _f(ag), _f(dg), n, _f(out))
    return out


def fk(task, q):
    q = np.ascontiguousarray(q, np.float64)
    p, R, hc = np.zeros(3), np.zeros(9), np.zeros(3)
    lib().or_fk(TASKS[task], _d(q), _d(p), _d(R), _d(hc))
    return p, R.reshape(3, 3), hc


def dynamics(task, q, qd, tau):
    q, qd, tau = (np.ascontiguousarray(x, np.float64) for x in (q, qd, tau))
    n = len(q)
    Minv, qdd = np.zeros(n * n), np.zeros(n)
    lib().or_dynamics(TASKS[task], _d(q), _d(qd), _d(tau), _d(Minv), _d(qdd))
    return Minv.reshape(n, n), qdd


def ik(task, arm, q, target):
    q = np.ascontiguousarray(q, np.float64)
    target = np.ascontiguousarray(target, np.float64)
    out = np.zeros_like(q)
    lib().or_ik(TASKS[task], arm, _d(q), _d(target), _d(out))
    return out


def box_box(A, B):
    """A, B: (c[3], R[3,3], h[3]) -> list of (pa, pb, n, depth)"""
    a = np.concatenate([np.asarray(A[0], float), np.asarray(A[1], float).reshape(9), np.asarray(A[2], float)])
    b = np.concatenate([np.asarray(B[0], float), np.asarray(B[1], float).reshape(9), np.asarray(B[2], float)])
    out = np.zeros(40)
    n = lib().or_box_box(_d(a), _d(b), _d(out))
    return [(out[10 * i:10 * i + 3].copy(), out[10 * i + 3:10 * i + 6].copy(), out[10 * i + 6:10 * i + 9].copy(), out[10 * i + 9]) for i in range(n)]


def assemble_obs(task, num_obj, hand_pos, hand_vel, finger_q, finger_qd, obj, goal):
    t = TASKS[task]
    a, o, g, _ = dims(t, max(num_obj, 1) if t in (1, 4) else num_obj)
    hp, hv = np.ascontiguousarray(hand_pos, np.float64), np.ascontiguousarray(hand_vel, np.float64)
    fq, fqd = np.ascontiguousarray(finger_q, np.float64), np.ascontiguousarray(finger_qd, np.float64)
    ob = np.ascontiguousarray(obj, np.float64) if obj is not None and len(obj) else np.zeros(13)
    gl = np.ascontiguousarray(goal, np.float32)
    obs, ag, dg = np.zeros(o, np.float32), np.zeros(g, np.float32), np.zeros(g, np.float32)
    rc = lib().or_debug_assemble_obs(t, num_obj, _d(hp), _d(hv), _d(fq), _d(fqd), _d(ob), _f(gl), _f(obs), _f(ag), _f(dg))
    assert rc == 0
    return {"observation": obs, "achieved_goal": ag, "desired_goal": dg}


def command(task, arm, action, eef, finger_q):
    action = np.ascontiguousarray(action, np.float32)
    eef = np.ascontiguousarray(eef, np.float64)
    tgt, g = np.zeros(3), C.c_double()
    assert lib().or_debug_command(TASKS[task], arm, _f(action), _d(eef), float(finger_q), _d(tgt), C.byref(g)) == 0
    return tgt, g.value


def lego_clamp(pos, quat):
    pos, quat = np.array(pos, np.float64), np.array(quat, np.float64)
    lib().or_debug_lego_clamp(_d(pos), _d(quat))
    return pos, quat


def dense_reward(task, num_obj, hand, ag, dg, grasp):
    """the staged dense reward of PickAndPlace / Handover from what the reference reads: link-9 COM position per arm,
    achieved / desired goal, grasp flags (PickAndPlace: after the step; Handover: as _set_action stored them)"""
    hand = np.ascontiguousarray(hand, np.float64).reshape(-1)
    ag, dg = np.ascontiguousarray(ag, np.float32), np.ascontiguousarray(dg, np.float32)
    g = (C.c_int32 * 2)(*[int(x) for x in list(grasp) + [0] * (2 - len(grasp))])
    return lib().or_debug_dense_reward(TASKS[task], num_obj, _d(hand), _f(ag), _f(dg), g)


def philox(seed, env, episode, block):
    out = (C.c_uint32 * 4)()
    lib().or_philox(seed, env, episode, block, out)
    return np.array(list(out), dtype=np.uint32)


def bench(task, n_envs, steps, n_threads, **kw):
    cfg = make_config(task, **kw)
    n = C.c_double()
    sec = lib().or_bench(C.byref(cfg), n_envs, steps, n_threads, C.byref(n))
    return n.value, sec
