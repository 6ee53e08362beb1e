"""ctypes binding of tests/hostsim/hostsim.cpp: the CUDA kernels' per-env bodies compiled for the host.

TEST HARNESS ONLY (see the header of hostsim.cpp): it lets `pytest -m "not gpu"` check the kernel logic against the
oracle where no GPU exists.  Nothing in gym_xarm_b200 imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hostsim.so")
_CSRC = os.path.join(_HERE, "..", "..", "gym_xarm_b200", "csrc")
_INC = os.path.join(_HERE, "..", "..", "include")


def build(force=False, double=False):
    global _SO
    if double:
        _SO = os.path.join(_HERE, "_hostsim_f64.so")
    deps = [os.path.join(_HERE, "hostsim.cpp")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith(".cuh")]
    deps += [os.path.join(_INC, f) for f in os.listdir(_INC)]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-Wno-narrowing"] + (["-DXARM_HOST_SIM_DOUBLE"] if double else []) + ["-o", _SO, os.path.join(_HERE, "hostsim.cpp")])
    return _SO


_lib = None
DOUBLE = bool(int(os.environ.get("XARM_HOSTSIM_DOUBLE", "0")))  # diagnostic: kernel logic in float64
FT = np.float64 if DOUBLE else np.float32
_CFT = C.c_double if DOUBLE else C.c_float


def lib():
    global _lib
    if _lib is None:
        build(double=DOUBLE)
        L = C.CDLL(_SO)
        fp, u8p, ip = C.POINTER(_CFT), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
        L.hs_create.restype = C.c_void_p
        L.hs_create.argtypes = [C.c_int] * 4 + [C.c_double] * 3 + [C.c_int] * 2 + [C.c_longlong] * 2 + [C.c_ulonglong]
        L.hs_destroy.argtypes = [C.c_void_p]
        L.hs_set_pipeline.argtypes = [C.c_void_p, C.c_int]
        L.hs_dims.argtypes = [C.c_void_p, ip, ip, ip, ip]
        L.hs_reset.argtypes = [C.c_void_p, u8p, fp, fp, fp]
        L.hs_get_obs.argtypes = [C.c_void_p, fp, fp, fp]
        L.hs_step.argtypes = [C.c_void_p, fp, fp, fp, fp, fp, u8p, fp, u8p]
        L.hs_get_state.argtypes = [C.c_void_p, fp]
        L.hs_get_forms.argtypes = [C.c_void_p, ip]
        L.hs_set_state.argtypes = [C.c_void_p, fp]
        L.hs_compute_reward.argtypes = [C.c_int, C.c_int, C.c_int, fp, fp, C.c_int64, fp]
        L.hs_box_box.argtypes = [fp, fp, fp]
        _lib = L
    return _lib


def _f(a):
    assert a.dtype == FT
    return a.ctypes.data_as(C.POINTER(_CFT))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class HostSimVec:
    """Batched env on the host build of the kernel bodies; same call shapes as the CUDA library wrapper."""

    def __init__(self, cfg, pipeline=False):
        """pipeline=False: the fused per-env bodies (body_step / body_reset); True: the same step as the CUDA path runs
        it - the kernel pipeline of xarm_pipeline.cuh, one host loop per kernel launch."""
        self.cfg = cfg
        self.L = lib()
        self.h = self.L.hs_create(cfg.task, cfg.reward_type, cfg.num_obj, cfg.goal_shape, cfg.init_grasp_rate, cfg.goal_ground_rate,
                                  cfg.same_side_rate, cfg.max_episode_steps, cfg.auto_reset, cfg.num_envs, cfg.env_index_base, cfg.seed)
        if not self.h:
            raise ValueError("hs_create failed")
        a, o, g, s = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.L.hs_dims(self.h, a, o, g, s)
        self.A, self.O, self.G, self.S = a.value, o.value, g.value, s.value
        self.n = cfg.num_envs
        self.L.hs_set_pipeline(self.h, 1 if pipeline else 0)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.hs_destroy(self.h)
            self.h = None

    def _obs(self):
        return (np.zeros((self.n, self.O), FT), np.zeros((self.n, self.G), FT), np.zeros((self.n, self.G), FT))

    def reset(self, mask=None):
        o, a, d = self._obs()
        m = None if mask is None else _u8(np.ascontiguousarray(mask, np.uint8))
        self.L.hs_reset(self.h, m, _f(o), _f(a), _f(d))
        return {"observation": o, "achieved_goal": a, "desired_goal": d}

    def get_obs(self):
        o, a, d = self._obs()
        self.L.hs_get_obs(self.h, _f(o), _f(a), _f(d))
        return {"observation": o, "achieved_goal": a, "desired_goal": d}

    def step(self, actions):
        actions = np.ascontiguousarray(actions, FT)
        assert actions.shape == (self.n, self.A)
        o, a, d = self._obs()
        r, s = np.zeros(self.n, FT), np.zeros(self.n, FT)
        done, tr = np.zeros(self.n, np.uint8), np.zeros(self.n, np.uint8)
        self.L.hs_step(self.h, _f(actions), _f(o), _f(a), _f(d), _f(r), _u8(done), _f(s), _u8(tr))
        return {"observation": o, "achieved_goal": a, "desired_goal": d}, r, done.astype(bool), s, tr.astype(bool)

    def get_state(self):
        s = np.zeros((self.n, self.S), FT)
        self.L.hs_get_state(self.h, _f(s))
        return s

    def get_forms(self):
        """1 where the setup pass of the last substep classified the env as heavy (pipeline mode)"""
        f = np.zeros(self.n, np.int32)
        self.L.hs_get_forms(self.h, f.ctypes.data_as(C.POINTER(C.c_int)))
        return f

    def set_state(self, s):
        s = np.ascontiguousarray(s, FT)
        assert s.shape == (self.n, self.S)
        self.L.hs_set_state(self.h, _f(s))


def compute_reward(task, reward_type, num_obj, ag, dg):
    ag = np.ascontiguousarray(ag, FT)
    dg = np.ascontiguousarray(dg, FT)
    n = ag.shape[0]
    out = np.zeros(n, FT)
    lib().hs_compute_reward(task, reward_type, num_obj, _f(ag), _f(dg), n, _f(out))
    return out


def box_box(A, B):
    a = np.concatenate([np.asarray(A[0], FT), np.asarray(A[1], FT).reshape(9), np.asarray(A[2], FT)]).astype(FT)
    b = np.concatenate([np.asarray(B[0], FT), np.asarray(B[1], FT).reshape(9), np.asarray(B[2], FT)]).astype(FT)
    out = np.zeros(40, FT)
    n = lib().hs_box_box(_f(a), _f(b), _f(out))
    return [(out[10 * i:10 * i + 3].copy(), out[10 * i + 3:10 * i + 6].copy(), out[10 * i + 6:10 * i + 9].copy(), out[10 * i + 9]) for i in range(n)]
