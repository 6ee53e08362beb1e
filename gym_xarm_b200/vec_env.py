"""XarmVecEnv - the batched, GPU-resident replacement of the reference env classes.

Host side of the drop-in boundary: the Python surface SB3 / gym callers rely on (`reset`, `step`,
`compute_reward(achieved_goal, desired_goal, info)`, `observation_space` Dict(observation / achieved_goal /
desired_goal), `action_space`, `_max_episode_steps`, `distance_threshold`, VecEnv auto-reset with
`terminal_observation`) implemented as thin calls into the C ABI of libxarm_b200.so.  PyTorch is used only for device
memory and streams.  There is no CPU path: without a CUDA device the constructor raises.

Reference methods mirrored: XarmPickAndPlace.step/reset/compute_reward [REF gym_xarm/envs/xarm_pick_and_place.py:107-190]
and the same trio of xarm_reach.py, xarm_stack_tower.py, xarm_push_with_door.py, xarm_handover.py.
"""
import ctypes as C

import numpy as np
import torch

from . import _native
from .spaces import Box, Dict
from .specs import GOAL_IDS, REWARD_IDS, SPECS, normalize_config


class InfoList:
    """Lazy `infos`: SB3 indexes a list of dicts; building 10^5 dicts per step on the host would dominate the step.
    Dicts are materialised on access from the step's host/device arrays."""

    def __init__(self, env, success, truncated, done, terminal_obs=None, key=None, terminal_key_obs=None):
        """terminal_obs: [N, O + 2 G] slab (rows of finished envs valid).  key: the observation is obs[key] (VecExtractDictObs) -
        terminal_observation is then the flat array.  terminal_key_obs: [N, O] replacement for the slab's observation part
        (VecNormalize: the normalised terminal observation)."""
        self._env, self._success, self._truncated, self._done, self._term = env, success, truncated, done, terminal_obs
        self._key, self._term_key = key, terminal_key_obs
        self._host = None

    def view(self, key=None, terminal_key_obs=None):
        """the same infos as seen through a wrapper (VecExtractDictObs / VecNormalize)"""
        return InfoList(self._env, self._success, self._truncated, self._done, self._term, key if key is not None else self._key,
                        terminal_key_obs if terminal_key_obs is not None else self._term_key)

    def __len__(self):
        return self._env.num_envs

    def _fetch(self):
        if self._host is None:
            def h(x):
                return x.cpu().numpy() if isinstance(x, torch.Tensor) else (None if x is None else np.asarray(x))
            self._host = (h(self._success), h(self._truncated), h(self._done), h(self._term), h(self._term_key))
        return self._host

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        s, t, d, term, term_key = self._fetch()
        info = {"is_success": float(s[i]), "TimeLimit.truncated": bool(t[i])}
        if d[i] and term is not None:
            O, G = self._env.obs_dim, self._env.goal_dim
            row = term[i]
            full = {"observation": row[:O].copy(), "achieved_goal": row[O:O + G].copy(), "desired_goal": row[O + G:].copy()}
            if term_key is not None:
                full[self._key or "observation"] = term_key[i].copy()
            info["terminal_observation"] = full[self._key] if self._key else full
        return info

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class XarmVecEnv:
    """N envs of one task on one GPU.

    task: 'reach' | 'pick_and_place' | 'stack_tower' | 'push_with_door' | 'handover'
    config: the reference's `config` dict (GUI, reward_type, num_obj, init_grasp_rate, goal_ground_rate, goal_shape,
            same_side_rate, use_stand); missing keys take the task defaults of specs.py
    output: 'torch' (device tensors, zero copies) or 'numpy' (host arrays through xarm_step_host - what SB3 consumes)
    env_index_base: global index of env 0 (RNG streams are keyed by the global env index, so a job sharded over
            several GPUs produces the same episodes as a single slab)
    stagger_phases: env i starts (construction, explicit reset) at step counter (global index mod episode length): the
            time-limit endings - and the auto-reset work they cause - then spread evenly over the steps instead of arriving
            as one wave every episode length.  Off by default (and in every parity test).
    """

    metadata = {"render.modes": []}

    def __init__(self, task, num_envs=1, config=None, device="cuda:0", seed=0, env_index_base=0, auto_reset=True,
                 output="torch", use_graph=True, max_episode_steps=None, stagger_phases=False):
        if task not in SPECS:
            raise ValueError(f"unknown task {task!r}")
        if not torch.cuda.is_available():
            raise _native.XarmError("XarmVecEnv needs a CUDA device: the env step runs only as sm_100a kernels (no CPU fallback)")
        self.spec_task = SPECS[task]
        self.config = normalize_config(self.spec_task, config)
        if self.config.get("GUI"):
            raise NotImplementedError("GUI/rendering is outside the batched step path")
        self.num_envs = int(num_envs)
        self.device = torch.device(device)
        self.output = output
        self.num_obj = self.config["num_obj"]
        self.act_dim, self.obs_dim, self.goal_dim = self.spec_task.dims(max(self.num_obj, 1))
        self._max_episode_steps = int(max_episode_steps or self.spec_task.max_episode_steps)
        self.distance_threshold = self.spec_task.distance_threshold
        self.reward_type = self.config["reward_type"]
        self._lib = _native.load()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._cfg = _native.XarmConfig(
            self.spec_task.task, REWARD_IDS[self.reward_type], self.num_obj, GOAL_IDS[self.config.get("goal_shape", "air")],
            float(self.config.get("init_grasp_rate", 0.0)), float(self.config.get("goal_ground_rate", 0.0)),
            float(self.config.get("same_side_rate", 0.5)), int(bool(self.config.get("use_stand", False))),
            int(max_episode_steps or 0), int(bool(auto_reset)), dev_index, int(bool(stagger_phases)), self.num_envs, int(env_index_base), int(seed))
        self._h = C.c_void_p()
        _native.check(self._lib.xarm_create(C.byref(self._cfg), C.byref(self._h)), "xarm_create")
        sw = C.c_int32()
        _native.check(self._lib.xarm_task_dims(self._cfg.task, self.num_obj, None, None, None, C.byref(sw)))
        self.state_words = sw.value
        N, A, O, G = self.num_envs, self.act_dim, self.obs_dim, self.goal_dim
        f32 = dict(dtype=torch.float32, device=self.device)
        self.actions = torch.zeros(N, A, **f32)
        self.obs_buf = {"observation": torch.zeros(N, O, **f32), "achieved_goal": torch.zeros(N, G, **f32),
                        "desired_goal": torch.zeros(N, G, **f32)}
        self.reward_buf = torch.zeros(N, **f32)
        self.done_buf = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.success_buf = torch.zeros(N, **f32)
        self.truncated_buf = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.terminal_buf = torch.zeros(N, O + 2 * G, **f32)
        self._bufs = _native.XarmBuffers(
            self.actions.data_ptr(), self.obs_buf["observation"].data_ptr(), self.obs_buf["achieved_goal"].data_ptr(),
            self.obs_buf["desired_goal"].data_ptr(), self.reward_buf.data_ptr(), self.done_buf.data_ptr(),
            self.success_buf.data_ptr(), self.truncated_buf.data_ptr(), self.terminal_buf.data_ptr())
        _native.check(self._lib.xarm_bind(self._h, C.byref(self._bufs)), "xarm_bind")
        self.action_space = Box(-1.0, 1.0, shape=(A,), dtype=np.float32)  # [REF xarm_pick_and_place.py:95]
        self.observation_space = Dict(dict(
            desired_goal=Box(-np.inf, np.inf, shape=(G,), dtype=np.float32),
            achieved_goal=Box(-np.inf, np.inf, shape=(G,), dtype=np.float32),
            observation=Box(-np.inf, np.inf, shape=(O,), dtype=np.float32)))  # [REF xarm_pick_and_place.py:96-100]
        self._stream = torch.cuda.Stream(device=self.device) if use_graph else None
        self._graph_ok = False
        self._pending = None
        self.reward_range = (-float("inf"), float("inf"))
        self.render_mode = None

    # ------------------------------------------------------------------ plumbing
    def _cur_stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None) and self._h:
            torch.cuda.synchronize(self.device)
            self._lib.xarm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def capture_graph(self):
        """Capture step(+auto-reset) as a CUDA graph on the env's own stream (xarm_graph_capture)."""
        if self._stream is None:
            raise ValueError("constructed with use_graph=False")
        torch.cuda.synchronize(self.device)
        _native.check(self._lib.xarm_graph_capture(self._h, C.c_void_p(self._stream.cuda_stream)), "xarm_graph_capture")
        self._graph_ok = True

    def set_profiling(self, on=True):
        """Device-side per-kernel timers of the step pipeline (xarm_set_profiling).  Turning them on or off drops a captured
        graph (capture again); calling it again while on clears the accumulators."""
        _native.check(self._lib.xarm_set_profiling(self._h, 1 if on else 0), "xarm_set_profiling")
        if bool(on) != getattr(self, "_profiling", False):
            self._graph_ok = False
        self._profiling = bool(on)

    def kernel_times(self):
        """{(branch, kernel): (launches, total_us)} accumulated since set_profiling (xarm_kernel_times)."""
        buf = C.create_string_buffer(1 << 16)
        n = self._lib.xarm_kernel_times(self._h, buf, len(buf))
        if n < 0:
            _native.check(n, "xarm_kernel_times")
        out = {}
        for line in buf.value.decode().splitlines():
            br, name, cnt, us = line.split()
            out[(br, name)] = (int(cnt), float(us))
        return out

    # ------------------------------------------------------------------ gym / VecEnv surface
    def seed(self, seed=None):
        return [seed]  # D11: RNG streams are fixed at construction (seed, global env index, episode)

    def _out(self, t):
        return t if self.output == "torch" else t.cpu().numpy()

    def _obs_out(self):
        return {k: self._out(v) for k, v in self.obs_buf.items()}

    def reset(self, mask=None):
        """Env.reset() for all envs (or those selected by a bool/uint8 mask tensor)."""
        if self.output == "numpy" and mask is None:
            N, O, G = self.num_envs, self.obs_dim, self.goal_dim
            o, a, d = np.empty((N, O), np.float32), np.empty((N, G), np.float32), np.empty((N, G), np.float32)
            _native.check(self._lib.xarm_reset_host(self._h, o.ctypes.data, a.ctypes.data, d.ctypes.data, None), "xarm_reset_host")
            return {"observation": o, "achieved_goal": a, "desired_goal": d}
        mp = None
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
            mp = C.c_void_p(mask.data_ptr())
        _native.check(self._lib.xarm_reset(self._h, mp, self._cur_stream()), "xarm_reset")
        return self._obs_out()

    def step_async(self, actions):
        self._pending = actions

    def step_wait(self):
        a, self._pending = self._pending, None
        return self.step(a)

    def step(self, actions):
        """Env.step for every env. actions: [N, A] float32 (device tensor for output='torch', ndarray for 'numpy')."""
        N, A = self.num_envs, self.act_dim
        if self.output == "numpy":
            actions = np.ascontiguousarray(actions, np.float32)
            assert actions.shape == (N, A), "action shape error"  # [REF xarm_pick_and_place.py:200]
            O, G = self.obs_dim, self.goal_dim
            o, ag, dg = np.empty((N, O), np.float32), np.empty((N, G), np.float32), np.empty((N, G), np.float32)
            r, s = np.empty(N, np.float32), np.empty(N, np.float32)
            d, t = np.empty(N, np.uint8), np.empty(N, np.uint8)
            term = np.zeros((N, O + 2 * G), np.float32)   # rows of the envs that finished are filled in (terminal_observation)
            _native.check(self._lib.xarm_step_host(self._h, actions.ctypes.data, o.ctypes.data, ag.ctypes.data, dg.ctypes.data,
                                                  r.ctypes.data, d.ctypes.data, s.ctypes.data, t.ctypes.data, term.ctypes.data, None),
                          "xarm_step_host")
            obs = {"observation": o, "achieved_goal": ag, "desired_goal": dg}
            return obs, r, d.astype(bool), InfoList(self, s, t, d, term)
        actions = torch.as_tensor(actions, device=self.device, dtype=torch.float32)
        assert tuple(actions.shape) == (N, A), "action shape error"
        if actions.data_ptr() != self.actions.data_ptr():
            self.actions.copy_(actions, non_blocking=True)
        if self._graph_ok:
            cur = torch.cuda.current_stream(self.device)
            self._stream.wait_stream(cur)
            _native.check(self._lib.xarm_step(self._h, C.c_void_p(self._stream.cuda_stream)), "xarm_step")
            cur.wait_stream(self._stream)
        else:
            _native.check(self._lib.xarm_step(self._h, self._cur_stream()), "xarm_step")
        infos = InfoList(self, self.success_buf, self.truncated_buf, self.done_buf, self.terminal_buf)
        return self.obs_buf, self.reward_buf, self.done_buf.bool(), infos

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        """HER entry point: compute_reward(achieved_goal, desired_goal, info), batch-safe reward types only
        [REF xarm_reach.py:107-116; xarm_pick_and_place.py:155-190; xarm_stack_tower.py:124-129; xarm_handover.py:153-183]."""
        is_np = not isinstance(achieved_goal, torch.Tensor)
        ag = torch.as_tensor(np.asarray(achieved_goal, np.float32) if is_np else achieved_goal, device=self.device, dtype=torch.float32).contiguous()
        dg = torch.as_tensor(np.asarray(desired_goal, np.float32) if is_np else desired_goal, device=self.device, dtype=torch.float32).contiguous()
        G = self.goal_dim
        single = ag.dim() == 1
        ag2, dg2 = ag.reshape(-1, G), dg.reshape(-1, G)
        if ag2.shape != dg2.shape:
            raise ValueError("achieved_goal and desired_goal shapes differ")
        out = torch.empty(ag2.shape[0], dtype=torch.float32, device=self.device)
        _native.check(self._lib.xarm_compute_reward(self._cfg.task, self._cfg.reward_type, max(self.num_obj, 1), ag2.data_ptr(),
                                                   dg2.data_ptr(), ag2.shape[0], out.data_ptr(), self._cur_stream()), "xarm_compute_reward")
        if single:
            return float(out.item()) if is_np else out[0]
        return out.cpu().numpy() if is_np else out

    def get_obs(self):
        _native.check(self._lib.xarm_get_obs(self._h, self._cur_stream()), "xarm_get_obs")
        return self._obs_out()

    # SB3 VecEnv protocol ------------------------------------------------------------------
    def get_attr(self, name, indices=None):
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [getattr(self, name)] * n

    def set_attr(self, name, value, indices=None):
        setattr(self, name, value)

    def env_method(self, method_name, *args, indices=None, **kwargs):
        res = getattr(self, method_name)(*args, **kwargs)
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [res] * n

    def env_is_wrapped(self, wrapper_class, indices=None):
        n = self.num_envs if indices is None else len(self._indices(indices))
        return [False] * n

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        return [indices] if isinstance(indices, int) else list(indices)

    def render(self, *a, **k):
        raise NotImplementedError("rendering is outside the batched step path (SURVEY.md 2.3 N12)")

    # simulator checkpoint ------------------------------------------------------------------
    def get_state(self):
        s = np.empty((self.num_envs, self.state_words), np.float32)
        _native.check(self._lib.xarm_get_state(self._h, s.ctypes.data), "xarm_get_state")
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, np.float32)
        assert s.shape == (self.num_envs, self.state_words)
        _native.check(self._lib.xarm_set_state(self._h, s.ctypes.data), "xarm_set_state")

    def episode_stats(self):
        """{episodes, return_sum, length_sum, success_sum, diverged} since the last call (this GPU's slab)."""
        out = (C.c_double * 5)()
        _native.check(self._lib.xarm_episode_stats(self._h, out, self._cur_stream()), "xarm_episode_stats")
        return dict(zip(("episodes", "return_sum", "length_sum", "success_sum", "diverged"), list(out)))
