"""Hindsight experience replay on the device.

The reference's sparse-reward training arm stores transitions in stable-baselines3's
`HerReplayBuffer(n_sampled_goal=4, goal_selection_strategy="future", max_episode_length=100, online_sampling=True)`
[REF benchmark/train.py:81-97], which relabels sampled transitions with goals achieved later in the same episode and
recomputes their rewards through `env.compute_reward` on batches.  `XarmHerReplayBuffer` is that buffer for an
`XarmVecEnv`: episodes live in HBM (time-major rings, one per env), `add()` consumes the buffers the env step just wrote,
`sample()` draws indices (Philox), gathers, relabels and recomputes rewards in two launches (xarm_her_sample,
include/xarm_abi.h).  Nothing visits the host.  No CPU fallback: without the CUDA library this module raises.
"""
import ctypes as C
from collections import namedtuple

import torch

from . import _native

# the field names of SB3's DictReplayBufferSamples
DictReplayBufferSamples = namedtuple("DictReplayBufferSamples", "observations actions next_observations dones rewards")


class XarmHerReplayBuffer:
    def __init__(self, venv=None, episodes_per_env=4, n_sampled_goal=4, goal_selection_strategy="future", max_episode_length=None,
                 online_sampling=True, seed=0, *, num_envs=None, obs_dim=None, goal_dim=None, action_dim=None, task=None,
                 reward_type=0, num_obj=1, device=None):
        if goal_selection_strategy != "future":
            raise NotImplementedError("only the 'future' strategy is built (the one the reference configures)")  # [REF train.py:86]
        if not online_sampling:
            raise NotImplementedError("offline sampling (relabelling at store time) is not built; the reference uses online_sampling=True")
        if venv is not None:
            num_envs, obs_dim, goal_dim, action_dim = venv.num_envs, venv.obs_dim, venv.goal_dim, venv.act_dim
            task, reward_type, num_obj, device = venv._cfg.task, venv._cfg.reward_type, max(venv.num_obj, 1), venv.device
            max_episode_length = max_episode_length or venv._max_episode_steps
        self.venv = venv
        self.device = torch.device(device if device is not None else "cuda:0")
        self.num_envs, self.obs_dim, self.goal_dim, self.action_dim = int(num_envs), int(obs_dim), int(goal_dim), int(action_dim)
        self.episodes_per_env, self.max_episode_length, self.n_sampled_goal = int(episodes_per_env), int(max_episode_length), int(n_sampled_goal)
        self.her_ratio = 1 - (1.0 / (self.n_sampled_goal + 1))
        self._lib = _native.load()
        cfg = _native.XarmHerConfig(num_envs=self.num_envs, episodes_per_env=self.episodes_per_env, max_episode_length=self.max_episode_length,
                                    obs_dim=self.obs_dim, goal_dim=self.goal_dim, action_dim=self.action_dim, task=int(task),
                                    reward_type=int(reward_type), num_obj=int(num_obj), n_sampled_goal=self.n_sampled_goal,
                                    device=self.device.index or 0, seed=int(seed))
        self._h = C.c_void_p()
        _native.check(self._lib.xarm_her_create(C.byref(cfg), C.byref(self._h)), "xarm_her_create")
        self._out = {}

    def close(self):
        if getattr(self, "_h", None) and self._h:
            torch.cuda.synchronize(self.device)
            self._lib.xarm_her_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    # ------------------------------------------------------------------ store
    def begin(self, obs=None, mask=None):
        """Call after venv.reset(): the first observation and the goal of the episode every (masked) env starts."""
        obs = obs if obs is not None else self.venv.obs_buf
        if mask is not None:
            mask = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        _native.check(self._lib.xarm_her_begin(self._h, self._p(obs["observation"]), self._p(obs["achieved_goal"]),
                                               self._p(obs["desired_goal"]), self._p(mask), self._stream()), "xarm_her_begin")

    def add(self, obs=None, action=None, reward=None, done=None, truncated=None, terminal=None):
        """HerReplayBuffer.add for the transition the env just stepped.  With no arguments the env's own buffers are read
        (obs_buf, actions, reward_buf, done_buf, truncated_buf, terminal_buf): call right after `venv.step`."""
        v = self.venv
        if obs is None:
            obs, action, reward, done, truncated, terminal = v.obs_buf, v.actions, v.reward_buf, v.done_buf, v.truncated_buf, v.terminal_buf
        for t in (obs["observation"], obs["achieved_goal"], obs["desired_goal"], action, reward):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.device == self.device
        assert done.dtype == torch.uint8 and (truncated is None or truncated.dtype == torch.uint8)
        _native.check(self._lib.xarm_her_add(self._h, self._p(obs["observation"]), self._p(obs["achieved_goal"]), self._p(obs["desired_goal"]),
                                             self._p(terminal), self._p(action), self._p(reward), self._p(done), self._p(truncated),
                                             self._stream()), "xarm_her_add")

    # ------------------------------------------------------------------ sample
    def _buffers(self, batch):
        if self._out.get("batch") != batch:
            f = dict(dtype=torch.float32, device=self.device)
            O, G, A = self.obs_dim, self.goal_dim, self.action_dim
            self._out = dict(batch=batch, obs=torch.empty(batch, O, **f), ag=torch.empty(batch, G, **f), dg=torch.empty(batch, G, **f),
                             act=torch.empty(batch, A, **f), nobs=torch.empty(batch, O, **f), nag=torch.empty(batch, G, **f),
                             rew=torch.empty(batch, **f), done=torch.empty(batch, dtype=torch.uint8, device=self.device),
                             index=torch.empty(batch, 4, dtype=torch.int32, device=self.device))
        return self._out

    def sample(self, batch_size, env=None, return_index=False):
        """HerReplayBuffer.sample(batch_size): device tensors, valid until the next call (the output buffers are reused)."""
        o = self._buffers(int(batch_size))
        _native.check(self._lib.xarm_her_sample(self._h, int(batch_size), self._p(o["obs"]), self._p(o["ag"]), self._p(o["dg"]), self._p(o["act"]),
                                                self._p(o["nobs"]), self._p(o["nag"]), self._p(o["rew"]), self._p(o["done"]), self._p(o["index"]),
                                                self._stream()), "xarm_her_sample")
        s = DictReplayBufferSamples(
            observations={"observation": o["obs"], "achieved_goal": o["ag"], "desired_goal": o["dg"]}, actions=o["act"],
            next_observations={"observation": o["nobs"], "achieved_goal": o["nag"], "desired_goal": o["dg"]},
            dones=o["done"].unsqueeze(1), rewards=o["rew"].unsqueeze(1))
        return (s, o["index"]) if return_index else s

    def stats(self):
        """{invalid_samples (since the last call), episodes, transitions, sample_calls}; synchronises."""
        out = (C.c_int64 * 4)()
        _native.check(self._lib.xarm_her_stats(self._h, out), "xarm_her_stats")
        return dict(zip(("invalid_samples", "episodes", "transitions", "sample_calls"), [int(x) for x in out]))

    def size(self):
        return self.stats()["transitions"]
