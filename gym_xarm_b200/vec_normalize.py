"""VecExtractDictObs + VecNormalize of the reference's training script, on the device.

The reference trains `XarmPDHandoverNoGoal-v1` (the flat `obs['observation']` of the Handover env) through
`VecNormalize(env, norm_obs=True, norm_reward=True, clip_obs=10.)` [REF benchmark/train.py:44-62,74-75].  `XarmVecNormalize`
wraps an `XarmVecEnv` the same way: `reset()` / `step(actions)` return the normalised flat observation and the normalised
reward as device tensors; the running statistics (float64) live on the device and are updated by three small kernels
(xarm_vecnorm_step, include/xarm_abi.h) right after the env step, so neither the observations nor the statistics visit the
host.  Semantics: stable-baselines3 1.x, pinned by the reference's saved `vec_normalize.pkl` (see include/xarm_abi.h and the CPU test suite).
"""
import ctypes as C
from collections import namedtuple

import numpy as np
import torch

from . import _native

RunningMeanStd = namedtuple("RunningMeanStd", "mean var count")


class XarmVecExtractDictObs:
    """The reference's VecExtractDictObs(venv, key) [REF benchmark/train.py:44-62]: a VecEnv whose observation is
    `obs[key]` (the flat Box the 'NoGoal' env id exposes).  No copy: the env's own device buffer is returned."""

    def __init__(self, venv, key="observation"):
        self.venv, self.key = venv, key
        self.observation_space = venv.observation_space.spaces[key]

    def __getattr__(self, name):
        return getattr(self.venv, name)

    def reset(self, *a, **k):
        return self.venv.reset(*a, **k)[self.key]

    def step_wait(self):
        obs, reward, done, info = self.venv.step_wait()
        return obs[self.key], reward, done, self._infos(info)

    def step(self, actions):
        obs, reward, done, info = self.venv.step(actions)
        return obs[self.key], reward, done, self._infos(info)

    def _infos(self, info):
        # the env this wrapper stands for ('...NoGoal') has a flat observation: so is its terminal_observation
        return info.view(key=self.key) if hasattr(info, "view") else info


class XarmVecNormalize:
    def __init__(self, venv, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99,
                 epsilon=1e-8, key="observation"):
        self._flat = isinstance(venv, XarmVecExtractDictObs)
        if self._flat:   # make_vec('XarmPDHandoverNoGoal-v1'): normalise the key it extracts
            venv, key = venv.venv, venv.key
        if getattr(venv, "output", "torch") != "torch":
            raise ValueError("XarmVecNormalize works on device tensors: construct the env with output='torch'")
        self.venv, self.key = venv, key
        self.num_envs, self.device = venv.num_envs, venv.device
        self.obs_dim = int(venv.obs_buf[key].shape[1])
        self.norm_obs, self.norm_reward, self.clip_obs, self.clip_reward = norm_obs, norm_reward, clip_obs, clip_reward
        self.gamma, self.epsilon = gamma, epsilon
        self._training = bool(training)
        self._lib = _native.load()
        cfg = _native.XarmVecNormConfig(num_envs=self.num_envs, obs_dim=self.obs_dim, device=self.device.index or 0, gamma=gamma,
                                        clip_obs=clip_obs, clip_reward=clip_reward, epsilon=epsilon, norm_obs=int(norm_obs),
                                        norm_reward=int(norm_reward), training=int(training), reserved=0)  # noqa: E501
        self._h = C.c_void_p()
        _native.check(self._lib.xarm_vecnorm_create(C.byref(cfg), C.byref(self._h)), "xarm_vecnorm_create")
        self.obs_out = torch.empty(self.num_envs, self.obs_dim, device=self.device)
        self.reward_out = torch.empty(self.num_envs, device=self.device)
        self.terminal_out = torch.zeros(self.num_envs, self.obs_dim, device=self.device)   # normalised terminal observations
        self.old_obs = self.old_reward = None

    # VecEnvWrapper surface
    @property
    def training(self):
        return self._training

    @training.setter
    def training(self, on):
        self._training = bool(on)
        _native.check(self._lib.xarm_vecnorm_set_training(self._h, int(bool(on))), "xarm_vecnorm_set_training")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self):
        obs = self.venv.reset()[self.key]
        self.old_obs = obs
        _native.check(self._lib.xarm_vecnorm_reset(self._h, C.c_void_p(obs.data_ptr()), C.c_void_p(self.obs_out.data_ptr()), self._stream()), "xarm_vecnorm_reset")
        return self.obs_out

    def step(self, actions):
        obs, rew, done, infos = self.venv.step(actions)
        o = obs[self.key]
        self.old_obs, self.old_reward = o, rew
        d8 = done.view(torch.uint8) if done.dtype == torch.bool else done
        _native.check(self._lib.xarm_vecnorm_step(self._h, C.c_void_p(o.data_ptr()), C.c_void_p(rew.data_ptr()), C.c_void_p(d8.data_ptr()),
                                                 C.c_void_p(self.obs_out.data_ptr()), C.c_void_p(self.reward_out.data_ptr()), self._stream()), "xarm_vecnorm_step")
        # SB3's VecNormalize.step_wait also normalises infos[i]['terminal_observation'] (with the statistics just updated)
        if self.key == "observation" and hasattr(infos, "view"):
            term = self.venv.terminal_buf
            _native.check(self._lib.xarm_vecnorm_normalize_obs(self._h, C.c_void_p(term.data_ptr()), self.num_envs, term.shape[1],
                                                              C.c_void_p(self.terminal_out.data_ptr()), self.obs_dim, 0, self._stream()), "xarm_vecnorm_normalize_obs")
            infos = infos.view(key=self.key if self._flat else None, terminal_key_obs=self.terminal_out)
        return self.obs_out, self.reward_out, done, infos

    def normalize_obs(self, obs):
        """VecNormalize.normalize_obs on any device batch [n, obs_dim]: apply only (statistics and returns untouched)."""
        return self._apply_obs(obs, 0)

    def unnormalize_obs(self, obs):
        return self._apply_obs(obs, 1)

    def _apply_obs(self, obs, inverse):
        obs = torch.as_tensor(obs, device=self.device, dtype=torch.float32)
        single = obs.dim() == 1
        o2 = obs.reshape(-1, obs.shape[-1]).contiguous()
        if o2.shape[1] != self.obs_dim:
            raise ValueError(f"expected [n, {self.obs_dim}] observations, got {tuple(obs.shape)}")
        out = torch.empty_like(o2)
        _native.check(self._lib.xarm_vecnorm_normalize_obs(self._h, C.c_void_p(o2.data_ptr()), o2.shape[0], self.obs_dim, C.c_void_p(out.data_ptr()),
                                                          self.obs_dim, inverse, self._stream()), "xarm_vecnorm_normalize_obs")
        return out[0] if single else out.reshape(obs.shape)

    def normalize_reward(self, reward):
        """VecNormalize.normalize_reward on any device batch [n]: apply only."""
        r = torch.as_tensor(reward, device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(r)
        _native.check(self._lib.xarm_vecnorm_normalize_reward(self._h, C.c_void_p(r.data_ptr()), r.numel(), C.c_void_p(out.data_ptr()), self._stream()),
                      "xarm_vecnorm_normalize_reward")
        return out

    def normalize(self, obs, reward, done=None):
        """normalize_obs + normalize_reward of a device batch of any size (replay data, evaluation): apply only."""
        return self.normalize_obs(obs), self.normalize_reward(reward)

    def get_original_obs(self):
        return self.old_obs

    def get_original_reward(self):
        return self.old_reward

    def _stats(self):
        m, v = (C.c_double * self.obs_dim)(), (C.c_double * self.obs_dim)()
        cnt, r3 = C.c_double(), (C.c_double * 3)()
        _native.check(self._lib.xarm_vecnorm_get_stats(self._h, m, v, C.byref(cnt), r3), "xarm_vecnorm_get_stats")
        return RunningMeanStd(np.array(m[:]), np.array(v[:]), cnt.value), RunningMeanStd(r3[0], r3[1], r3[2])

    @property
    def obs_rms(self):
        return self._stats()[0]

    @property
    def ret_rms(self):
        return self._stats()[1]

    def set_stats(self, obs_rms, ret_rms):
        m = (C.c_double * self.obs_dim)(*np.asarray(obs_rms.mean, np.float64))
        v = (C.c_double * self.obs_dim)(*np.asarray(obs_rms.var, np.float64))
        r3 = (C.c_double * 3)(float(ret_rms.mean), float(ret_rms.var), float(ret_rms.count))
        _native.check(self._lib.xarm_vecnorm_set_stats(self._h, m, v, float(obs_rms.count), r3), "xarm_vecnorm_set_stats")

    def save(self, path):
        """VecNormalize.save [REF benchmark/train.py:107-108] (npz instead of a pickle)."""
        o, r = self._stats()
        np.savez(path, obs_mean=o.mean, obs_var=o.var, obs_count=o.count, ret=np.array([r.mean, r.var, r.count]),
                 cfg=np.array([self.gamma, self.clip_obs, self.clip_reward, self.epsilon]))

    def load(self, path):
        d = np.load(path)
        self.set_stats(RunningMeanStd(d["obs_mean"], d["obs_var"], float(d["obs_count"])), RunningMeanStd(*d["ret"]))

    def close(self):
        if getattr(self, "_h", None) and self._h:
            torch.cuda.synchronize(self.device)
            self._lib.xarm_vecnorm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
