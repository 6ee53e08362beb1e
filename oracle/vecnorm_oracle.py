"""ORACLE (test infrastructure, not product code): numpy float64 restatement of stable-baselines3 1.x VecNormalize +
RunningMeanStd as the reference's training script uses them [REF benchmark/train.py:74-75: VecNormalize(env, norm_obs=True,
norm_reward=True, clip_obs=10.)].  stable-baselines3 is not installable here, so the algorithm is restated from its
published source (common/vec_env/vec_normalize.py, common/running_mean_std.py of the 1.x line the reference's pickle was
written by) and PINNED by the reference's saved benchmark/saved_data/XarmPDHandoverNoGoal-v1/vec_normalize.pkl:
epsilon 1e-8, gamma 0.99, clip_obs = clip_reward = 10, RunningMeanStd.count starts at 1e-4, and after 6250 steps of 4 envs
obs_rms.count == 25000.0001 while ret_rms.count == 25004.0001 - i.e. reset() feeds the zeroed returns to ret_rms and does
not update obs_rms (tests/test_oracle_golden.py::test_vecnorm_oracle_counts_match_reference_pickle).
"""
import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean, self.var, self.count = np.zeros(shape, np.float64), np.ones(shape, np.float64), epsilon

    def update(self, arr):
        arr = np.asarray(arr, np.float64)
        self.update_from_moments(arr.mean(axis=0), arr.var(axis=0), arr.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot
        m2 = self.var * self.count + batch_var * batch_count + np.square(delta) * self.count * batch_count / tot
        self.mean, self.var, self.count = new_mean, m2 / tot, tot


class VecNormalizeOracle:
    def __init__(self, num_envs, obs_dim, training=True, norm_obs=True, norm_reward=True, clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8):
        self.obs_rms, self.ret_rms = RunningMeanStd(shape=(obs_dim,)), RunningMeanStd(shape=())
        self.ret = np.zeros(num_envs)
        self.training, self.norm_obs, self.norm_reward = training, norm_obs, norm_reward
        self.clip_obs, self.clip_reward, self.gamma, self.epsilon = clip_obs, clip_reward, gamma, epsilon

    def normalize_obs(self, obs):
        if not self.norm_obs:
            return obs
        return np.clip((obs - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.epsilon), -self.clip_obs, self.clip_obs)

    def normalize_reward(self, reward):
        if not self.norm_reward:
            return reward
        return np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)

    def _update_reward(self, reward):
        self.ret = self.ret * self.gamma + reward
        self.ret_rms.update(self.ret)

    def reset(self, obs):
        self.ret = np.zeros_like(self.ret)
        if self.training:
            self._update_reward(self.ret)
        return self.normalize_obs(obs)

    def step(self, obs, reward, done):
        if self.training:
            if self.norm_obs:
                self.obs_rms.update(obs)
            self._update_reward(reward)
        o, r = self.normalize_obs(obs), self.normalize_reward(reward)
        self.ret[np.asarray(done, bool)] = 0
        return o, r
