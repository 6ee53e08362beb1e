"""Single-env facades with the reference's class names and gym.GoalEnv-shaped API (dict obs, float reward, bool done,
info dict), each a 1-env XarmVecEnv underneath.  They exist so that `gym.make(id, config=...)`-style callers and the
reference's smoke script [REF test.py:16-28] run unchanged; throughput callers use XarmVecEnv directly."""
import numpy as np

from .vec_env import XarmVecEnv


class _SingleEnv:
    _task = None

    def __init__(self, config=None, device="cuda:0", seed=0, env_index=0):
        self.config = dict(config or {})
        self._vec = XarmVecEnv(self._task, 1, config=config, device=device, seed=seed, env_index_base=env_index,
                               auto_reset=False, output="numpy", use_graph=False)
        self.action_space = self._vec.action_space
        self.observation_space = self._vec.observation_space
        self._max_episode_steps = self._vec._max_episode_steps
        self.distance_threshold = self._vec.distance_threshold
        self.reward_type = self._vec.reward_type
        self.metadata = {"render.modes": []}
        self.goal = None
        self._needs_reset = True  # D15: stepping before reset() starts an episode instead of raising

    def seed(self, seed=None):
        return [seed]

    @staticmethod
    def _row(obs):
        return {k: v[0].copy() for k, v in obs.items()}

    def reset(self):
        obs = self._row(self._vec.reset())
        self.goal = obs["desired_goal"].copy()
        self._needs_reset = False
        return obs

    def step(self, action):
        action = np.asarray(action, np.float32)
        assert action.shape == (self._vec.act_dim,), "action shape error"  # [REF xarm_pick_and_place.py:200]
        if self._needs_reset:
            self.reset()
        obs, r, d, infos = self._vec.step(action[None])
        return self._row(obs), float(r[0]), bool(d[0]), dict(infos[0])

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        return self._vec.compute_reward(achieved_goal, desired_goal, info)

    def _is_success(self, achieved_goal, desired_goal):
        ag = np.asarray(achieved_goal, np.float32)
        dg = np.asarray(desired_goal, np.float32)
        return np.float32(np.linalg.norm(ag - dg, axis=-1) < np.float32(self.distance_threshold))

    def render(self, *a, **k):
        raise NotImplementedError("rendering is outside the batched step path")

    def close(self):
        self._vec.close()

    @property
    def unwrapped(self):
        return self


class XarmReachEnv(_SingleEnv):          # [REF gym_xarm/envs/xarm_reach.py:9]
    _task = "reach"


class XarmPickAndPlace(_SingleEnv):      # [REF gym_xarm/envs/xarm_pick_and_place.py:16]
    _task = "pick_and_place"


class XarmStackTowerEnv(_SingleEnv):     # [REF gym_xarm/envs/xarm_stack_tower.py:13]
    _task = "stack_tower"


class XarmPushWithDoorEnv(_SingleEnv):   # [REF gym_xarm/envs/xarm_push_with_door.py:13]
    _task = "push_with_door"


class XarmHandover(_SingleEnv):          # [REF gym_xarm/envs/xarm_handover.py:19]
    _task = "handover"

    def ezpolicy(self, obs):             # [REF gym_xarm/envs/xarm_handover.py:404-446] the reference's scripted handover
        from .policies import ezpolicy
        return ezpolicy(obs)


ENV_CLASSES = {c._task: c for c in (XarmReachEnv, XarmPickAndPlace, XarmStackTowerEnv, XarmPushWithDoorEnv, XarmHandover)}
