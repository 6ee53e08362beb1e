"""Registration IDs: both the IDs the reference snapshot registers [REF gym_xarm/__init__.py:6-22]
and the README/north-star spellings.  Registrations carry default kwargs so that `make(id)` works without `config`
(D14).  If gym or gymnasium is importable the IDs are also registered there; `gym_xarm_b200.make` always works."""
from .envs import ENV_CLASSES
from .specs import SPECS
from .vec_env import XarmVecEnv

# id -> (task, default config overrides)
REGISTRY = {
    "XarmReach-v0": ("reach", {}),
    "XarmPickAndPlace-v1": ("pick_and_place", {}),
    "XarmPDPickAndPlace-v0": ("pick_and_place", {}),
    "XarmPDStackTower-v0": ("stack_tower", {}),
    "XarmPDPushWithDoor-v0": ("push_with_door", {}),
    "XarmHandover-v0": ("handover", {}),
    "XarmPDHandover-v0": ("handover", {}),
    "XarmPDHandover-v1": ("handover", {"reward_type": "dense"}),
}

# flat-observation ids: id -> (dict-observation id, key).  The reference trains on 'XarmPDHandoverNoGoal-v1'
# [REF benchmark/train.py:66,74] (the Handover env behind VecExtractDictObs [REF benchmark/train.py:44-62]); its snapshot
# does not register the id, so it exists here for make_vec only.
FLAT_IDS = {"XarmPDHandoverNoGoal-v1": ("XarmPDHandover-v1", "observation")}


class TimeLimit:
    """gym.wrappers.TimeLimit stand-in (a20): sets done and info['TimeLimit.truncated'] at max_episode_steps."""

    def __init__(self, env, max_episode_steps):
        self.env, self._max_episode_steps, self._elapsed = env, max_episode_steps, 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self):
        self._elapsed = 0
        return self.env.reset()

    def step(self, action):
        obs, r, d, info = self.env.step(action)
        self._elapsed += 1
        if self._elapsed >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not d or info.get("TimeLimit.truncated", False)
            d = True
        return obs, r, d, info


def make(env_id, config=None, **kwargs):
    """gym.make(id, config=...) equivalent: a single env wrapped in TimeLimit."""
    if env_id not in REGISTRY:
        raise KeyError(f"unknown env id {env_id!r}; known: {sorted(REGISTRY)}")
    task, defaults = REGISTRY[env_id]
    cfg = dict(defaults)
    cfg.update(config or {})
    env = ENV_CLASSES[task](cfg, **kwargs)
    return TimeLimit(env, SPECS[task].max_episode_steps)


def make_vec(env_id, num_envs, config=None, **kwargs):
    """Batched counterpart of SB3's make_vec_env(env_id, n_envs=...) [REF benchmark/train.py:74]."""
    if env_id in FLAT_IDS:
        from .vec_normalize import XarmVecExtractDictObs
        base, key = FLAT_IDS[env_id]
        return XarmVecExtractDictObs(make_vec(base, num_envs, config=config, **kwargs), key)
    task, defaults = REGISTRY[env_id]
    cfg = dict(defaults)
    cfg.update(config or {})
    return XarmVecEnv(task, num_envs, config=cfg, **kwargs)


def register_with_gym():
    """Register every ID with gym / gymnasium when one of them is installed (it is not in the build image)."""
    done = []
    for modname in ("gymnasium", "gym"):
        try:
            mod = __import__(modname)
            reg = mod.envs.registration.register
        except Exception:  # noqa: BLE001
            continue
        for env_id, (task, defaults) in REGISTRY.items():
            try:
                reg(id=env_id, entry_point=f"gym_xarm_b200.envs:{ENV_CLASSES[task].__name__}",
                    max_episode_steps=SPECS[task].max_episode_steps, kwargs={"config": dict(defaults)})
                done.append((modname, env_id))
            except Exception:  # noqa: BLE001  (already registered)
                pass
    return done
