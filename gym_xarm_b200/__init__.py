"""gym_xarm_b200 - B200-native batched implementation of the gym-xarm environment step.

The package holds only what the hot path needs: csrc/ (hand-written sm_100a kernels + the C ABI of include/xarm_abi.h)
and the host-side mirror of the reference's env interface (XarmVecEnv, the five env classes, the registration IDs).
"""
from .specs import SPECS, TaskSpec  # noqa: F401
from .vec_env import XarmVecEnv, InfoList  # noqa: F401
from .envs import XarmReachEnv, XarmPickAndPlace, XarmStackTowerEnv, XarmPushWithDoorEnv, XarmHandover  # noqa: F401
from .registration import REGISTRY, make, make_vec, register_with_gym  # noqa: F401
from . import distributed  # noqa: F401
from .policies import ezpolicy  # noqa: F401
from .vec_normalize import XarmVecNormalize, XarmVecExtractDictObs  # noqa: F401
from .her import XarmHerReplayBuffer  # noqa: F401

register_with_gym()
__version__ = "0.1.0"
