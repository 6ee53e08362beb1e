"""ctypes binding of libxarm_b200.so (include/xarm_abi.h).  There is no fallback: if the library is missing or a call
fails, this module raises."""
import ctypes as C
import os

from . import build as _build


class XarmConfig(C.Structure):
    _fields_ = [
        ("task", C.c_int32), ("reward_type", C.c_int32), ("num_obj", C.c_int32), ("goal_shape", C.c_int32),
        ("init_grasp_rate", C.c_float), ("goal_ground_rate", C.c_float), ("same_side_rate", C.c_float),
        ("use_stand", C.c_int32), ("max_episode_steps", C.c_int32), ("auto_reset", C.c_int32),
        ("device", C.c_int32), ("stagger_phases", C.c_int32),
        ("num_envs", C.c_int64), ("env_index_base", C.c_int64), ("seed", C.c_uint64),
    ]


class XarmBuffers(C.Structure):
    _fields_ = [
        ("actions", C.c_void_p), ("observation", C.c_void_p), ("achieved_goal", C.c_void_p), ("desired_goal", C.c_void_p),
        ("reward", C.c_void_p), ("done", C.c_void_p), ("success", C.c_void_p), ("truncated", C.c_void_p),
        ("terminal_observation", C.c_void_p),
    ]


class XarmVecNormConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int64), ("obs_dim", C.c_int32), ("device", C.c_int32),
        ("gamma", C.c_float), ("clip_obs", C.c_float), ("clip_reward", C.c_float), ("epsilon", C.c_float),
        ("norm_obs", C.c_int32), ("norm_reward", C.c_int32), ("training", C.c_int32), ("reserved", C.c_int32),
    ]


class XarmHerConfig(C.Structure):
    _fields_ = [
        ("num_envs", C.c_int64), ("episodes_per_env", C.c_int32), ("max_episode_length", C.c_int32),
        ("obs_dim", C.c_int32), ("goal_dim", C.c_int32), ("action_dim", C.c_int32),
        ("task", C.c_int32), ("reward_type", C.c_int32), ("num_obj", C.c_int32),
        ("n_sampled_goal", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64),
    ]


ABI_VERSION = 2   # XARM_ABI_VERSION of include/xarm_abi.h

# every symbol include/xarm_abi.h declares
ABI_SYMBOLS = [
    "xarm_task_dims", "xarm_create", "xarm_destroy", "xarm_bind", "xarm_reset", "xarm_step", "xarm_step_host",
    "xarm_reset_host", "xarm_compute_reward", "xarm_get_state", "xarm_set_state", "xarm_get_obs", "xarm_graph_capture",
    "xarm_episode_stats", "xarm_launch_count", "xarm_last_error", "xarm_abi_version", "xarm_set_profiling", "xarm_kernel_times",
    "xarm_measure_fp32_peak",
    "xarm_vecnorm_create", "xarm_vecnorm_destroy", "xarm_vecnorm_reset", "xarm_vecnorm_step", "xarm_vecnorm_set_training",
    "xarm_vecnorm_get_stats", "xarm_vecnorm_set_stats", "xarm_vecnorm_normalize_obs", "xarm_vecnorm_normalize_reward",
    "xarm_her_create", "xarm_her_destroy", "xarm_her_begin", "xarm_her_add", "xarm_her_sample", "xarm_her_stats",
]

_lib = None


class XarmError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """Load the native library (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("XARM_B200_LIB") or _build.LIB   # XARM_B200_LIB: development override (A/B builds)
    if path == _build.LIB and _build.is_stale():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(path):
                raise XarmError(f"libxarm_b200.so is missing and could not be built ({e}); there is no CPU fallback") from e
    L = C.CDLL(path)
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    L.xarm_task_dims.argtypes = [C.c_int32, C.c_int32, i32p, i32p, i32p, i32p]
    L.xarm_create.argtypes = [C.POINTER(XarmConfig), C.POINTER(vp)]
    L.xarm_destroy.argtypes = [vp]
    L.xarm_bind.argtypes = [vp, C.POINTER(XarmBuffers)]
    L.xarm_reset.argtypes = [vp, vp, vp]
    L.xarm_step.argtypes = [vp, vp]
    L.xarm_step_host.argtypes = [vp] + [vp] * 9 + [vp]
    L.xarm_reset_host.argtypes = [vp, vp, vp, vp, vp]
    L.xarm_compute_reward.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, vp, C.c_int64, vp, vp]
    L.xarm_get_state.argtypes = [vp, vp]
    L.xarm_set_state.argtypes = [vp, vp]
    L.xarm_get_obs.argtypes = [vp, vp]
    L.xarm_graph_capture.argtypes = [vp, vp]
    L.xarm_episode_stats.argtypes = [vp, C.POINTER(C.c_double), vp]
    L.xarm_set_profiling.argtypes = [vp, C.c_int32]
    L.xarm_kernel_times.argtypes = [vp, C.c_char_p, C.c_int64]
    L.xarm_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double)]
    dp = C.POINTER(C.c_double)
    L.xarm_vecnorm_create.argtypes = [C.POINTER(XarmVecNormConfig), C.POINTER(vp)]
    L.xarm_vecnorm_destroy.argtypes = [vp]
    L.xarm_vecnorm_reset.argtypes = [vp, vp, vp, vp]
    L.xarm_vecnorm_step.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.xarm_vecnorm_set_training.argtypes = [vp, C.c_int32]
    L.xarm_vecnorm_normalize_obs.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int32, vp]
    L.xarm_vecnorm_normalize_reward.argtypes = [vp, vp, C.c_int64, vp, vp]
    L.xarm_vecnorm_get_stats.argtypes = [vp, dp, dp, dp, dp]
    L.xarm_vecnorm_set_stats.argtypes = [vp, dp, dp, C.c_double, dp]
    L.xarm_her_create.argtypes = [C.POINTER(XarmHerConfig), C.POINTER(vp)]
    L.xarm_her_destroy.argtypes = [vp]
    L.xarm_her_begin.argtypes = [vp, vp, vp, vp, vp, vp]
    L.xarm_her_add.argtypes = [vp] + [vp] * 8 + [vp]
    L.xarm_her_sample.argtypes = [vp, C.c_int64] + [vp] * 9 + [vp]
    L.xarm_her_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.xarm_launch_count.restype = C.c_int64
    L.xarm_last_error.restype = C.c_char_p
    if L.xarm_abi_version() != ABI_VERSION:
        raise XarmError(f"libxarm_b200.so has ABI version {L.xarm_abi_version()}, this package needs {ABI_VERSION}: rebuild (python -m gym_xarm_b200.build --force)")
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = load().xarm_last_error().decode("utf-8", "replace")
        if rc == -1:
            if "not implemented" in msg or "reads simulator state" in msg:
                raise NotImplementedError(msg)
            raise ValueError(msg)
        raise XarmError(f"{what} failed ({rc}): {msg}")
