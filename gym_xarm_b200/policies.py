"""Scripted policies of the reference, batched.

`ezpolicy` restates XarmHandover.ezpolicy [REF gym_xarm/envs/xarm_handover.py:404-446], the hand-written two-arm handover
script the reference uses as a test fixture: arm 1 approaches the lego from -x and grasps it,
lifts it towards arm 2, arm 2 approaches from +x, and once both hold it arm 2 pulls back.  It reads the 29-dim
Handover observation (num_obj = 1): lego pos 0:3, quat 3:7, linvel 7:10, angvel 10:13, hand-1 pos 13:16, vel 16:19,
finger-1 joint 19, its velocity 20, hand-2 pos 21:24, vel 24:27, finger-2 joint 27, its velocity 28.

Works on one observation ([29] numpy, as the reference) or on a batch ([N, 29] numpy array or torch tensor on any
device); the arithmetic is the reference's, in the array's own dtype.  The CPU test suite checks it bit for bit against
vectors produced by the reference's own function (tests/golden/make_golden_ezpolicy.py).
"""
import numpy as np


def ezpolicy(obs):
    """obs: dict with 'observation', or the observation array itself.  Returns actions [..., 8]."""
    if isinstance(obs, dict):
        obs = obs["observation"]
    try:
        import torch
        is_torch = isinstance(obs, torch.Tensor)
    except ImportError:  # pragma: no cover
        is_torch = False
    if is_torch:
        import torch
        o = obs
        norm = lambda x: torch.linalg.norm(x, dim=-1)  # noqa: E731
        where, zeros = torch.where, torch.zeros(o.shape[:-1] + (8,), dtype=o.dtype, device=o.device)
        full = lambda v: torch.full(o.shape[:-1], v, dtype=o.dtype, device=o.device)  # noqa: E731
        vec = lambda x, y, z: torch.tensor([x, y, z], dtype=o.dtype, device=o.device)  # noqa: E731
    else:
        o = np.asarray(obs)

        def norm(x):  # the reference takes np.linalg.norm of one 3-vector (sqrt(dot(x, x)) through BLAS); row by row keeps its bits
            if x.ndim == 1:
                return np.linalg.norm(x)
            return np.array([np.linalg.norm(v) for v in x.reshape(-1, x.shape[-1])], dtype=x.dtype).reshape(x.shape[:-1])

        where, zeros = np.where, np.zeros(o.shape[:-1] + (8,), dtype=o.dtype if o.dtype.kind == "f" else np.float64)
        full = lambda v: np.full(o.shape[:-1], v, dtype=zeros.dtype)  # noqa: E731
        vec = lambda x, y, z: np.array([x, y, z], dtype=zeros.dtype)  # noqa: E731
    obj, g1, g2 = o[..., 0:3], o[..., 13:16], o[..., 21:24]
    f1, f2 = o[..., 19], o[..., 27]
    d1, d2 = norm(obj - g1), norm(obj - g2)
    grasp1 = (f1 < 0.25) & (d1 < 0.05)          # [REF :421-422] (finger joints never exceed 0.04: the first term always holds)
    grasp2 = (f2 < 0.25) & (d2 < 0.05)
    delta1 = obj - g1 + vec(-0.07, 0.0, 0.0)    # approach points 7 cm to either side of the lego [REF :423-426]
    delta2 = obj - g2 + vec(0.07, 0.0, 0.0)
    dir1 = delta1 / norm(delta1)[..., None]
    dir2 = delta2 / norm(delta2)[..., None]
    act = zeros
    act[..., 3] = where(d1 < 0.1, full(-0.5), full(0.5))     # close a gripper within 10 cm of the lego, open it otherwise
    act[..., 7] = where(d2 < 0.1, full(-0.5), full(0.5))
    lift = vec(0.5, 0.0, 0.5)
    m1 = (~grasp1)[..., None]                                   # arm 1 still approaching
    m2 = (grasp1 & ~grasp2)[..., None]                          # arm 1 holds: lift towards arm 2, arm 2 approaches
    m3 = grasp1 & grasp2                                        # both hold: arm 2 pulls back
    act[..., 0:3] = where(m1, dir1, where(m2, lift + 0 * dir1, 0 * dir1))
    act[..., 4:7] = where(m2, dir2, 0 * dir2)
    act[..., 4] = where(m3, full(-0.5), act[..., 4])
    return act
