"""Shared comparison of a whole env state / observation dict with the oracle's (tests/test_gpu_parity.py on the CUDA path,
tests/test_kernel_logic_cpu.py on the host build of the kernel bodies).

Tolerances.  Positions (joint angles, object pose, door travel, the position words of the observation): 1e-3 rad / 1e-3 m
(north_star).  Velocities (joint rates, object linear / angular velocity, door rate, hand COM velocity, the relative
velocity words): north_star gives none; a substep velocity is (position change) / h with h = 1/900 .. 1/240 s, so a
position agreement of 1e-5 over a substep is a velocity agreement of ~1e-2.  VTOL = 2e-2 rad/s | m/s is asserted; the
measured maxima are printed by the tests (host build: 2e-4; CUDA path: see DESIGN.md 7)."""
import numpy as np

TOL = 1e-3
VTOL = 2e-2


def obs_velocity_mask(task, nobj, obs_dim):
    """True for the words of obs['observation'] that are velocities [REF xarm_reach.py:144-161; xarm_pick_and_place.py:220-248;
    xarm_stack_tower.py:164-199; xarm_push_with_door.py:162-193; xarm_handover.py:299-336]."""
    m = np.zeros(obs_dim, bool)
    if task in ("reach", "pick_and_place"):
        m[[3, 4, 5, 7]] = True                      # hand COM velocity, finger rate
        for i in range(nobj if task == "pick_and_place" else 0):
            b = 8 + 16 * i
            m[b + 7:b + 13] = True                  # lego velocity relative to the hand, angular velocity
    else:
        m[7 * nobj:13 * nobj] = True                # linear and angular velocities of the cubes
        w = 6 if task == "push_with_door" else 8
        for a in range(2):
            b = 13 * nobj + w * a
            m[b + 3:b + 6] = True
            if w == 8:
                m[b + 7] = True
    return m


class Worst:
    def __init__(self):
        self.d = {}

    def add(self, key, diff):
        if diff.size:
            self.d[key] = max(self.d.get(key, 0.0), float(np.abs(diff).max()))

    def __str__(self):
        return ", ".join(f"{k} {v:.1e}" for k, v in self.d.items())


def compare_full(task, nobj, st, rst, obs, robs, clean, worst, msg="", arm_joints=None):
    """st / rst: [n, S] state records (CUDA or host build / oracle); obs / robs: dicts of [n, dim] arrays; clean: bool [n] -
    the envs under the strict tolerance (gripper touched nothing so far)."""
    ndof = 13 if task == "reach" else 9
    narm = 1 if task in ("reach", "pick_and_place") else 2
    nq = 3 * ndof * narm
    na = arm_joints or (7 if task == "reach" else 9)   # Reach: the 7 arm joints (its gripper: DESIGN.md 7)
    c = clean
    D = (np.asarray(st, np.float64) - np.asarray(rst, np.float64))[c]

    def chk(key, sl, tol):
        worst.add(key, D[:, sl])
        np.testing.assert_allclose(D[:, sl], 0, atol=tol, err_msg=f"{task} {msg} {key}")

    for arm in range(narm):
        b = arm * 3 * ndof
        chk("q", slice(b, b + na), TOL)
        chk("qd", slice(b + ndof, b + ndof + na), VTOL)
        chk("q_target", slice(b + 2 * ndof, b + 2 * ndof + na), TOL)
    for o in range(nobj):
        b = nq + 13 * o
        chk("obj_pose", slice(b, b + 7), TOL)
        chk("obj_vel", slice(b + 7, b + 13), VTOL)
    if task == "push_with_door":
        b = nq + 13 * nobj
        chk("door_q", slice(b, b + 1), TOL)
        chk("door_qd", slice(b + 1, b + 2), VTOL)
    O = np.asarray(obs["observation"], np.float64)[c] - np.asarray(robs["observation"], np.float64)[c]
    vm = obs_velocity_mask(task, nobj, O.shape[1])
    pm = ~vm
    worst.add("obs_pos", O[:, pm]); worst.add("obs_vel", O[:, vm])
    np.testing.assert_allclose(O[:, pm], 0, atol=TOL, err_msg=f"{task} {msg} observation (position words)")
    np.testing.assert_allclose(O[:, vm], 0, atol=VTOL, err_msg=f"{task} {msg} observation (velocity words)")
    A = np.asarray(obs["achieved_goal"], np.float64)[c] - np.asarray(robs["achieved_goal"], np.float64)[c]
    worst.add("achieved_goal", A)
    np.testing.assert_allclose(A, 0, atol=TOL, err_msg=f"{task} {msg} achieved_goal")
    assert np.array_equal(np.asarray(obs["desired_goal"]), np.asarray(robs["desired_goal"])), f"{task} {msg} desired_goal"


def run_dense_staged(task, make_env, n, fn_tol=1e-6):
    """body of test_dense_staged_reward_in_step (CUDA path) / test_dense_staged_reward_host_build: make_env(n, cfg) returns an
    adapter with reset / set_state / get_state / get_obs / step(actions) -> (obs dict of numpy arrays, reward) / close"""
    from gym_xarm_b200.policies import ezpolicy
    from oracle import oracle as orc
    pp = task == "pick_and_place"
    cfg = {"reward_type": "dense"}
    if pp:
        cfg.update(init_grasp_rate=0.5, goal_shape="air")
    env = make_env(n, cfg)
    ref = orc.OracleBatch(task, n, seed=31, auto_reset=0, reward_type="dense", goal_shape="air" if pp else "ground", init_grasp_rate=0.5 if pp else 0.0)
    env.reset()
    ref.reset()
    ref.arm_contacts()
    st0 = ref.get_state()
    env.set_state(st0)
    ref.set_state(st0)
    obs = env.get_obs()
    clean = np.ones(n, bool)
    rng = np.random.default_rng(6)
    grip = np.where(rng.random(n) < 0.7, -1.0, rng.uniform(-1, 1, n)).astype(np.float32)
    stages, worst_fn, worst_env = set(), 0.0, 0.0
    for t in range(30):
        if pp:
            a = np.zeros((n, 4), np.float32)
            if t >= 3:
                a[:, 0], a[:, 2] = -0.3, 1.0
            a[:, 3] = grip
            a[n // 2:] = rng.uniform(-1, 1, (n - n // 2, 4))
        else:
            a = np.clip(np.asarray(ezpolicy(obs["observation"]), np.float64), -1, 1).astype(np.float32)
            a[n // 2:] = rng.uniform(-1, 1, (n - n // 2, 8))
        obs, r = env.step(a)
        robs, rrew, _, _, _ = ref.step(a)
        clean &= ref.arm_contacts() == 0
        st = env.get_state()
        o, ag, dg = obs["observation"], obs["achieved_goal"], obs["desired_goal"]
        g = st[:, -2:].astype(np.int64)
        for i in range(n):
            if pp:
                want = orc.dense_reward(task, 1, o[i, 0:3], ag[i], dg[i], [g[i, 0] & 1])
                stages.add(0 if not g[i, 0] & 1 else (2 if ag[i, 2] > np.float32(0.05) else 1))
            else:
                hand = np.stack([o[i, 13:16], o[i, 21:24]]).astype(np.float64)
                hand[:, 2] += 0.088 - 0.021
                flags = [(g[i, 0] >> 1) & 1, (g[i, 1] >> 1) & 1]
                want = orc.dense_reward(task, 1, hand, ag[i], dg[i], flags)
                stages.add(tuple(flags))
            worst_fn = max(worst_fn, abs(float(r[i]) - want))
        worst_env = max(worst_env, float(np.abs(r - rrew)[clean].max()) if clean.any() else 0.0)
    print(f"{task} dense: stages seen {sorted(stages)}; worst |reward - oracle function| {worst_fn:.2e}; vs the oracle env on "
          f"{int(clean.sum())} contact-free envs {worst_env:.2e}")
    assert worst_fn < fn_tol and worst_env < 2.5e-4
    assert len(stages) >= (3 if pp else 2), stages
    env.close()
