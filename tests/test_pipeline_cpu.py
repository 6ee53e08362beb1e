"""CPU check of the step PIPELINE: the per-kernel bodies of gym_xarm_b200/csrc/xarm_pipeline.cuh (action -> NSUB x {setup ->
light || heavy} -> finish -> auto-reset passes), run on the host by tests/hostsim one loop per kernel launch, must give
bit-for-bit the results of the fused per-env form (body_step / body_reset), which test_kernel_logic_cpu.py checks against
the oracle.  Covers the scratch-slab round trip of the rows, the light/heavy classification, the island split of the
light solver, list-mode launches and the reset stages."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.hostsim import hostsim as hs

CASES = [("pick_and_place", 24, 60, 0), ("reach", 8, 30, 0), ("stack_tower", 4, 12, 8), ("push_with_door", 4, 12, 8), ("handover", 4, 12, 8)]


@pytest.mark.parametrize("task,n,steps,limit", CASES)
def test_pipeline_equals_fused_step(task, n, steps, limit):
    gs = "air" if task == "pick_and_place" else "ground"
    cfg = orc.make_config(task, num_envs=n, seed=7, auto_reset=1, goal_shape=gs, max_episode_steps=limit)
    a, b = hs.HostSimVec(cfg, pipeline=False), hs.HostSimVec(cfg, pipeline=True)
    oa, ob = a.reset(), b.reset()
    assert all(np.array_equal(oa[k], ob[k]) for k in oa)
    assert np.array_equal(a.get_state(), b.get_state())
    rng = np.random.default_rng(1)
    n_done = 0
    for t in range(steps):
        act = rng.uniform(-1, 1, (n, a.A)).astype(np.float32)
        ra, rb = a.step(act), b.step(act)
        for k in ra[0]:
            assert np.array_equal(ra[0][k], rb[0][k]), (t, k)
        for x, y in zip(ra[1:], rb[1:]):
            assert np.array_equal(x, y), t
        assert np.array_equal(a.get_state(), b.get_state()), t
        n_done += int(ra[2].sum())
    assert n_done >= n  # every env went through at least one auto-reset
    m = np.zeros(n, np.uint8)
    m[::3] = 1
    oa, ob = a.reset(m), b.reset(m)  # masked reset = list-mode passes
    assert all(np.array_equal(oa[k], ob[k]) for k in oa)
    assert np.array_equal(a.get_state(), b.get_state())


def test_light_solver_early_exit_matches_joint_loop():
    """An arm at rest over a resting lego: every row is below the residual threshold after a few sweeps, so Bullet's joint
    loop stops early.  The island-split light solver has to stop both islands at that same sweep (pipeline == fused is
    not enough here: both use it), so compare with the oracle's joint loop."""
    task = "pick_and_place"
    cfg = orc.make_config(task, num_envs=2, seed=3, auto_reset=0, goal_shape="air")
    v = hs.HostSimVec(cfg, pipeline=True)
    ref = [orc.OracleEnv(task, env_index=i, seed=3, auto_reset=0, goal_shape="air") for i in range(2)]
    v.reset()
    for r in ref:
        r.reset()
    st0 = np.stack([r.get_state() for r in ref])
    st0[:, 27:30] = [0.65, 0.42, 0.04]           # lego parked on the table, away from the gripper
    st0[:, 30:40] = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    v.set_state(st0)
    for r, s in zip(ref, st0):
        r.set_state(s)
    z = np.zeros((2, 4), np.float32)
    for t in range(40):                           # zero actions: the arm settles on its IK target
        v.step(z)
        for i, r in enumerate(ref):
            r.step(z[i])
    st, rst = v.get_state(), np.stack([r.get_state() for r in ref])
    np.testing.assert_allclose(st[:, :18], rst[:, :18], atol=2e-5)
    np.testing.assert_allclose(st[:, 27:40], rst[:, 27:40], atol=2e-5)
    assert min(r.solver_sweeps_last() for r in ref) < 50 if hasattr(ref[0], "solver_sweeps_last") else True


@pytest.mark.parametrize("task", ["stack_tower", "push_with_door", "handover"])
def test_island_split_of_the_generic_solve_matches_the_joint_loop(task):
    """Heavy envs of the multi-island tasks: the generic substep takes the simple islands (an arm that touches nothing, an object
    on one static box, the untouched door) out of the joint loop (solve_generic_islands).  Islands are independent, so the result
    must be the joint loop's up to float32 rounding of the manifold form.  Scripted push towards the nearest cube; before every
    step the island-split run is re-synchronised to the joint-loop run, so the comparison is per step from identical states
    (gripper contacts amplify differences over several steps)."""
    import os
    n = 24
    cfg = orc.make_config(task, num_envs=n, seed=5, auto_reset=0, goal_shape="ground")
    a = hs.HostSimVec(cfg, pipeline=True)      # island split (default)
    b = hs.HostSimVec(cfg, pipeline=True)      # joint loop (XARM_NO_ISLANDS read per call)
    oa = a.reset()
    b.reset()
    nobj = {"stack_tower": 3, "push_with_door": 1, "handover": 1}[task]
    hw = {"stack_tower": 8, "push_with_door": 6, "handover": 8}[task]
    worst, heavy_steps = 0.0, 0
    obs = oa["observation"]
    for t in range(30):
        act = np.zeros((n, a.A), np.float32)
        cubes = obs[:, :3 * nobj].reshape(n, nobj, 3)
        for arm in range(2):
            hp = 13 * nobj + hw * arm
            hand = obs[:, hp:hp + 3]
            d = np.linalg.norm(cubes[:, :, :2] - hand[:, None, :2], axis=2)
            tgt = cubes[np.arange(n), d.argmin(1)]
            delta = tgt - hand
            delta[:, 2] = -1.0
            k = a.A // 2
            act[:, k * arm:k * arm + 3] = np.clip(8.0 * delta, -1, 1)
        if t in (6, 14, 22) and task != "handover":   # park cube 0 under one finger of arm 0's lowered gripper: a gripper contact for sure
            st = a.get_state()
            hp = 13 * nobj
            st[:, 54:57] = np.stack([obs[:, hp], obs[:, hp + 1] + 0.045, np.full(n, 0.025, np.float32)], axis=1)
            st[:, 57:67] = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
            a.set_state(st)
        b.set_state(a.get_state())
        st0 = a.get_state()
        ra = a.step(act)
        os.environ["XARM_NO_ISLANDS"] = "1"
        try:
            b.step(act)
        finally:
            del os.environ["XARM_NO_ISLANDS"]
        sa, sb = a.get_state(), b.get_state()
        assert np.isfinite(sa).all() and np.isfinite(sb).all()
        moved = np.abs(sa - st0).max()
        diff = np.abs(sa - sb)
        # positions within 2e-5, velocities within 2e-3 of the joint loop after one env step (15 substeps) from identical states
        worst = max(worst, float(diff.max()))
        assert diff.max() < 5e-3, (t, float(diff.max()), np.unravel_index(diff.argmax(), diff.shape))
        heavy_steps += int((a.get_forms() == 1).sum())   # envs that ended the step in the generic (heavy) form
        obs = ra[0]["observation"]
        assert moved > 0
    print(f"{task}: island split vs joint loop, worst |state difference| after one step {worst:.2e}; env-steps that ended in the heavy form: {heavy_steps}")
    assert heavy_steps > 0, "the scenario never reached a heavy env: nothing was compared"


@pytest.mark.parametrize("task", ["stack_tower", "push_with_door", "handover"])
def test_lean_multi_island_setup_classifies_like_the_generic_setup(task):
    """The lean setup of the multi-island tasks calls an env light when NO pair outside "object x one static box" has a point; the
    generic setup (full collision record) is the reference for that decision.  Fused per-env form with the lean path (default)
    against the same form with XARM_NO_LEAN=1 (generic setup + island-split solve for every env), per step from re-synchronised
    states, full-range random actions (Handover: fingers reach the tables): a pair the lean setup overlooked would show as a
    missing contact, i.e. a difference far above rounding.  Same compilation on both sides (no FMA contraction on the host), so
    what remains is the order of operations of the manifold rows."""
    import os
    n = 96
    cfg = orc.make_config(task, num_envs=n, seed=9, auto_reset=0, goal_shape="ground")
    a, b = hs.HostSimVec(cfg, pipeline=False), hs.HostSimVec(cfg, pipeline=False)
    a.reset()
    b.reset()
    rng = np.random.default_rng(4)
    worst = 0.0
    for t in range(25):
        act = rng.uniform(-1, 1, (n, a.A)).astype(np.float32)
        if t % 3 == 0:
            act[:, 2] = -1.0          # arm 0 all the way down every third step: gripper - table / object contacts for sure
            act[:, a.A // 2 + 2] = -1.0
        b.set_state(a.get_state())
        a.step(act)
        os.environ["XARM_NO_LEAN"] = "1"
        try:
            b.step(act)
        finally:
            del os.environ["XARM_NO_LEAN"]
        d = np.abs(a.get_state().astype(np.float64) - b.get_state())
        worst = max(worst, float(d.max()))
        assert d.max() < 1e-4, (t, float(d.max()), np.unravel_index(d.argmax(), d.shape))
    print(f"{task}: lean multi-island path vs generic setup for every env, worst |state difference| after one step {worst:.2e}")
