"""CPU check of the CUDA path's per-env logic: the kernel bodies (gym_xarm_b200/csrc/*.cuh) compiled for the host by
tests/hostsim (float32, same code as the device minus FMA contraction) against the float64 oracle.  This is a
development/test harness, not a product path - on the GPU box the same comparisons run against the real kernels
(tests/test_gpu_parity.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.hostsim import hostsim as hs
from tests.parity_util import Worst, compare_full, run_dense_staged

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_logic.npz"))
TASKS = ["reach", "pick_and_place", "stack_tower", "push_with_door", "handover"]
NOBJ = {"reach": 0, "pick_and_place": 1, "stack_tower": 3, "push_with_door": 1, "handover": 1}
TOL = 1e-3  # north_star: 1e-3 rad / 1e-3 m over 50 contact-free steps


def _pair(task, n, seed):
    gs = "air" if task == "pick_and_place" else "ground"
    v = hs.HostSimVec(orc.make_config(task, num_envs=n, seed=seed, auto_reset=0, goal_shape=gs))
    ref = [orc.OracleEnv(task, env_index=i, seed=seed, auto_reset=0, goal_shape=gs) for i in range(n)]
    return v, ref


@pytest.mark.parametrize("task", TASKS)
def test_constructor_state_and_reset_goals_bit_exact(task):
    v, ref = _pair(task, 6, 21)
    assert np.array_equal(v.get_state(), np.stack([r.get_state() for r in ref]))
    o = v.reset()
    ro = [r.reset() for r in ref]
    assert np.array_equal(o["desired_goal"], np.stack([x["desired_goal"] for x in ro]))   # goal sampling: bit-exact
    st, rst = v.get_state(), np.stack([r.get_state() for r in ref])
    G = v.G
    assert np.array_equal(st[:, -5 - G:-5], rst[:, -5 - G:-5]) and np.array_equal(st[:, -5:-3], rst[:, -5:-3])  # goal, step, episode


@pytest.mark.parametrize("task", TASKS)
def test_step_parity_contact_free(task):
    """50 steps (25 for Reach) from identical states and action tapes; envs whose gripper touches something drop out of
    the strict comparison from then on (stiff contacts amplify float32 rounding; covered statistically on the GPU)."""
    n = 6
    v, ref = _pair(task, n, 5)
    v.reset()
    for r in ref:
        r.reset()
    st0 = np.stack([r.get_state() for r in ref])
    if task == "pick_and_place":
        # park the lego of half the envs on the table outside the gripper's workspace (x<=0.5, |y|<=0.3): their whole
        # tape is gripper-contact-free while the lego-table contact rows stay active
        st0[::2, 27:30] = [0.65, 0.42, 0.04]
        st0[::2, 30:40] = [0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    v.set_state(st0)
    for r, s in zip(ref, st0):
        r.set_state(s)
        r.arm_contacts()
    clean = np.ones(n, bool)
    worst = Worst()
    rng = np.random.default_rng(3)
    ndof = 13 if task == "reach" else 9
    narm = 1 if task in ("reach", "pick_and_place") else 2
    nq = 3 * ndof * narm
    for t in range(25 if task == "reach" else 50):
        a = rng.uniform(-1, 1, (n, v.A)).astype(np.float32)
        if task == "reach":  # keep the xArm gripper inside its joint range (tests/test_gpu_parity.py::_actions explains)
            a[:, 3] = 0.02 * np.abs(a[:, 3])
        if task == "handover":  # keep the fingertips above the table and the lego (eef z can go down to 0.1 there)
            a[:, 2] = 0.5 + 0.5 * np.abs(a[:, 2])
            a[:, 6] = 0.5 + 0.5 * np.abs(a[:, 6])
        o, r, d, s, tr = v.step(a)
        res = [e.step(a[i]) for i, e in enumerate(ref)]
        clean &= np.array([e.arm_contacts() == 0 for e in ref])
        st, rst = v.get_state(), np.stack([e.get_state() for e in ref])
        assert np.isfinite(st).all()
        c = clean
        robs_t = {k: np.stack([x[0][k] for x in res]) for k in o}
        compare_full(task, NOBJ[task], st, rst, o, robs_t, c, worst, msg=f"step {t}")   # whole state record + whole observation dict
        assert np.array_equal(d[c], np.array([x[2] for x in res])[c])
        assert np.array_equal(s[c], np.array([x[3]["is_success"] for x in res], np.float32)[c])
        want = orc.compute_reward(task, "sparse", max(NOBJ[task], 1), o["achieved_goal"], o["desired_goal"])
        assert np.array_equal(r.view(np.uint32), want.view(np.uint32))
    print(f"{task}: worst |host build - oracle|: {worst}")
    assert clean.sum() >= 2


@pytest.mark.parametrize("task,prefix", [("reach", "reach_sparse"), ("pick_and_place", "pap"), ("stack_tower", "stack"),
                                         ("push_with_door", "push"), ("handover", "handover")])
def test_kernel_reward_matches_reference_golden(task, prefix):
    ag, dg = GOLD[f"{prefix}_ag"], GOLD[f"{prefix}_dg"]
    want = GOLD[f"{prefix}_reward"] if f"{prefix}_reward" in GOLD else GOLD[f"{prefix}_sparse_reward"]
    got = hs.compute_reward(orc.TASKS[task], 0, max(NOBJ[task], 1), ag, dg)
    if want.dtype == np.float32:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    else:
        assert np.array_equal(got.astype(np.float64), want)


def _rand_rot(rng, small=False):
    q = rng.normal(size=4)
    if small:
        q = np.array([0, 0, 0, 1.0]) + 0.05 * q
    q /= np.linalg.norm(q)
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_box_box_matches_oracle():
    """Same contact set (count, normal, points, depth) from the float32 kernel routine and the float64 oracle routine."""
    rng = np.random.default_rng(0)
    n_hit = n_tie = 0
    for trial in range(400):
        hA, hB = rng.uniform(0.02, 0.08, 3), rng.uniform(0.02, 0.08, 3)
        RA, RB = _rand_rot(rng, trial % 2 == 0), _rand_rot(rng, trial % 3 == 0)
        cA = rng.uniform(-0.05, 0.05, 3)
        cB = cA + rng.uniform(-0.09, 0.09, 3)
        co = orc.box_box((cA, RA, hA), (cB, RB, hB))
        ck = hs.box_box((cA, RA, hA), (cB, RB, hB))
        if len(co) != len(ck):
            # float32 may flip a near-tie; accept only if the penetration is marginal
            assert max([c[3] for c in co] + [c[3] for c in ck] + [0]) < 1e-4, trial
            continue
        for (pa, pb, nn, d), (pa2, pb2, n2, d2) in zip(co, ck):
            if np.abs(nn - n2).max() > 1e-3:   # near-tie between two SAT axes
                continue
            if np.abs(pa2 - pa).max() > 2e-5:  # manifold reduction picked another (equally extreme) candidate: a float32 tie
                n_tie += 1
                continue
            n_hit += 1
            np.testing.assert_allclose(pb2, pb, atol=2e-5)
            assert abs(d - d2) < 2e-5 and d >= -1e-9
            assert abs(np.linalg.norm(nn) - 1) < 1e-9
    assert n_hit > 200 and n_tie <= 4


def test_box_on_table_gives_four_corner_contacts():
    I = np.eye(3)
    pts = orc.box_box(((0.4, 0.1, 0.035), I, (0.025, 0.025, 0.04)), ((0, 0, -0.025), I, (0.75, 0.5, 0.025)))
    assert len(pts) == 4
    for pa, pb, nn, d in pts:
        np.testing.assert_allclose(nn, [0, 0, 1], atol=1e-12)
        assert abs(d - 0.005) < 1e-12 and abs(abs(pa[0] - 0.4) - 0.025) < 1e-12
    assert orc.box_box(((0.4, 0.1, 0.0401), I, (0.025, 0.025, 0.04)), ((0, 0, -0.025), I, (0.75, 0.5, 0.025))) == []


def test_auto_reset_and_stats_semantics():
    """VecEnv semantics of the step kernel: a finished env is reset inside step(), done/truncated flags, new goal."""
    cfg = orc.make_config("reach", num_envs=3, seed=4, auto_reset=1)
    v = hs.HostSimVec(cfg)
    v.reset()
    g0 = v.get_obs()["desired_goal"].copy()
    for t in range(25):
        o, r, d, s, tr = v.step(np.zeros((3, 4), np.float32))
        assert d.all() == (t == 24) and tr.all() == (t == 24)
    assert not np.array_equal(o["desired_goal"], g0)          # already the next episode's goal
    assert np.array_equal(v.get_state()[:, -5], np.zeros(3))  # step counter back to 0
    assert np.array_equal(v.get_state()[:, -4], np.full(3, 2.0))  # second episode


@pytest.mark.parametrize("task", ["pick_and_place", "handover"])
def test_dense_staged_reward_host_build(task):
    """The staged dense rewards of the kernel code (host build) in a scripted rollout - every stage - against the oracle's
    staged-reward function, itself pinned on the reference's compute_reward (same body as the -m gpu test)."""
    class Host:
        def __init__(self, n, cfg):
            pp = task == "pick_and_place"
            self.v = hs.HostSimVec(orc.make_config(task, num_envs=n, seed=31, auto_reset=0, reward_type="dense", goal_shape="air" if pp else "ground",
                                                   init_grasp_rate=cfg.get("init_grasp_rate", 0.0)))
        def reset(self): self.v.reset()
        def set_state(self, st): self.v.set_state(st)
        def get_state(self): return self.v.get_state()
        def get_obs(self): return self.v.get_obs()
        def step(self, a):
            o, r, d, s, tr = self.v.step(a)
            return o, r
        def close(self): pass
    run_dense_staged(task, Host, 48)
