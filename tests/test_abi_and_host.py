"""CPU suite for the boundary and the host logic: the C-ABI library loads and exports every symbol include/xarm_abi.h
declares (no compute calls without a GPU), struct layouts, the config/registry mirror of the reference interface,
slab partitioning and the episode-statistics exchange over a world_size-2 gloo group."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "xarm_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xarm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gym_xarm_b200 import _native
    L = _native.load()
    names = _header_functions()
    assert sorted(names) == sorted(_native.ABI_SYMBOLS)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/xarm_abi.h but not exported"
    assert L.xarm_abi_version() == _native.ABI_VERSION == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _native.lib_path()], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", out), n


def test_struct_layouts_match_header():
    from gym_xarm_b200 import _native
    assert C.sizeof(_native.XarmConfig) == 72 and _native.XarmConfig.num_envs.offset == 48 and _native.XarmConfig.seed.offset == 64
    assert C.sizeof(_native.XarmBuffers) == 9 * 8
    assert C.sizeof(_native.XarmHerConfig) == 56 and _native.XarmHerConfig.seed.offset == 48 and _native.XarmHerConfig.device.offset == 44
    from oracle import oracle as orc
    assert C.sizeof(orc.XarmConfig) == C.sizeof(_native.XarmConfig)


def test_her_create_rejects_bad_configs_without_gpu():
    """argument errors of xarm_her_create surface as ValueError / NotImplementedError before any CUDA call"""
    import pytest
    from gym_xarm_b200 import _native
    L = _native.load()
    base = dict(num_envs=8, episodes_per_env=4, max_episode_length=50, obs_dim=24, goal_dim=3, action_dim=4, task=1, reward_type=0,
                num_obj=1, n_sampled_goal=4, device=0, seed=0)
    for bad, exc in [(dict(episodes_per_env=1), ValueError), (dict(goal_dim=10), ValueError), (dict(num_envs=0), ValueError),
                     (dict(task=9), ValueError), (dict(reward_type=1), NotImplementedError), (dict(n_sampled_goal=-1), ValueError)]:
        cfg = _native.XarmHerConfig(**{**base, **bad})
        h = C.c_void_p()
        rc = L.xarm_her_create(C.byref(cfg), C.byref(h))
        assert rc == -1
        with pytest.raises(exc):
            _native.check(rc, "xarm_her_create")
    from gym_xarm_b200.her import XarmHerReplayBuffer
    with pytest.raises(NotImplementedError):
        XarmHerReplayBuffer(num_envs=8, obs_dim=24, goal_dim=3, action_dim=4, task=1, max_episode_length=50, goal_selection_strategy="final")


def test_use_stand_is_refused_not_ignored():
    """config['use_stand']=True [REF xarm_handover.py:391-392] puts a static box under every goal; the collider is not built, so
    xarm_create must refuse the config (before any CUDA call) instead of simulating a different world silently."""
    import pytest
    from gym_xarm_b200 import _native
    L = _native.load()
    cfg = _native.XarmConfig(task=4, reward_type=1, num_obj=1, goal_shape=1, init_grasp_rate=0.0, goal_ground_rate=0.0, same_side_rate=0.5,
                             use_stand=1, max_episode_steps=0, auto_reset=1, device=0, stagger_phases=0, num_envs=4, env_index_base=0, seed=0)
    h = C.c_void_p()
    rc = L.xarm_create(C.byref(cfg), C.byref(h))
    assert rc == -1 and b"use_stand" in L.xarm_last_error()
    with pytest.raises(NotImplementedError):
        _native.check(rc, "xarm_create")


def test_task_dims_without_gpu():
    from gym_xarm_b200 import _native, SPECS
    L = _native.load()
    for name, spec in SPECS.items():
        a, o, g, s = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        assert L.xarm_task_dims(spec.task, 1, a, o, g, s) == 0
        nobj = {"stack_tower": 3}.get(name, 1)
        assert (a.value, o.value, g.value) == spec.dims(nobj)
        from oracle import oracle as orc
        assert s.value == orc.dims(name, {"reach": 0, "stack_tower": 3}.get(name, 1))[3]
    assert L.xarm_task_dims(99, 1, None, None, None, None) == -1 and b"unsupported" in L.xarm_last_error()


def test_no_cpu_fallback():
    import torch
    from gym_xarm_b200 import XarmVecEnv, _native, make
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.XarmError):
        XarmVecEnv("reach", 4)
    with pytest.raises(_native.XarmError):
        make("XarmReach-v0")
    # the product package never imports the oracle or the host-sim harness
    for root, _, files in os.walk(os.path.join(ROOT, "gym_xarm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(root, f)).read()
                assert "oracle" not in txt.replace("oracle stand-in", "") or f.endswith((".cuh", ".cu")) and "import" not in txt, f
                assert "hostsim" not in txt or "tests/hostsim" in txt, f


def test_registry_and_config_mirror_reference():
    from gym_xarm_b200 import REGISTRY, SPECS
    from gym_xarm_b200.specs import normalize_config
    # ids the snapshot registers [REF gym_xarm/__init__.py:6-22] and the north-star spellings
    for env_id in ("XarmReach-v0", "XarmHandover-v0", "XarmPickAndPlace-v1", "XarmPDPickAndPlace-v0", "XarmPDStackTower-v0",
                   "XarmPDPushWithDoor-v0", "XarmPDHandover-v1"):
        assert env_id in REGISTRY
    assert [SPECS[t].max_episode_steps for t in ("reach", "handover", "pick_and_place")] == [25, 100, 50]
    cfg = normalize_config(SPECS["pick_and_place"], {"GUI": False, "init_grasp_rate": 0.0, "goal_ground_rate": 0.0, "num_obj": 1,
                                                     "reward_type": "sparse", "goal_shape": "ground"})
    assert cfg["goal_shape"] == "ground" and cfg["num_obj"] == 1
    for bad in ("incremental", "dense_diff_o2g", "nonsense"):   # D6: broken in the reference -> NotImplementedError like unknown types
        with pytest.raises(NotImplementedError):
            normalize_config(SPECS["pick_and_place"], {"reward_type": bad})
    assert normalize_config(SPECS["handover"], {"goal_shape": "any"})["goal_shape"] == "air"   # [REF test.py:9-15]
    assert normalize_config(SPECS["stack_tower"], {})["num_obj"] == 3


def test_spaces_and_infolist():
    from gym_xarm_b200.spaces import Box, Dict
    from gym_xarm_b200.vec_env import InfoList
    b = Box(-1.0, 1.0, shape=(4,), dtype=np.float32)
    assert b.contains(b.sample()) and not b.contains(np.full(4, 2, np.float32))
    d = Dict(dict(observation=Box(-np.inf, np.inf, shape=(8,), dtype=np.float32)))
    assert d.contains({"observation": np.zeros(8, np.float32)})

    class E:
        num_envs, obs_dim, goal_dim = 3, 2, 1
    term = np.arange(12, dtype=np.float32).reshape(3, 4)
    infos = InfoList(E(), np.array([0, 1, 0], np.float32), np.array([0, 0, 1], np.uint8), np.array([0, 1, 1], np.uint8), term)
    assert len(infos) == 3 and infos[0] == {"is_success": 0.0, "TimeLimit.truncated": False}
    assert infos[1]["is_success"] == 1.0 and list(infos[1]["terminal_observation"]["observation"]) == [4.0, 5.0]
    assert infos[2]["TimeLimit.truncated"] and list(infos[2]["terminal_observation"]["desired_goal"]) == [11.0]


def test_slab_partition():
    from gym_xarm_b200.distributed import slab
    for total, world in ((1048576, 8), (10, 3), (7, 8)):
        parts = [slab(total, r, world) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == total
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    assert slab(1048576, 3, 8) == (393216, 131072)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import torch, torch.distributed as dist
from gym_xarm_b200 import distributed as xd
rank, world, local = xd.init_from_env(backend="gloo")
first, count = xd.slab(10, rank, world)
stats = {"episodes": count, "return_sum": float(rank + 1), "length_sum": 50.0 * count, "success_sum": float(rank), "diverged": 0}
tot = xd.gather_episode_stats(stats)
mx = xd.max_over_ranks(10.0 * (rank + 1))
xd.barrier()
if rank == 0:
    print("RESULT", tot["episodes"], tot["return_sum"], tot["mean_length"], tot["success_rate"], mx, len(tot["per_rank"]))
dist.destroy_process_group()
"""


def test_episode_stats_all_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    assert [float(x) for x in line[1:]] == [10.0, 3.0, 50.0, 0.1, 20.0, 2.0]


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` prints one JSON line with the contract's keys (tiny run)."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_committed_traffic_capture_matches_its_launch_list():
    """profiles/traffic_pick_and_place.json (what bench.py quotes as roofline.traffic) must be what tools/ncu_traffic.py computes
    from the committed ncu launch list of the same capture, and it must carry the hash of the kernel sources in this tree."""
    import glob
    import importlib.util
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(root, "profiles", "traffic_pick_and_place.json")))
    lists = sorted(glob.glob(os.path.join(root, "profiles", "r2z_launches_one_step.csv.gz")))
    assert lists, "the launch list of the capture is not committed"
    spec = importlib.util.spec_from_file_location("ncu_traffic", os.path.join(root, "tools", "ncu_traffic.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = m.parse(lists[-1], "pick_and_place", 131072, write=False)
    assert abs(out["dram_bytes_per_step"] - d["dram_bytes_per_step"]) <= 1e-6 * d["dram_bytes_per_step"]
    assert {k["kernel"] for k in out["kernels"]} == {k["kernel"] for k in d["kernels"]}
    import bench
    import pytest
    if d["kernel_source_hash"] != bench.kernel_source_hash():   # bench.py then prints traffic = null with a note; not a test failure
        pytest.skip("the committed capture is from other kernel sources than this tree: re-run tools/ncu_traffic.py on a GPU box")
