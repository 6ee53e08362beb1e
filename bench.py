#!/usr/bin/env python3
"""bench.py - env-steps/sec of the batched XarmPDPickAndPlace-v0 step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--task pick_and_place] [--envs-per-gpu 131072]
    python bench.py --impl reference ...      # the CPU arm: PyBullet when importable, else the oracle port, on all host cores

One "step" = one Env.step (action ingest, IK, 15 substeps of collide/dynamics/PGS, obs/reward/done, amortised
auto-resets) over one batch of `envs-per-gpu` envs per GPU, synthetic U(-1,1) actions.  N>1: launched by torchrun, one
rank per GPU, envs sharded with no data-path collective (weak scaling); NCCL only for the timing reduction and the
episode statistics.  Prints ONE JSON line from rank 0.

Workload phase: the envs' episode phases are STAGGERED (env i starts at step counter i mod episode length,
XarmVecEnv(stagger_phases=True)) and the population is pre-rolled for `--preroll` untimed steps after reset(), so that any
timed window - the driver's 20 steps as well as 200 - holds its fair share of time-limit endings, successes and auto-reset
passes.  The SYNCHRONISED schedule (all envs start together: one reset wave of every env per episode length) is timed
beside it over --sync-steps (200 = 4 episodes) and printed as `synchronised`.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {"reach": 333, "pick_and_place": 445, "stack_tower": 1001, "push_with_door": 601, "handover": 625}
ENV_ID = {"reach": "XarmReach-v0", "pick_and_place": "XarmPDPickAndPlace-v0", "stack_tower": "XarmPDStackTower-v0",
          "push_with_door": "XarmPDPushWithDoor-v0", "handover": "XarmPDHandover-v1"}
FP32_NOMINAL_TFLOPS = 74.4  # 148 SM x 128 lanes x 2 x 1.965 GHz; the measured figure comes from xarm_measure_fp32_peak
KERNEL_NAMES = {"setup": "k_pipe_setup", "light": "k_pipe_light", "heavy_rows": "k_heavy_rows", "heavy_solve": "k_heavy_solve2",
                "action": "k_pipe_action", "finish": "k_pipe_finish", "reset_stage": "k_pipe_reset_stage", "heavy_all": "k_pipe_heavy_all",
                "heavy_fused": "k_heavy_fused", "heavy_local": "k_pipe_heavy_local", "heavy_rec": "k_pipe_heavy_rec"}


def bench_config(task):
    cfg = {"reward_type": "sparse"}
    if task == "pick_and_place":
        cfg.update(num_obj=1, goal_shape="air", init_grasp_rate=0.0, goal_ground_rate=0.0)
    if task == "handover":   # BASELINE.json configs[4]: XarmPDHandover-v1 = the dense staged reward [REF benchmark/train.py:65-66]
        cfg.update(num_obj=1, goal_shape="ground", same_side_rate=0.5, use_stand=False, reward_type="dense")
    return cfg


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 8:
                self.rows.append(p)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for p in self.rows:
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_arm(task, seconds, min_steps):
    """the CPU path on all host cores (baseline/run_cpu_baseline.py): the real reference envs under PyBullet when pybullet + gym
    import (kind "pybullet"), else the oracle port (kind "port")"""
    from baseline import run_cpu_baseline as rcb
    cfg = {k: v for k, v in bench_config(task).items() if k != "use_stand"}
    return rcb.run(task, cfg, os.cpu_count() or 1, seconds, min_steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    task = args.task
    # blocks of steps until >= 5 s have elapsed and every worker has done >= max(2000, --steps) steps
    r = cpu_arm(task, 5.0, max(2000, args.steps))
    steps_per_worker = r["env_steps"] / r["cores"]
    line = {
        "impl": "reference", "metric": f"env-steps/sec (whole box) {ENV_ID[task]}", "value": r["value"], "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(steps_per_worker, 1.0),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{ENV_ID[task]} on the host CPU, one env per core ({r['cores']} workers, the SubprocVecEnv shape): "
                               + ("the reference's own env classes under PyBullet DIRECT" if r["kind"] == "pybullet" else
                                  "oracle port (float64 restatement of the PyBullet pipeline) - PyBullet is not installable offline"),
                   "env_id": ENV_ID[task], "steps_per_worker": steps_per_worker},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "pybullet_unavailable" in r:
        line["cpu_baseline"]["pybullet_unavailable"] = r["pybullet_unavailable"]
    print(json.dumps(line))


def kernel_source_hash():
    """sha256 over the kernel sources: a committed ncu traffic capture is only quoted for the sources it was taken from"""
    h = hashlib.sha256()
    for d in ("gym_xarm_b200/csrc", "include"):
        for f in sorted(os.listdir(os.path.join(ROOT, d))):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(ROOT, d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic_per_step(task, n):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one env step from profiles/traffic_<task>.json, written by
    tools/ncu_traffic.py from an `ncu --set full` capture of the substep kernels at this size.  Refused (None) when the capture
    was taken from other kernel sources than the ones this run uses, or at another size."""
    path = os.path.join(ROOT, "profiles", f"traffic_{task}.json")
    try:
        d = json.load(open(path))
    except Exception:  # noqa: BLE001
        return None, "no ncu traffic capture committed for this workload (tools/ncu_traffic.py)"
    if d.get("envs") != n:
        return None, f"capture is for {d.get('envs')} envs, this run has {n}"
    if d.get("kernel_source_hash") != kernel_source_hash():
        return None, f"capture {d.get('kernel_source_hash')} is stale: the kernel sources are {kernel_source_hash()} now (re-run tools/ncu_traffic.py)"
    return d["dram_bytes_per_step"], d.get("note", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", default="pick_and_place", choices=list(ALGO_BYTES))
    ap.add_argument("--envs-per-gpu", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stagger", action="store_true", help="headline on the synchronised schedule (all envs start together)")
    ap.add_argument("--preroll", type=int, default=-1, help="untimed steps after reset() (default: 1.2 episode lengths with stagger, 0 without)")
    ap.add_argument("--sync-steps", type=int, default=200, help="timed steps of the synchronised-schedule measurement (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed steps of the end-to-end arm (default: --steps)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: park fd 1 on stderr while libraries run (NCCL prints its version banner to
    # stdout at init) and restore it for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from gym_xarm_b200 import XarmVecEnv, _native, distributed as xd
    from gym_xarm_b200.specs import SPECS

    rank, world, local = xd.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    task = args.task
    n = args.envs_per_gpu or (65536 if task != "pick_and_place" else 131072)
    W, K = max(args.warmup, 3), args.steps
    ep_len = SPECS[task].max_episode_steps
    stagger = not args.no_stagger
    preroll = args.preroll if args.preroll >= 0 else (int(1.2 * ep_len) if stagger else 0)
    L = _native.load()
    fp32_peak = C.c_double(0.0)
    if rank == 0:
        _native.check(L.xarm_measure_fp32_peak(local, C.byref(fp32_peak)), "xarm_measure_fp32_peak")
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = [torch.rand(n, SPECS[task].act_dim, generator=g, device=dev) * 2 - 1 for _ in range(64)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed_run(stag, pre, w, k, profile):
        """reset -> `pre` untimed steps -> `w` warm-up steps -> `k` timed steps (CUDA events around every step on the launching
        stream, L2 flushed between steps).  Returns per-step ms, launches, kernel times, episode stats, wall time, clocks."""
        env = XarmVecEnv(task, n, config=bench_config(task), device=dev, seed=0, env_index_base=rank * n, auto_reset=True, stagger_phases=stag)
        env.reset()
        if profile:
            env.set_profiling(True)   # device-side %globaltimer stamps per pipeline launch (works inside the graph)
        if not args.no_graph:
            env.capture_graph()
        stream = env._stream if not args.no_graph else torch.cuda.current_stream(dev)

        def one_step(i, ev0=None, ev1=None):
            env.actions.copy_(ring[i % 64])
            flush.zero_()
            stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(stream):
                if ev0 is not None:
                    ev0.record(stream)
                _native.check(L.xarm_step(env._h, C.c_void_p(stream.cuda_stream)), "xarm_step")
                if ev1 is not None:
                    ev1.record(stream)
            torch.cuda.current_stream(dev).wait_stream(stream)

        for i in range(pre + w):
            one_step(i)
        torch.cuda.synchronize(dev)
        xd.barrier()
        torch.cuda.synchronize(dev)
        env.episode_stats()
        if profile:
            env.set_profiling(True)   # clear the per-kernel accumulators: only the timed region counts
        sampler = ClockSampler(local)
        if rank == 0 and profile:
            sampler.start()
        launches0 = L.xarm_launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        t0 = time.perf_counter()
        for i in range(k):
            one_step(pre + w + i, *evs[i])
        torch.cuda.synchronize(dev)
        xd.barrier()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if rank == 0 and profile else None
        out = {"ms": [a.elapsed_time(b) for a, b in evs], "launches": L.xarm_launch_count() - launches0,
               "ktimes": env.kernel_times() if profile else {}, "stats": env.episode_stats(), "wall": wall, "clocks": clocks,
               "act_dim": env.act_dim, "obs_dim": env.obs_dim, "goal_dim": env.goal_dim}
        env.close()
        return out

    R = timed_run(stagger, preroll, W, K, True)
    per_step_ms = R["ms"]
    dev_ms = sum(per_step_ms)
    ktimes = R["ktimes"]
    cnt = {(br, nm): c for (br, nm), (c, _) in ktimes.items() if nm.startswith("#")}
    heavy = sum(c for (br, nm), c in cnt.items() if nm == "#heavy_envs")
    allsub = sum(c for (br, nm), c in cnt.items() if nm == "#setup_envs")
    # per-rank rows (one small all_gather, outside the timed region): device time, slowest step, heavy / all env-substeps
    row = torch.tensor([dev_ms, max(per_step_ms), float(heavy), float(allsub)], dtype=torch.float64, device=dev)
    if world > 1:
        rows = torch.empty(world, 4, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(rows, row.unsqueeze(0).contiguous())
    else:
        rows = row.unsqueeze(0)
    rows = rows.cpu().tolist()
    dev_ms_max = max(r[0] for r in rows)
    stats = xd.gather_episode_stats(R["stats"], device=dev)
    total_env_steps = float(n) * K * world
    value = total_env_steps / (dev_ms_max * 1e-3)

    # the synchronised schedule beside it: all envs start together -> one reset wave per episode length inside the window
    sync = None
    if args.sync_steps > 0 and stagger:
        S = timed_run(False, 0, 20, args.sync_steps, False)
        s_ms = xd.max_over_ranks(sum(S["ms"]), device=dev)
        sync = {"value": float(n) * args.sync_steps * world / (s_ms * 1e-3), "unit": "env-steps/s", "steps": args.sync_steps, "warmup": 20,
                "ms_per_step": s_ms / args.sync_steps, "step_ms_p50": statistics.median(S["ms"]), "step_ms_max": max(S["ms"]),
                "note": "all envs start their episodes together (no stagger, no pre-roll): every 50th step is a reset wave of ~all envs"}

    # end-to-end arm: numpy in / numpy out through xarm_step_host (H2D of the actions, graph replay, D2H of obs / reward / flags and
    # of the finished envs' terminal observations inside), same staggered + pre-rolled phase and the same number of steps as `value`
    Ke = args.e2e_steps or K
    env_np = XarmVecEnv(task, n, config=bench_config(task), device=dev, seed=0, env_index_base=rank * n, auto_reset=True, output="numpy",
                        stagger_phases=stagger)
    env_np.reset()
    host_ring = [torch.empty(n, R["act_dim"]).uniform_(-1, 1).pin_memory().numpy() for _ in range(8)]
    for i in range(preroll + 3):
        env_np.step(host_ring[i % 8])
    xd.barrier()
    t0 = time.perf_counter()
    n_term = 0
    for i in range(Ke):
        _, _, d_, _ = env_np.step(host_ring[i % 8])
        n_term += int(d_.sum())
    torch.cuda.synchronize(dev)
    e2e_sec = xd.max_over_ranks(time.perf_counter() - t0, device=dev)
    e2e_value = float(n) * Ke * world / e2e_sec
    Wt = R["obs_dim"] + 2 * R["goal_dim"]
    h2d = n * R["act_dim"] * 4
    d2h = n * (Wt + 2) * 4 + 2 * n + 4 + int(n_term / max(Ke, 1)) * (Wt * 4 + 4)
    env_np.close()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        k_ms = dev_ms / K   # mean device time of one xarm_step over the timed region (CUDA events on the launching stream)
        # per-kernel device time inside the timed region (device-side %globaltimer stamps around every launch of the captured
        # graph; a CUDA event cannot be read inside a graph).  The step is a pipeline of small kernels on concurrent streams, so
        # the shares are of the summed kernel time, not of the wall time.
        real = {k_: v for k_, v in ktimes.items() if not k_[1].startswith("#")}
        tot_us = sum(v[1] for v in real.values()) or 1.0
        by_kernel = {}
        for (br, name), (c_, us) in real.items():
            e_ = by_kernel.setdefault(name, {"launches": 0, "total_us": 0.0})
            e_["launches"] += c_; e_["total_us"] += us
        kernels = [{"kernel": KERNEL_NAMES.get(k_, k_), "launches_with_work": v["launches"], "avg_us": v["total_us"] / max(v["launches"], 1),
                    "share_of_kernel_time": v["total_us"] / tot_us} for k_, v in sorted(by_kernel.items(), key=lambda kv: -kv[1]["total_us"])]
        branch_ms = {br: sum(us for (b_, _), (_, us) in real.items() if b_ == br) / 1e3 / K for br in "MEL"}
        dom = kernels[0] if kernels else {"kernel": "xarm_step", "avg_us": k_ms * 1e3}
        achieved = ALGO_BYTES[task] * n / (k_ms * 1e-3) / 1e9
        traffic, traffic_note = ncu_traffic_per_step(task, n)
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "xarm_step pipeline (dominant kernel: %s)" % dom["kernel"], "kernel_ms": k_ms,
                "algorithmic_bytes_per_env_step": ALGO_BYTES[task], "peak_source": peak_src,
                "dominant_kernel": dom, "kernels": kernels[:9], "kernel_ms_per_step_by_branch": branch_ms,
                "heavy_env_substep_fraction": {br: (cnt.get((br, "#heavy_envs"), 0) / cnt[(br, "#setup_envs")] if cnt.get((br, "#setup_envs")) else None)
                                               for br in "MEL"},
                "env_substeps_per_step": {br: cnt.get((br, "#setup_envs"), 0) / K for br in "MEL"},
                "note": "the step is FP32-issue/latency bound, not HBM bound (SURVEY.md 8d, DESIGN.md 4): state traffic is tiny, so the HBM "
                        "fraction is ~1e-4 by construction; see fp32 (oracle-counted FLOPs against the MEASURED FP32 peak) and the per-kernel "
                        "list (profiles/ holds the ncu captures of the same kernels); branches: M main, E envs that may finish + their "
                        "auto-reset passes, L late tail"}
        cb = None
        if not args.no_cpu_baseline and world == 1:   # the CPU arm is timed beside the GPU arm at N=1 only
            try:
                r_ = cpu_arm(task, 10.0, 2000)
                cb = {k_: r_[k_] for k_ in ("value", "unit", "cores", "kind", "sample")}
                if "pybullet_unavailable" in r_:
                    cb["pybullet_unavailable"] = r_["pybullet_unavailable"]
            except Exception as e:  # noqa: BLE001
                cb = {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        if not args.no_cpu_baseline:
            try:
                from oracle import oracle as orc
                e1 = orc.OracleEnv(task, seed=0, **{k: v for k, v in bench_config(task).items() if k != "use_stand"})
                e1.reset(); e1.flops()
                rr = np.random.default_rng(0)
                for _ in range(50):
                    _, _, d_, _ = e1.step(rr.uniform(-1, 1, e1.act_dim).astype(np.float32))
                    if d_:
                        e1.reset()
                fl = e1.flops() / 50
                tf = fl * n / (k_ms * 1e-3) / 1e12
                roof["fp32"] = {"oracle_flops_per_env_step": fl, "achieved_tflops": tf, "measured_peak_tflops": fp32_peak.value,
                                "frac_of_measured": tf / fp32_peak.value if fp32_peak.value else None,
                                "nominal_peak_tflops": FP32_NOMINAL_TFLOPS, "frac_of_nominal": tf / FP32_NOMINAL_TFLOPS,
                                "peak_how": "xarm_measure_fp32_peak: 8 independent FMA chains per thread, 2048 threads per SM, best of 5 x ~10 ms, CUDA events"}
            except Exception:  # noqa: BLE001
                pass
        slow = max(range(world), key=lambda r_: rows[r_][0])
        line = {
            "metric": f"env-steps/sec (whole box) {ENV_ID[task]}", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{ENV_ID[task]} batched, {n} envs per GPU x {world} GPU(s), {bench_config(task)['reward_type']} reward, auto-reset on, "
                                   f"U(-1,1) actions from a 64-batch device ring", "env_id": ENV_ID[task], "envs_per_gpu": n,
                       "phase": (f"staggered episode phases (env i starts at step i mod {ep_len}) + {preroll} untimed pre-roll steps after reset(): "
                                 "every timed step holds its share of time-limit endings, successes and auto-reset passes") if stagger else
                                "synchronised: all envs start their episodes together",
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)",
                       "timing": "CUDA events on the launching stream around every step, summed; max over ranks",
                       "cuda_graph": not args.no_graph, "mapping": "one thread per env (16 lanes per env in the coupled contact solver); step = kernel pipeline action -> 15 x {setup -> light | fused heavy (collision + rows + joint loop)} -> finish -> auto-reset passes; envs that may finish run as an early branch on reserved SMs"},
            "clocks": R["clocks"], "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                           "steps": Ke, "path": "XarmVecEnv(output='numpy').step -> xarm_step_host (pinned staging, CUDA-graph replay, "
                                                                 "terminal observations of the finished envs gathered on the device); same phase and step count as value"},
            "synchronised": sync,
            "gpu_launches": int(R["launches"]), "roofline": roof, "cpu_baseline": cb,
            "episode_stats": {k: stats[k] for k in ("episodes", "mean_return", "mean_length", "success_rate", "diverged")},
            "per_rank": {"ms_per_step": [r_[0] / K for r_ in rows], "slowest_step_ms": [r_[1] for r_ in rows],
                         "heavy_env_substep_fraction": [r_[2] / r_[3] if r_[3] else None for r_ in rows], "slow_rank": slow},
            "wall_s_timed_region": R["wall"],
            "step_ms_quantiles": {"min": min(per_step_ms), "p10": sorted(per_step_ms)[len(per_step_ms) // 10], "p50": statistics.median(per_step_ms),
                                  "p90": sorted(per_step_ms)[(9 * len(per_step_ms)) // 10], "max": max(per_step_ms)},
            "step_ms_first_60": [round(x, 2) for x in per_step_ms[:60]],
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
