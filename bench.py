#!/usr/bin/env python3
"""bench.py - env-steps/sec of the batched XarmPDPickAndPlace-v0 step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--task pick_and_place] [--envs-per-gpu 131072]
    python bench.py --impl reference ...      # the CPU arm: the oracle port on all host cores (PyBullet is not installable)

One "step" = one Env.step (action ingest, IK, 15 substeps of collide/dynamics/PGS, obs/reward/done, amortised
auto-resets) over one batch of `envs-per-gpu` envs per GPU, synthetic U(-1,1) actions.  N>1: launched by torchrun, one
rank per GPU, envs sharded with no data-path collective (weak scaling); NCCL only for the timing reduction and the
episode statistics.  Prints ONE JSON line from rank 0.
"""
import argparse
import json
import re
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
FP32_NOMINAL_TFLOPS = 74.4  # 148 SM x 128 lanes x 2 x 1.965 GHz (not in MEASURED_PEAKS.json)
KERNEL_NAMES = {"setup": "k_pipe_setup", "light": "k_pipe_light", "heavy_rows": "k_heavy_rows", "heavy_solve": "k_heavy_solve2",
                "action": "k_pipe_action", "finish": "k_pipe_finish", "reset_stage": "k_pipe_reset_stage"}


def bench_config(task):
    cfg = {"reward_type": "sparse"}
    if task == "pick_and_place":
        cfg.update(num_obj=1, goal_shape="air", init_grasp_rate=0.0, goal_ground_rate=0.0)
    if task == "handover":
        cfg.update(num_obj=1, goal_shape="ground", same_side_rate=0.5, use_stand=False)
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


def cpu_baseline(task, cores, env_steps_target):
    """The oracle port on `cores` host threads over a bounded sample of the same workload."""
    from oracle import oracle as orc
    n_envs = cores * 8
    steps = max(50, int(env_steps_target // n_envs))
    done, sec = orc.bench(task, n_envs, steps, cores, seed=0, **{k: v for k, v in bench_config(task).items() if k != "use_stand"})
    return {"value": done / sec, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_envs} envs x {steps} steps of {ENV_ID[task]} (auto-reset, U(-1,1) actions) on {cores} threads, {sec:.1f} s; "
                      "oracle stand-in for PyBullet, which is not installable offline"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    task = args.task
    n_envs = cores * 8
    cfg = {k: v for k, v in bench_config(task).items() if k != "use_stand"}
    t_w = 0.0
    if args.warmup > 0:
        _, t_w = orc.bench(task, n_envs, args.warmup, cores, seed=1, **cfg)
    done, sec = orc.bench(task, n_envs, args.steps, cores, seed=0, **cfg)
    value = done / sec
    line = {
        "impl": "reference", "metric": f"env-steps/sec (whole box) {ENV_ID[task]}", "value": value, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{ENV_ID[task]} oracle port (double precision restatement of the PyBullet pipeline), "
                               f"{n_envs} envs per step on {cores} host threads", "env_id": ENV_ID[task]},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n_envs} envs x {args.steps} steps, {sec:.1f} s (+{t_w:.1f} s warm-up); PyBullet itself is not installable offline"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



def ncu_traffic_per_step(task, n):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one env step, from the committed `ncu --set full` capture of
    the four substep kernels of a full-size main pass (profiles/r1c_ncu_full_main_pass_kernels_summary.txt: one launch each of
    setup / heavy_rows / heavy_solve2 / light at 131 072 PickAndPlace envs) x the 15 substeps of a step.  None when the capture
    does not describe this workload.  It is implementation traffic (thread-local link arrays and solver records that spill
    through L2), not the 445 algorithmic bytes per env-step."""
    try:
        if task != "pick_and_place" or n != 131072:
            return None, "no ncu capture for this workload"
        path = os.path.join(ROOT, "profiles", "r1c_ncu_full_main_pass_kernels_summary.txt")
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total, launches = 0.0, 0
        for line in open(path):
            got = re.findall(r"dram_(?:rd|wr)=([0-9.]+)([KMG]?byte)", line)
            if got:
                total += sum(float(v) * unit[u] for v, u in got)
                launches += 1
        if launches != 4:
            return None, "capture not understood"
        return 15.0 * total, ("15 substeps x one ncu --set full launch each of k_pipe_setup, k_heavy_rows, k_heavy_solve2, k_pipe_light "
                              "(profiles/r1c_ncu_full_main_pass_kernels_summary.txt, 131 072 envs, two steps after a reset); implementation "
                              "traffic per step, to compare with algorithmic_bytes_per_env_step x envs")
    except Exception as e:  # noqa: BLE001
        return None, f"unavailable: {e}"


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
    ap.add_argument("--e2e-steps", type=int, default=30)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly one JSON line: park fd 1 on stderr while libraries run (NCCL prints its version banner to
    # stdout at init) and restore it for the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    from gym_xarm_b200 import XarmVecEnv, _native, distributed as xd

    rank, world, local = xd.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    task = args.task
    n = args.envs_per_gpu or (65536 if task != "pick_and_place" else 131072)
    W, K = max(args.warmup, 3), args.steps
    env = XarmVecEnv(task, n, config=bench_config(task), device=dev, seed=0, env_index_base=rank * n, auto_reset=True)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = [torch.rand(n, env.act_dim, generator=g, device=dev) * 2 - 1 for _ in range(64)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    env.reset()
    env.set_profiling(True)   # device-side %globaltimer stamps per pipeline launch (works inside the graph)
    if not args.no_graph:
        env.capture_graph()
    stream = env._stream if not args.no_graph else torch.cuda.current_stream(dev)
    L = _native.load()

    def one_step(i, ev0=None, ev1=None):
        env.actions.copy_(ring[i % 64])
        flush.zero_()
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            if ev0 is not None:
                ev0.record(stream)
            _native.check(L.xarm_step(env._h, __import__("ctypes").c_void_p(stream.cuda_stream)), "xarm_step")
            if ev1 is not None:
                ev1.record(stream)
        torch.cuda.current_stream(dev).wait_stream(stream)

    for i in range(W):
        one_step(i)
    torch.cuda.synchronize(dev)
    xd.barrier()
    torch.cuda.synchronize(dev)
    env.episode_stats()
    env.set_profiling(True)   # clear the per-kernel accumulators: only the timed region counts
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.xarm_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for i in range(K):
        one_step(W + i, *evs[i])
    torch.cuda.synchronize(dev)
    xd.barrier()
    torch.cuda.synchronize(dev)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = L.xarm_launch_count() - launches0
    ktimes = env.kernel_times()
    per_step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = sum(per_step_ms)
    dev_ms_max = xd.max_over_ranks(dev_ms, device=dev)
    stats = xd.gather_episode_stats(env.episode_stats(), device=dev)
    total_env_steps = float(n) * K * world
    value = total_env_steps / (dev_ms_max * 1e-3)

    # end-to-end arm: numpy in / numpy out through xarm_step_host (H2D of the actions, D2H of obs/reward/done inside)
    env_np = XarmVecEnv(task, n, config=bench_config(task), device=dev, seed=0, env_index_base=rank * n, auto_reset=True, output="numpy")
    env_np.reset()
    host_ring = [torch.empty(n, env.act_dim).uniform_(-1, 1).pin_memory().numpy() for _ in range(8)]
    for i in range(3):
        env_np.step(host_ring[i % 8])
    xd.barrier()
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        env_np.step(host_ring[i % 8])
    torch.cuda.synchronize(dev)
    e2e_sec = xd.max_over_ranks(time.perf_counter() - t0, device=dev)
    e2e_value = float(n) * args.e2e_steps * world / e2e_sec
    h2d = n * env.act_dim * 4
    d2h = n * (env.obs_dim + 2 * env.goal_dim + 2) * 4 + 2 * n
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
        tot_us = sum(v[1] for v in ktimes.values()) or 1.0
        by_kernel = {}
        for (br, name), (cnt, us) in ktimes.items():
            e_ = by_kernel.setdefault(name, {"launches": 0, "total_us": 0.0})
            e_["launches"] += cnt; e_["total_us"] += us
        kernels = [{"kernel": KERNEL_NAMES.get(k_, k_), "launches_with_work": v["launches"], "avg_us": v["total_us"] / max(v["launches"], 1),
                    "share_of_kernel_time": v["total_us"] / tot_us} for k_, v in sorted(by_kernel.items(), key=lambda kv: -kv[1]["total_us"])]
        branch_ms = {br: sum(us for (b_, _), (_, us) in ktimes.items() if b_ == br) / 1e3 / K for br in "MEL"}
        dom = kernels[0] if kernels else {"kernel": "xarm_step", "avg_us": k_ms * 1e3}
        achieved = ALGO_BYTES[task] * n / (k_ms * 1e-3) / 1e9
        traffic, traffic_note = ncu_traffic_per_step(task, n)
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "xarm_step pipeline (dominant kernel: %s)" % dom["kernel"], "kernel_ms": k_ms,
                "algorithmic_bytes_per_env_step": ALGO_BYTES[task], "peak_source": peak_src,
                "dominant_kernel": dom, "kernels": kernels[:8], "kernel_ms_per_step_by_branch": branch_ms,
                "note": "the step is FP32-issue/latency bound, not HBM bound (SURVEY.md 8d, DESIGN.md 4): state traffic is tiny, so the HBM "
                        "fraction is ~1e-4 by construction; see fp32 (oracle-counted FLOPs against the nominal FP32 peak) and the per-kernel "
                        "list (profiles/ holds the ncu captures of the same kernels)"}
        cb = None
        if not args.no_cpu_baseline and world == 1:   # the CPU arm is timed beside the GPU arm at N=1 only
            try:
                cb = cpu_baseline(task, os.cpu_count() or 1, 300000)
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
                roof["fp32"] = {"oracle_flops_per_env_step": fl, "achieved_tflops": fl * n / (k_ms * 1e-3) / 1e12,
                                "nominal_peak_tflops": FP32_NOMINAL_TFLOPS, "frac_of_nominal": fl * n / (k_ms * 1e-3) / 1e12 / FP32_NOMINAL_TFLOPS}
            except Exception:  # noqa: BLE001
                pass
        line = {
            "metric": f"env-steps/sec (whole box) {ENV_ID[task]}", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{ENV_ID[task]} batched, {n} envs per GPU x {world} GPU(s), num_obj=1, sparse reward, auto-reset on, "
                                   f"U(-1,1) actions from a 64-batch device ring", "env_id": ENV_ID[task], "envs_per_gpu": n,
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pairs)",
                       "timing": "CUDA events on the launching stream around every step, summed; max over ranks",
                       "cuda_graph": not args.no_graph, "mapping": "one thread per env (16 lanes per env in the coupled contact solver); step = kernel pipeline action -> 15 x {setup -> light | heavy_rows -> heavy_solve} -> finish -> auto-reset passes; envs that may finish run as an early branch on 32 reserved SMs"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "steps": args.e2e_steps, "path": "XarmVecEnv(output='numpy').step -> xarm_step_host (pinned staging)"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cb,
            "episode_stats": {k: stats[k] for k in ("episodes", "mean_return", "mean_length", "success_rate", "diverged")},
            "wall_s_timed_region": t_wall,
            "step_ms_quantiles": {"min": min(per_step_ms), "p10": sorted(per_step_ms)[len(per_step_ms) // 10], "p50": statistics.median(per_step_ms),
                                  "p90": sorted(per_step_ms)[(9 * len(per_step_ms)) // 10], "max": max(per_step_ms)},
            "step_ms_first_60": [round(x, 2) for x in per_step_ms[:60]],
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    env.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
