"""Multi-GPU plumbing: envs shard embarrassingly - one process per GPU, one contiguous slab of envs per
rank, RNG streams keyed by the GLOBAL env index so results do not depend on the number of ranks.  No collective runs on
the step path; the only exchange is an all_gather of five episode-statistics doubles per rank (NCCL over NVLink on
GPUs, gloo in CPU tests)."""
import os

import torch
import torch.distributed as dist


def slab(total_envs, rank, world_size):
    """[first, count) of the global env range owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def init_from_env(backend=None):
    """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world_size, local_rank)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # NCCL writes its banner ("NCCL version ...") to stdout when NCCL_DEBUG is VERSION or above: bench.py's stdout is one
            # JSON line, so the banner goes to stderr
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_episode_stats(stats, device=None):
    """all_gather of {episodes, return_sum, length_sum, success_sum, diverged} -> job-wide totals + per-rank rows."""
    keys = ("episodes", "return_sum", "length_sum", "success_sum", "diverged")
    row = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        rows = torch.empty(dist.get_world_size(), len(keys), dtype=torch.float64, device=row.device)
        dist.all_gather_into_tensor(rows, row.unsqueeze(0).contiguous())
    else:
        rows = row.unsqueeze(0)
    total = rows.sum(0)
    out = {k: float(total[i]) for i, k in enumerate(keys)}
    ep = max(out["episodes"], 1.0)
    out["mean_return"] = out["return_sum"] / ep
    out["mean_length"] = out["length_sum"] / ep
    out["success_rate"] = out["success_sum"] / ep
    out["per_rank"] = rows.cpu().tolist()
    return out


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
