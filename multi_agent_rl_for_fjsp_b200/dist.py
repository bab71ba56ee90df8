"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous global index, NO collective in the step.

The only collectives are (a) the max-over-ranks / sum-over-ranks of timings and counters (bench, stats) and (b) the
A2C gradient all-reduce in ``a2c_batched.py``.  Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total_envs: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of the global env index range for `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(gpu_index: int) -> list:
    """Pin this process to the CPUs NVML reports as local to `gpu_index`, so pinned host buffers allocated afterwards
    are first-touched on the GPU's own NUMA node (matters when 8 ranks stream results to the host at once).
    Returns the CPU list it bound to ([] if NVML or the affinity call is unavailable)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001
        return []


def init(backend: str | None = None, device: torch.device | None = None):
    rank, local_rank, world = world_info()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if device is not None and device.type == "cuda":
            kw["device_id"] = device
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), **kw)
    return rank, local_rank, world


def _reduce(x: float, op, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=op)
    return float(t.item())


def max_over_ranks(x: float, device=None) -> float:
    """Timings are reported as the MAX over ranks (never a wall clock of one rank)."""
    return _reduce(x, dist.ReduceOp.MAX, device)


def sum_over_ranks(x: float, device=None) -> float:
    return _reduce(x, dist.ReduceOp.SUM, device)


def barrier(device=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def make_sharded_env(total_envs: int, seed: int, num_orders: int = 30, config=None, autoreset: bool = True, **kw):
    """The local shard of a global batch of `total_envs` envs on this rank's GPU (cuda:LOCAL_RANK)."""
    from .env import BatchedFJSPEnv

    rank, local_rank, world = world_info()
    lo, hi = shard_range(total_envs, rank, world)
    return BatchedFJSPEnv(hi - lo, config=config, device="cuda:%d" % local_rank, first_env=lo, seed=seed,
                          num_orders=num_orders, autoreset=autoreset, **kw)
