"""Multi-GPU layout of the path: independent utterances, sharded by index, no collective.

Every utterance's result depends only on its own samples and its own plan, so a batch of B utterances is split
into contiguous index ranges, one per rank (one process per GPU). ``torch.distributed`` is used only for the
barrier and the max-over-ranks reduction of the timing, never on the data path.
"""
from __future__ import annotations

import os
from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of ``total`` items owned by ``rank``; sizes differ by at most one, low ranks first."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched plainly."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str):
    """Join the job's process group (MASTER_ADDR/PORT from the environment, 127.0.0.1 by default)."""
    import torch.distributed as dist
    rank, _, world = env_rank_world()
    if world == 1 or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    kw = {}
    if backend == "nccl":  # bind the communicator to this rank's GPU (LOCAL_RANK) instead of letting NCCL guess
        import torch
        kw["device_id"] = torch.device("cuda", env_rank_world()[1])
    dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX all-reduce of a scalar (the step time); identity for a single process."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
