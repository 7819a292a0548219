"""Multi-GPU plumbing: stereo pairs are independent units (CSPatchMatch and the
plane cost are per-pair objects, main.cc:96-116), so pair i goes to rank
i mod world with NO data-path collective.  torch.distributed (NCCL on GPUs,
gloo in the CPU tests) is used only for the barrier and for the two scalars of
the throughput line: sum of units, max of elapsed time."""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def pairs_for_rank(n_pairs: int, world: int, rank: int) -> list[int]:
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_pairs, world))


def aggregate(units_local: float, seconds_local, device=None):
    """Whole-job (units, seconds): units summed over ranks, every entry of
    seconds_local maxed over ranks.  Works uninitialised (world 1)."""
    import torch
    import torch.distributed as dist

    secs = list(seconds_local) if isinstance(seconds_local, (list, tuple)) else [seconds_local]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(units_local), [float(s) for s in secs]
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    t = torch.tensor([float(s) for s in secs], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), [float(x) for x in t.tolist()]
