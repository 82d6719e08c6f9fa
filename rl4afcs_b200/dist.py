"""Multi-GPU layout of the path: agents are independent (functions.py:39-42 builds a fresh
IDHPsp per seed), so GPU g simply owns the contiguous block [g*B/G, (g+1)*B/G) of the agent
index.  No collective touches the data path; the only exchange is one all-gather / all-reduce of
episode statistics at the end (replaces the pickled result dicts of functions.py:131-166).
"""
from __future__ import annotations

import torch


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous block of the agent index owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def episode_summary_tensor(stats: dict) -> torch.Tensor:
    """Per-rank partial sums for the MC_run metrics (functions.py:161-227): [n, n_diverged,
    n_unsteady(converged_time > 30 s), sum(sum_c over non-diverged), sum(converged_time over
    non-diverged), sum(mean_abs_e over non-diverged)] as float64."""
    div = stats["diverged"]
    ok = ~div
    conv = stats["converged_time"]
    vals = [
        torch.tensor(float(div.numel()), dtype=torch.float64, device=div.device),
        div.sum().double(),
        (conv > 30.0).sum().double(),
        torch.where(ok, stats["sum_c"], torch.zeros_like(stats["sum_c"])).sum(),
        torch.where(ok, conv, torch.zeros_like(conv)).sum(),
        torch.where(ok, stats["mean_abs_e"], torch.zeros_like(stats["mean_abs_e"])).sum(),
    ]
    return torch.stack(vals)


def reduce_summary(parts: torch.Tensor) -> dict:
    """parts: (world, 6) -> the metrics dict of MC_run (functions.py:223-227) over all ranks."""
    tot = parts.sum(dim=0)
    n, nd = float(tot[0]), float(tot[1])
    ok = max(n - nd, 1.0)
    return {"agents": int(n), "diverged": int(nd), "unsteady_convergence": int(tot[2]),
            "avg_c": float(tot[3]) / ok, "avg_t": float(tot[4]) / ok, "avg_abs_e": float(tot[5]) / ok}


def gather_episode_summary(engine, world: int, group=None) -> dict:
    """One all-gather of the 6-number per-rank summary (NCCL over NVLink when world > 1)."""
    part = episode_summary_tensor(engine.stats())
    if world > 1:
        import torch.distributed as dist

        parts = torch.empty((world, part.numel()), dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(parts, part.contiguous(), group=group)
    else:
        parts = part[None]
    return reduce_summary(parts)


def gather_per_agent(t: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """All-gather a per-agent statistics tensor (agents, k) from equal-sized shards."""
    if world == 1:
        return t
    import torch.distributed as dist

    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out
