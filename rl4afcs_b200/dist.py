"""Multi-GPU layout of the path: agents are independent (functions.py:39-42 builds a fresh
IDHPsp per seed), so GPU g simply owns the contiguous block [g*B/G, (g+1)*B/G) of the agent
index.  No collective touches the data path; the only exchange is one all-gather of episode
statistics at the end (replaces the pickled result dicts of functions.py:131-166).

The per-rank numbers come from the CUDA kernels behind ``rl4_sp_agent_stats`` / ``rl4_stats_reduce``
(include/rl4afcs_b200.h); this module only moves them between ranks (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch

from . import _lib

# layout of the per-rank summary produced by rl4_stats_reduce over the SPS planes:
#   [2 f] = sum of field f over the non-diverged agents, [2 f + 1] = over all agents, then n_kept, n_excluded
_S = _lib.SPS
SUMMARY_LEN = 2 * _S["COUNT"] + 2


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous block of the agent index owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def episode_summary_tensor(engine, n_steps=None) -> torch.Tensor:
    """Per-rank partial sums for the MC_run metrics (functions.py:161-227) as ``SUMMARY_LEN`` float64 numbers on the
    engine's GPU: one ``rl4_sp_agent_stats`` launch + one deterministic ``rl4_stats_reduce`` (diverged runs excluded
    from the 'kept' sums, functions.py:176-178)."""
    import ctypes

    L = engine.lib
    planes = engine.stats_planes(n_steps)                  # view (COUNT, n) of a (COUNT, stride) buffer
    out = torch.empty(SUMMARY_LEN, dtype=torch.float64, device=engine.device)
    nwork = int(L.rl4_stats_reduce_work_doubles())
    work = torch.empty(nwork, dtype=torch.float64, device=engine.device)
    with torch.cuda.device(engine.device):
        rc = L.rl4_stats_reduce(planes.data_ptr(), engine.stride, _S["COUNT"],
                                planes[_S["DIVERGED"]].data_ptr(), engine.n, out.data_ptr(), work.data_ptr(), nwork,
                                ctypes.c_void_p(torch.cuda.current_stream(engine.device).cuda_stream))
        _lib.check(rc, "rl4_stats_reduce")
    return out


def reduce_summary(parts: torch.Tensor) -> dict:
    """parts: (world, SUMMARY_LEN) -> the metrics of MC_run (functions.py:223-227) over all ranks.  ``avg_c`` averages the
    non-diverged runs; ``avg_t`` does the same here (the reference's own ``avg_t`` runs over ALL runs because of its
    second ``np.delete`` quirk -- that value is ``avg_t_all``); ``avg_nmae`` = mean over the non-diverged runs of
    mean|e| / (max ref - min ref)."""
    tot = parts.to(torch.float64).sum(dim=0).tolist()
    kept = lambda f: tot[2 * _S[f]]            # noqa: E731
    every = lambda f: tot[2 * _S[f] + 1]       # noqa: E731
    n_kept, n_excl = tot[2 * _S["COUNT"]], tot[2 * _S["COUNT"] + 1]
    n = n_kept + n_excl
    ok = max(n_kept, 1.0)
    return {"agents": int(n), "diverged": int(n_excl), "unsteady_convergence": int(every("UNSTEADY")),
            "avg_c": kept("SUM_C") / ok, "avg_t": kept("CONV_TIME") / ok, "avg_t_all": every("CONV_TIME") / max(n, 1.0),
            "avg_abs_e": kept("MEAN_ABS_E") / ok, "avg_nmae": kept("NMAE") / ok}


def gather_episode_summary(engine, world: int, group=None, part: torch.Tensor = None) -> dict:
    """One all-gather of the per-rank summary (NCCL over NVLink when world > 1).  ``part`` overrides the summary of this
    rank (the CPU tests pass a hand-made one; on a GPU it is ``episode_summary_tensor(engine)``)."""
    if part is None:
        part = episode_summary_tensor(engine)
    part = part.contiguous()
    if world > 1:
        import torch.distributed as dist

        flat = torch.empty(world * part.numel(), dtype=part.dtype, device=part.device)   # concatenated form: NCCL and gloo both take it
        dist.all_gather_into_tensor(flat, part.reshape(-1), group=group)
        parts = flat.reshape(world, part.numel())
    else:
        parts = part[None]
    return reduce_summary(parts)


def gather_per_agent(t: torch.Tensor, world: int, group=None, n_total: int = None) -> torch.Tensor:
    """All-gather a per-agent tensor (agents, k) whose shards follow ``shard_bounds`` (they differ by at most one agent
    when ``n_total % world != 0``): every shard is padded to the largest one, gathered with ONE collective and trimmed."""
    if world == 1:
        return t
    import torch.distributed as dist

    rows = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    if n_total is None:                                  # sizes unknown to the caller: exchange them first
        sizes = torch.empty(world, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(sizes, rows, group=group)
        sizes = [int(s) for s in sizes.tolist()]
    else:
        sizes = [b - a for a, b in (shard_bounds(n_total, world, r) for r in range(world))]
    big = max(sizes)
    padded = t.contiguous()
    if t.shape[0] < big:
        padded = torch.cat([padded, padded.new_zeros((big - t.shape[0],) + tuple(t.shape[1:]))], dim=0)
    out = torch.empty((world * big,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(s == big for s in sizes):
        return out
    return torch.cat([out[r * big: r * big + sizes[r]] for r in range(world)], dim=0)
