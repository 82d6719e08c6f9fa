"""Multi-GPU host logic on CPU: index sharding and the episode-statistics gather with a
world_size-2 gloo group (the GPU path uses the same code over NCCL)."""
import os

import pytest

torch = pytest.importorskip("torch")


def _summary_part(n, n_div, conv_time, sum_c, mean_abs_e, nmae):
    """What rl4_stats_reduce writes for one rank (rl4afcs_b200/dist.py layout): per SPS field the sum over the
    non-diverged agents and over all agents, then n_kept, n_excluded."""
    from rl4afcs_b200 import _lib
    from rl4afcs_b200 import dist as rdist

    S = _lib.SPS
    part = torch.zeros(rdist.SUMMARY_LEN, dtype=torch.float64)
    kept = n - n_div
    for f, v in (("SUM_C", sum_c), ("CONV_TIME", conv_time), ("MEAN_ABS_E", mean_abs_e), ("NMAE", nmae),
                 ("UNSTEADY", 1.0 if conv_time > 30 else 0.0)):
        part[2 * S[f]] = kept * v
        part[2 * S[f] + 1] = n * v
    part[2 * S["DIVERGED"] + 1] = n_div
    part[2 * S["COUNT"]] = kept
    part[2 * S["COUNT"] + 1] = n_div
    return part


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from rl4afcs_b200 import dist as rdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rdist.shard_bounds(11, world, rank)          # 6 + 5 agents: unequal shards
    n = hi - lo
    part = _summary_part(n, 1 if rank == 1 else 0, 10.0 + 25.0 * rank, -1.0 - rank, 0.01, 0.05)
    out = rdist.gather_episode_summary(None, world, part=part)          # the function bench.py calls, under a process group
    mine = torch.arange(lo, hi, dtype=torch.float64)[:, None]
    per_agent = rdist.gather_per_agent(mine, world, n_total=11)
    per_agent2 = rdist.gather_per_agent(mine, world)                     # sizes exchanged first
    if rank == 0:
        q.put((out, per_agent.ravel().tolist(), per_agent2.ravel().tolist()))
    dist.destroy_process_group()


def test_shard_bounds_cover_the_index_exactly():
    from rl4afcs_b200 import dist as rdist

    for n in (0, 1, 7, 8, 1 << 20, (1 << 22) + 3):
        for w in (1, 2, 4, 8):
            b = [rdist.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_episode_summary_gather_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, per_agent, per_agent2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["agents"] == 11 and out["diverged"] == 1 and out["unsteady_convergence"] == 5
    assert abs(out["avg_c"] - (6 * -1.0 + 4 * -2.0) / 10) < 1e-12
    assert abs(out["avg_t"] - (6 * 10.0 + 4 * 35.0) / 10) < 1e-12 and abs(out["avg_t_all"] - (6 * 10.0 + 5 * 35.0) / 11) < 1e-12
    assert abs(out["avg_nmae"] - 0.05) < 1e-15 and abs(out["avg_abs_e"] - 0.01) < 1e-15
    assert per_agent == [float(i) for i in range(11)] and per_agent2 == per_agent
