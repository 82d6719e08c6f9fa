"""Multi-GPU host logic on CPU: index sharding and the episode-statistics gather with a
world_size-2 gloo group (the GPU path uses the same code over NCCL)."""
import os

import pytest

torch = pytest.importorskip("torch")


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from rl4afcs_b200 import dist as rdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rdist.shard_bounds(10, world, rank)
    n = hi - lo
    stats = {"diverged": torch.tensor([False] * (n - 1) + [rank == 1]),
             "converged_time": torch.full((n,), 10.0 + 25.0 * rank, dtype=torch.float64),
             "sum_c": torch.full((n,), -1.0 - rank, dtype=torch.float64),
             "mean_abs_e": torch.full((n,), 0.01, dtype=torch.float64)}
    part = rdist.episode_summary_tensor(stats)
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part)
    out = rdist.reduce_summary(torch.stack(parts))
    per_agent = rdist.gather_per_agent(torch.arange(lo, hi, dtype=torch.float64)[:, None], world) if n * world == 10 else None
    if rank == 0:
        q.put((out, None if per_agent is None else per_agent.ravel().tolist()))
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
    out, per_agent = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out["agents"] == 10 and out["diverged"] == 1 and out["unsteady_convergence"] == 5
    assert abs(out["avg_c"] - (5 * -1.0 + 4 * -2.0) / 9) < 1e-12
    assert per_agent == [float(i) for i in range(10)]
