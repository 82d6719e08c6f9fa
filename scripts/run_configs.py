"""BASELINE.json configs[3] and configs[4] as runnable jobs (single GPU or torchrun, one rank per GPU).

    python scripts/run_configs.py --config sweep   [--agents-total 4194304]     # hyper-parameter sweep grid, linear plant
    python scripts/run_configs.py --config faults  [--agents-total 8388608]     # Monte-Carlo fault study, nonlinear plant
    python scripts/run_configs.py --config nonlinear [--integrator rk4|ode5]    # configs[2]: 256K agents, 90 s flights
    torchrun --nproc-per-node 8 scripts/run_configs.py --config sweep

sweep  (SURVEY 8d config 4): grid over eta_a_h in [2.5,4.7] x eta_c_h in [0.45,0.55] x rls_gamma in [0.99,1.0] x reference
       amplitude in [1,10] deg (the reference has no excitation signal; the amplitude of the tracked sine stands in),
       flattened onto the agent index and sharded contiguously over the GPUs.
faults (config 5): per-agent fault family sampled from the nine names of envs/nonlinear/env.py:134-158, fault time
       U(30,70) s, damping factor U(0.2,0.5), c.g. shift U(-0.5,0); statistics gathered with one NCCL all-gather.
Prints one JSON line with throughput (CUDA events, max over ranks) and the gathered statistics."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rl4afcs_b200 import _lib, dist as rdist, nl_engine, sp_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", required=True, choices=["sweep", "faults", "nonlinear"])
ap.add_argument("--integrator", default="rk4", choices=["rk4", "ode5"])
ap.add_argument("--plant", default="surrogate", choices=["surrogate", "dasmat"], help="nonlinear config: calibrated stand-in or the reference's own model")
ap.add_argument("--agents-total", type=int, default=None)
ap.add_argument("--steps", type=int, default=None)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.stdout.flush(); _saved = os.dup(1); os.dup2(2, 1)          # NCCL's version banner goes to stderr
    dist.init_process_group("nccl", device_id=torch.device(dev)); dist.barrier(); torch.cuda.synchronize()
    os.dup2(_saved, 1); os.close(_saved)


def timed(fn):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


if a.config == "sweep":
    total = a.agents_total or (1 << 22)
    steps = a.steps or 3000
    lo, hi = rdist.shard_bounds(total, world, rank)
    n = hi - lo
    idx = np.arange(lo, hi)
    # grid 64 x 64 x 32 x 32 (scaled down proportionally for smaller totals), flattened row-major
    dims = [64, 64, 32, 32]
    while np.prod(dims) > total:
        dims[int(np.argmax(dims))] //= 2
    i3 = idx % dims[3]; i2 = (idx // dims[3]) % dims[2]; i1 = (idx // (dims[3] * dims[2])) % dims[1]; i0 = (idx // (dims[3] * dims[2] * dims[1])) % dims[0]
    lin = lambda i, d, a_, b_: a_ + (b_ - a_) * (i / max(d - 1, 1))  # noqa: E731
    eng = sp_engine.SpEngine(n, policy="mixed", device=dev)
    import bench
    sp_engine.apply_idhp_config(eng, bench.default_idhp_config(), dt=0.02)
    eng.set_hp("ETA_A_H", lin(i0, dims[0], 2.5, 4.7)); eng.set_hp("ETA_C_H", lin(i1, dims[1], 0.45, 0.55))
    eng.set_hp("RLS_GAMMA", lin(i2, dims[2], 0.99, 1.0)); eng.set_hp("REF_AMP", np.deg2rad(lin(i3, dims[3], 1.0, 10.0)))
    eng.set_hpi("FAULT_STEP", -1); eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(bench.reference_table(steps))
    w = sp_engine.truncated_normal_weights(n, 7, 0.1, dev)      # same initial weights draw for every grid point block
    eng.init(torch.zeros((n, 2), dtype=torch.float64, device=dev), w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    ms = timed(lambda: eng.run(steps))
    summary = rdist.gather_episode_summary(eng, world)
    out = {"config": "hyper-parameter sweep grid (BASELINE.json configs[3])", "grid": dims, "agents_total": total, "n_gpus": world,
           "steps": steps, "policy": "mixed", "seconds": ms * 1e-3, "agent_steps_per_s": total * steps / (ms * 1e-3), "stats": summary}
elif a.config == "nonlinear":
    # BASELINE.json configs[2]: nonlinear aircraft IDHP attitude tracking, 256K agents, the 90 s flight of idhp_nonlin.py
    # (hyper-parameters :123-146, no fault), through the reference-shaped objects: Ce500NonLinear + IDHPnonlin.train()
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear
    from rl4afcs_b200.objects import IDHPnonlin

    total = a.agents_total or (1 << 18)
    steps = a.steps or 9000
    lo, hi = rdist.shard_bounds(total, world, rank)
    n = hi - lo
    th = nl_engine.theta_reference()
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    env_config = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01, "t_end": steps * 0.01,
                  "total_steps": steps, "fault_time": 60, "fault_scenario": "none",
                  "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    idhp_config = {"gamma": 0.6, "multistep": 0, "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": 0.95, "kappa": [1, 2, 1],
                   "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 4.0, "error_thresh": 1, "tau": 0.02, "in_dims": 4,
                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 35.0, "eta_l": 5.0, "elig": "accumulating"},
                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.4, "eta_l": 0.7, "elig": 1233},
                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}
    env = Ce500NonLinear(env_config, batch=n, device=dev, dtype="mixed", integrator="ode5" if a.plant == "dasmat" else a.integrator, plant=a.plant)
    idhp = IDHPnonlin(env, idhp_config, seed=8 + rank, verbose=0, log=None, chunk=1000)
    idhp.train(2)                                             # loads the kernels (lazy module loading) outside the timed region
    ms = timed(lambda: idhp.train(steps))                     # reset (trim) + prologue + the fused launches + noise draws
    st = idhp.stats()
    alive = ~st["diverged"]
    rse = st["rse"][:, 0][alive]
    flight = st["rse_flight"][:, 0][alive]
    part = torch.stack([torch.tensor(float(n), device=dev, dtype=torch.float64), st["diverged"].sum().double(), rse.sum(), flight.sum(),
                        st["nz_peak"][alive].max(), rse.median(), flight.median()])
    parts = rdist.gather_per_agent(part[None], world)
    ok = parts[:, 0].sum() - parts[:, 1].sum()
    out = {"config": "nonlinear aircraft IDHP attitude tracking (BASELINE.json configs[2])", "agents_total": total, "n_gpus": world,
           "steps": steps, "policy": "mixed", "plant": a.plant, "integrator": "ode5 (the model's own)" if a.plant == "dasmat" else a.integrator, "seconds": ms * 1e-3,
           "agent_steps_per_s": total * steps / (ms * 1e-3),
           "stats": {"agents": int(parts[:, 0].sum()), "diverged": int(parts[:, 1].sum()),
                     "mean_RSE_theta_per_step_deg": float(np.rad2deg(float(parts[:, 2].sum() / ok) / steps)),
                     "mean_RSE_theta_per_step_flight_deg": float(np.rad2deg(float(parts[:, 3].sum() / ok) / max(1, steps - 5500))),
                     "median_RSE_theta_per_step_deg_rank0": float(np.rad2deg(float(parts[0, 5]) / steps)),
                     "peak_nz": float(parts[:, 4].max())}}
else:
    total = a.agents_total or (1 << 20)
    steps = a.steps or 9000
    lo, hi = rdist.shard_bounds(total, world, rank)
    n = hi - lo
    rng = np.random.default_rng(1000 + rank)
    names = ["damp_elevator", "damp_aileron", "damp_rudder", "damp_all", "shift_cg", "slow_all", "saturate_elevator",
             "saturate_aileron", "saturate_rudder"]
    pick = rng.integers(0, len(names), n)
    ds = np.asarray([nl_engine.split_fault(names[p]) for p in pick], dtype=np.int32)
    eng = nl_engine.NlEngine(n, policy="mixed", device=dev)
    eng.set_hpi("FAULT_DAMP", ds[:, 0]); eng.set_hpi("FAULT_SAT", ds[:, 1])
    eng.set_hpi("FAULT_STEP", (rng.uniform(30, 70, n) / 0.01).astype(np.int32))
    eng.set_hp("DAMP_FACTOR", rng.uniform(0.2, 0.5, n)); eng.set_hp("CG_SHIFT", rng.uniform(-0.5, 0.0, n))
    eng.set_reference(nl_engine.theta_reference())
    g = torch.Generator(device=dev); g.manual_seed(11 + rank)
    wd = lambda k: (torch.randn((n, k), generator=g, device=dev).clamp_(-2, 2) * 0.1).double()  # noqa: E731
    eng.init(wd(40), wd(10), wd(40), wd(30))
    chunk = 500

    def episode():
        k = 0
        while k < steps:
            c = min(chunk, steps - k)
            eng.run(c, torch.randn((c, n), generator=g, device=dev, dtype=torch.float32)); k += c
    ms = timed(episode)
    st = eng.stats()
    part = torch.stack([torch.tensor(float(n), device=dev, dtype=torch.float64), st["diverged"].sum().double(),
                        st["rse"][:, 0].nan_to_num().sum(), st["nz_peak"].nan_to_num().max()])
    parts = rdist.gather_per_agent(part[None], world)
    out = {"config": "Monte-Carlo fault study, nonlinear plant (BASELINE.json configs[4])", "agents_total": total, "n_gpus": world,
           "steps": steps, "policy": "mixed", "seconds": ms * 1e-3, "agent_steps_per_s": total * steps / (ms * 1e-3),
           "stats": {"agents": int(parts[:, 0].sum()), "diverged": int(parts[:, 1].sum()),
                     "mean_cumulative_RSE_theta": float(parts[:, 2].sum() / parts[:, 0].sum()), "peak_nz": float(parts[:, 3].max())}}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
