#!/usr/bin/env python
"""Benchmark of the fused env + IDHP hot path (BASELINE.json metric: agent-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A bench "step" is one IDHP time step (env step + critic/actor/target update + RLS + adapt) of
EVERY agent of the batch: config[1] of BASELINE.json -- linear short-period IDHP, 2^20 agents
per GPU with randomised initial states and weights, default hyper-parameters of idhp_sp.py.
The K timed steps run inside ONE persistent kernel launch per GPU (state in registers), after a
W-step warm-up launch; the per-GPU state (0.73 GB in fp64) is far larger than L2, so nothing is
cache-resident between launches.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every key.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# Algorithmic FLOPs per agent-step of the fp64/fp32 path with the idhp_sp.py defaults (multistep
# on, no traces); FMA = 2, div = sqrt = 1.  Derivation in DESIGN.md "Roofline accounting".
FLOP_PER_AGENT_STEP = 388 + 13 * 38 + 13 + 2          # body + 13 tanh (37 flop + 1 div) + 13 div + 2 sqrt
# FP64 warp-instructions per agent-step of the fp64 kernel, measured with ncu (profiles/prof_sp_fp64_r01c_raw.csv):
# sm__pipe_fp64_cycles_active 69.5 % x 2450 cycles per warp-step / 2 issue cycles per FP64 instruction
FP64_INSTR_PER_AGENT_STEP = 850
# DRAM traffic of the fused kernel per agent and launch (dram__bytes_read + dram__bytes_write of the same capture,
# 300.7 MB for 2^18 agents): state planes in once, out once (the trace planes are write-only when traces are off)
DRAM_BYTES_PER_AGENT_LAUNCH_FP64 = 1147


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents", type=int, default=1 << 20, help="agents per GPU")
    ap.add_argument("--policy", default="fp64", choices=["fp64", "fp32", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nonlinear", action="store_true")
    return ap.parse_args()


def reference_table(n: int) -> np.ndarray:
    """sin(2 pi t/10) on the idhp_sp.py grid (spacing 60/2999 s, Q10), extended periodically."""
    t = np.arange(n) * (60.0 / 2999.0)
    return np.sin(2 * np.pi * t / 10.0)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def cpu_baseline_c_port(n_steps: int, budget_s: float = 15.0) -> dict:
    """The C oracle (port of the reference path) on all host cores; bounded sample."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import sp_c

    sp_c.build()
    cores = os.cpu_count() or 1
    ic = sp_c.default_idhp_config()
    n_steps = min(n_steps, 3000)
    base = reference_table(n_steps)
    amp = float(np.deg2rad(5))
    per_core = 256
    # calibrate on one core, then size the sample for ~budget_s
    def work(seed, n):
        rng = np.random.default_rng(seed)
        x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
        w = sp_c.init_weights(n, seed)
        cfg = sp_c.make_cfg(ic, ref_amp=amp)
        st = sp_c.init_states("fp64", cfg, x0, w)
        sp_c.run("fp64", cfg, base, st, 0, n_steps, tanh="libm")
        return n
    t0 = time.perf_counter(); work(0, 64); t1 = time.perf_counter()
    rate1 = 64 * n_steps / (t1 - t0)
    per_core = int(max(64, min(1 << 16, rate1 * budget_s / n_steps)))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:          # ctypes releases the GIL: real parallelism
        done = sum(ex.map(lambda s: work(s, per_core), range(cores)))
    t1 = time.perf_counter()
    return {"value": done * n_steps / (t1 - t0), "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": f"C oracle (oracle/sp_oracle.c, fp64, libm tanh), {done} agents x {n_steps} steps, "
                      f"{cores} threads; the TensorFlow reference itself cannot run (no TF in the image)"}


def _numpy_agent(args):
    seed, n_steps = args
    from oracle import sp_c, sp_numpy

    ic = sp_c.default_idhp_config()
    base = reference_table(n_steps)
    ref = float(np.deg2rad(5)) * base
    rng = np.random.default_rng(seed)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(2, 1)))
    env = sp_numpy.ShortPeriodPlant({"x0": x0, "dt": 0.02, "t_end": n_steps * 0.02, "fault_time": 20,
                                     "fault_scenario": None, "reference": {"signal": [ref]}})
    w = sp_c.init_weights(1, seed)
    loop = sp_numpy.IDHPspLoop(env, ic, {k: v[0] for k, v in w.items()})
    loop.train(n_steps)
    return n_steps


def run_reference(args) -> dict:
    """--impl reference: the reference-shaped CPU loop (numpy restatement of IDHPsp.train, one agent
    per task on a process pool like functions.py:131-139).  The TensorFlow original cannot be
    imported in this image; env + RLS of this loop are checked bit-for-bit against the verbatim
    reference classes in the build container."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    cores = os.cpu_count() or 1
    k = min(args.steps, 3000)
    w = min(args.warmup, 50)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_numpy_agent, [(i, max(w, 3)) for i in range(cores)])          # warm-up
        t0 = time.perf_counter()
        done = sum(pool.map(_numpy_agent, [(100 + i, k) for i in range(2 * cores)]))
        t1 = time.perf_counter()
    val = done / (t1 - t0)
    sample = (f"numpy restatement of IDHPsp.train (oracle/sp_numpy.py), {2 * cores} agents x {k} steps, "
              f"one agent per task on a {cores}-process pool")
    return {
        "impl": "reference", "metric": "fused env+IDHP agent-steps/s", "value": val, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (t1 - t0) / k,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "linear short-period IDHP (idhp_sp.py hyper-parameters), CPU sample", "agents": 2 * cores,
                   "episode_steps": k},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ------------------------------------------------------------------------------------------
def make_engine(policy, n, device, seed, n_table):
    import torch

    from rl4afcs_b200 import sp_engine

    eng = sp_engine.SpEngine(n, policy=policy, device=device)
    ic = default_idhp_config()
    sp_engine.apply_idhp_config(eng, ic, dt=0.02)
    eng.set_hp("REF_AMP", float(np.deg2rad(5)))
    eng.set_hpi("FAULT_STEP", -1)
    eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(reference_table(n_table))
    g = torch.Generator(device=device)
    g.manual_seed(1234 + seed)
    x0 = (torch.rand((n, 2), generator=g, device=device, dtype=torch.float64) * 4.0 - 2.0) * (np.pi / 180.0)
    w = sp_engine.truncated_normal_weights(n, 99 + seed, 0.1, device)
    return eng, x0, w


def default_idhp_config() -> dict:
    """idhp_sp.py:150-173."""
    return {
        "gamma": 0.6, "multistep": 2, "gamma_rls": 1.0, "lambda_h": 0.576, "lambda_l": 0.296, "kappa": 1140,
        "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 3.0, "error_thresh": 1, "tau": 0.01, "in_dims": 1,
        "actor_config": {"layers": {4: "tanh", 1: "tanh"}, "eta_h": 3.55, "eta_l": 0.054, "elig": None},
        "critic_config": {"layers": {4: "tanh", 2: "linear"}, "eta_h": 0.338, "eta_l": 0.00, "elig": None},
        "rls_config": {"state_dim": 2, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6},
    }


def timed_run(eng, x0, w, warmup, steps, dist, world):
    """init + W warm-up steps (untimed), then exactly K steps timed with CUDA events on the launch
    stream, barrier + synchronize on both sides, max over ranks."""
    import torch

    from rl4afcs_b200 import _lib

    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.load().rl4_launch_count()
    e0.record()
    eng.run(steps)
    e1.record()
    torch.cuda.synchronize()
    l1 = _lib.load().rl4_launch_count()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=eng.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, int(l1 - l0)


def e2e_run(policy, n, steps, device_index, seed):
    """The host-buffer entry point a reference user would call (IDHPsp(...).train() for a batch):
    pinned host inputs -> H2D -> init + fused run -> D2H of the full final state."""
    import torch

    from rl4afcs_b200 import _lib
    from rl4afcs_b200._lib import SPE, SPI, SPN
    from rl4afcs_b200 import sp_engine

    L = _lib.load()
    eng = sp_engine.SpEngine(1, policy=policy, device=f"cuda:{device_index}")   # only to build the params struct
    sp_engine.apply_idhp_config(eng, default_idhp_config(), dt=0.02)
    eng.set_hp("REF_AMP", float(np.deg2rad(5)))
    eng.set_hpi("FAULT_STEP", -1)
    eng.set_hpi("FAULT_KIND", 0)
    tn, te = sp_engine.policy_dtypes(policy)
    g = torch.Generator(); g.manual_seed(77 + seed)
    pin = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype).pin_memory()  # noqa: E731
    x0 = pin(2, n); x0.copy_((torch.rand((2, n), generator=g, dtype=torch.float64) * 4 - 2) * (np.pi / 180))
    ws = {}
    for nm, wd in (("w1a", 4), ("w2a", 4), ("w1c", 4), ("w2c", 8)):
        ws[nm] = pin(wd, n)
        ws[nm].copy_((torch.randn((wd, n), generator=g, dtype=torch.float32).clamp_(-2, 2) * 0.1).double())
    ref = pin(steps); ref.copy_(torch.from_numpy(reference_table(steps)))
    out_env = pin(SPE["COUNT"], n, dtype=te); out_net = pin(SPN["COUNT"], n, dtype=tn)
    out_ints = pin(SPI["COUNT"], n, dtype=torch.int32)
    ctx = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(device_index, _lib.POLICY[policy], n, steps, ctypes.byref(ctx)), "rl4_ctx_create")
    io = _lib.SpHostIO(x0.data_ptr(), ws["w1a"].data_ptr(), ws["w2a"].data_ptr(), ws["w1c"].data_ptr(),
                       ws["w2c"].data_ptr(), ref.data_ptr(), out_env.data_ptr(), out_net.data_ptr(), out_ints.data_ptr())
    try:
        _lib.check(L.rl4_sp_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, min(steps, 30), 0), "warm-up")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.rl4_sp_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps, 0), "rl4_sp_episode_host")
        t1 = time.perf_counter()
    finally:
        L.rl4_ctx_destroy(ctx)
    h2d = 22 * n * 8 + steps * 8
    d2h = out_env.numel() * out_env.element_size() + out_net.numel() * out_net.element_size() + out_ints.numel() * 4
    div = int((out_ints[SPI["DIVERGED_STEP"]] >= 0).sum())
    return (t1 - t0), h2d, d2h, div


def nonlinear_workload(device, n=1 << 18, steps=300, warmup=700) -> dict:
    """BASELINE.json configs[2]: nonlinear aircraft IDHP attitude tracking, 256K agents, dt = 0.01, reported next to
    the headline (never mixed into it).  The plant is the documented surrogate (reference plant: source-less binary).
    The timed window starts after 700 steps: past the 4 s warm-up of the learning rates and past the wave of early
    divergences, i.e. the regime the remaining 92 % of the 9000-step episode runs in (the first 400 steps are ~12 % faster,
    scripts/prof_nl_episode.py)."""
    import torch

    from rl4afcs_b200 import _lib, nl_engine

    out = {}
    for policy, integ in (("mixed", "ode5"), ("mixed", "rk4"), ("fp64", "ode5")):
        eng = nl_engine.NlEngine(n, policy=policy, device=device)
        eng.params.integrator = _lib.INTEGRATOR[integ]
        eng.set_reference(nl_engine.theta_reference())
        g = torch.Generator(device=device); g.manual_seed(5)
        w = lambda k: (torch.randn((n, k), generator=g, device=device).clamp_(-2, 2) * 0.1).double()  # noqa: E731
        eng.init(w(40), w(10), w(40), w(30))
        nz = torch.randn((max(steps, warmup), n), generator=g, device=device)
        eng.run(warmup, nz[:warmup])
        torch.cuda.synchronize()
        del nz
        nz = torch.randn((steps, n), generator=g, device=device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run(steps, nz[:steps]); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[f"{policy}_{integ}"] = {"value": n * steps / (ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms / steps,
                                    "diverged": int(eng.stats()["diverged"].sum())}
        del eng, nz
        torch.cuda.empty_cache()
    return {"workload": "nonlinear aircraft IDHP attitude tracking (BASELINE.json configs[2]), surrogate 6-DOF plant",
            "agents": n, "steps": steps, "warmup": warmup, "results": out}


def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist

    from rl4afcs_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout to the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(device))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    L = _lib.load()
    n, K, W = args.agents, args.steps, max(args.warmup, 0)

    eng, x0, w = make_engine(args.policy, n, device, rank, W + K)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = timed_run(eng, x0, w, W, K, dist, world)
    # episode statistics: the only cross-GPU exchange of this path (one all-gather at episode end)
    from rl4afcs_b200 import dist as rdist
    stats = rdist.gather_episode_summary(eng, world)
    clocks = sampler.stop() if rank == 0 else {}
    value = world * n * K / (ms * 1e-3)

    # e2e through the host-buffer C-ABI call (per rank, max over ranks)
    e2e = None
    if not args.no_e2e:
        if world > 1:
            dist.barrier()
        t, h2d, d2h, _ = e2e_run(args.policy, n, K, local, rank)
        if world > 1:
            tt = torch.tensor([t], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        e2e = {"value": world * n * K / t, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d / K,
               "d2h_bytes_per_step": d2h / K, "seconds": t,
               "note": "rl4_sp_episode_host: pinned host x0/weights/ref -> GPU, init + K fused steps, full final state -> host"}

    out = None
    if rank == 0:
        is_double = args.policy != "fp32"
        peak = ctypes.c_double(0.0)
        _lib.check(L.rl4_peak_fma(1 if is_double else 0, ctypes.byref(peak), None), "rl4_peak_fma")
        achieved = FLOP_PER_AGENT_STEP * n * K / (ms * 1e-3)          # rank 0's kernel; per GPU
        roof = {"bound": "fp64" if is_double else "fp32", "achieved": achieved / 1e12, "peak": peak.value / 1e12,
                "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": None,
                "kernel": "sp_run_kernel", "flop_per_agent_step": FLOP_PER_AGENT_STEP,
                "peak_source": "rl4_peak_fma measured live on this GPU (MEASURED_PEAKS.json has no FP64/FP32 vector peak)",
                "hbm_bytes_per_agent_step": (eng.env.element_size() * 45 + eng.net.element_size() * 40 + 16) * 2 / K}
        if args.policy == "fp64":
            # pipe view: the parity contract forces separately rounded mul/add (1 flop per FP64 issue slot), so the
            # FLOP fraction understates how busy the binding pipe is; instruction count from ncu (profiles/README.md)
            roof["fp64_instr_per_agent_step"] = FP64_INSTR_PER_AGENT_STEP
            roof["pipe_frac_of_measured_dfma_rate"] = FP64_INSTR_PER_AGENT_STEP * n * K / (ms * 1e-3) / (peak.value / 2.0)
            roof["ncu_fp64_pipe_active_pct"] = 69.5
            # 897 algorithmic FLOP are issued as 850 FP64 instructions (numpy's separately rounded mul / add cannot be
            # fused), so even a 100 % busy pipe reaches only 897 / (2 * 850) of the DFMA FLOP peak
            roof["flop_frac_ceiling_under_parity_contract"] = FLOP_PER_AGENT_STEP / (2.0 * FP64_INSTR_PER_AGENT_STEP)
            roof["traffic"] = DRAM_BYTES_PER_AGENT_LAUNCH_FP64 * n
            roof["traffic_note"] = "HBM bytes per launch from ncu (scaled per agent); the kernel is FP64-pipe bound, not HBM bound"
        variants = {}
        if not args.no_variants and world == 1:
            for pol in ("mixed", "fp32", "fp64"):
                if pol == args.policy:
                    continue
                e2, x2, w2 = make_engine(pol, n, device, rank, W + K)
                ms2, _ = timed_run(e2, x2, w2, W, K, dist, 1)
                variants[pol] = {"value": n * K / (ms2 * 1e-3), "ms_per_step": ms2 / K}
                del e2, x2, w2
                torch.cuda.empty_cache()
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline_c_port(K)
        nonlinear = None
        if not args.no_nonlinear and world == 1:
            nonlinear = nonlinear_workload(device)
        out = {
            "metric": "fused env+IDHP agent-steps/s", "value": value, "unit": "agent-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp64": "f64", "fp32": "f32", "mixed": "f32 nets + f64 env/RLS"}[args.policy],
            "data": "synthetic",
            "config": {"workload": "linear short-period IDHP, independent agents with randomised ICs/weights "
                                   "(BASELINE.json configs[1]; idhp_sp.py hyper-parameters)",
                       "agents_per_gpu": n, "policy": args.policy, "tanh": "t13",
                       "l2": "state planes (%.0f MB/GPU) exceed L2; one persistent launch for the K steps"
                             % ((eng.env.numel() * eng.env.element_size() + eng.net.numel() * eng.net.element_size()) / 1e6)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "variants": variants, "nonlinear": nonlinear, "stats": stats, "impl": "ours",
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        out = run_reference(args)
    else:
        out = run_ours(args)
    if out:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
