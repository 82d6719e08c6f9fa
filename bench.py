#!/usr/bin/env python
"""Benchmark of the fused env + IDHP hot path (BASELINE.json metric: agent-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config sp|sweep|faults]

A bench "step" is one IDHP time step (env step + critic/actor/target update + RLS + adapt) of EVERY agent of the batch.

  --config sp      (default, the headline) BASELINE.json configs[1]: linear short-period IDHP, 2^20 agents per GPU with
                   randomised initial states and weights, default hyper-parameters of idhp_sp.py, fp64.
  --config sweep   BASELINE.json configs[3]: hyper-parameter sweep grid (actor / critic learning rate, RLS forgetting factor,
                   reference amplitude) -- 4M agents over 8 GPUs = 2^19 agents per GPU, mixed policy.
  --config faults  BASELINE.json configs[4]: Monte-Carlo fault study on the nonlinear plant -- 8M agents over 8 GPUs = 2^20
                   agents per GPU, per-agent fault family / time / severity, per-agent statistics gathered with ONE NCCL
                   all-gather (timed separately from the step loop).

The K timed steps run inside persistent kernel launches (state in registers / shared memory) after a W-step warm-up
launch; the per-GPU state (0.73 GB in fp64) is far larger than L2, so nothing is cache-resident between launches.
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

# Algorithmic work per short-period agent-step, SURVEY.md section 8(d) (add and mul counted separately, FMA = 2, multistep
# on, traces off): env 22 + 3 forwards 65 + traces 28 + da/dz, M 18 + TD 24 + critic VJP/SGD 44 + Polyak 36 + actor 30 +
# RLS 71 + adapt 6 = 344 FLOP, with 13 tanh, 12 divisions and 1 square root reported separately.  This is the numerator of
# roofline.achieved.  The kernel's own software tanh / IEEE division expand those primitives into FP64 instructions; that
# expanded count is reported beside it under roofline.expanded, never as the headline.
SURVEY_FLOP_PER_AGENT_STEP = 344
SURVEY_SPECIAL_PER_AGENT_STEP = {"tanh": 13, "div": 12, "sqrt": 1}
# nonlinear agent part (no plant), SURVEY.md section 8(d)
SURVEY_NL_FLOP_PER_AGENT_STEP = 1600
# ncu-derived constants are NOT hard-coded here: they live in profiles/roofline_counters.json next to the name of the
# capture they were read from, and the bench line carries them under the key "from_profile".
COUNTERS_FILE = os.path.join(ROOT, "profiles", "roofline_counters.json")
METRIC = "fused env+IDHP agent-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=300)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="sp", choices=["sp", "sweep", "faults"])
    ap.add_argument("--agents", type=int, default=None, help="agents per GPU (default: the BASELINE.json size of the config)")
    ap.add_argument("--policy", default=None, choices=["fp64", "fp32", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nonlinear", action="store_true")
    return ap.parse_args()


def reference_table(n: int) -> np.ndarray:
    """sin(2 pi t/10) on the idhp_sp.py grid (spacing 60/2999 s, Q10), extended periodically."""
    t = np.arange(n) * (60.0 / 2999.0)
    return np.sin(2 * np.pi * t / 10.0)


def load_counters() -> dict:
    try:
        with open(COUNTERS_FILE) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return {}


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
# CPU arms (the only places bench.py executes anything under oracle/)
# ------------------------------------------------------------------------------------------
def cpu_baseline_c_port(n_steps: int, budget_s: float = 12.0) -> dict:
    """The C oracle (port of the reference path) on all host cores; bounded sample."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import sp_c

    sp_c.build()
    cores = os.cpu_count() or 1
    ic = sp_c.default_idhp_config()
    n_steps = min(n_steps, 3000)
    base = reference_table(n_steps)
    amp = float(np.deg2rad(5))

    def prepare(seed, n):
        rng = np.random.default_rng(seed)
        x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
        w = sp_c.init_weights(n, seed)
        cfg = sp_c.make_cfg(ic, ref_amp=amp)
        return cfg, sp_c.init_states("fp64", cfg, x0, w)

    def work(job):
        cfg, st = job
        sp_c.run("fp64", cfg, base, st, 0, n_steps, tanh="libm")
        return st.shape[0]
    job = prepare(0, 64)
    t0 = time.perf_counter(); work(job); t1 = time.perf_counter()
    rate1 = 64 * n_steps / (t1 - t0)
    per_core = int(max(64, min(1 << 18, rate1 * budget_s / n_steps)))
    jobs = [prepare(s, per_core) for s in range(cores)]      # states are built outside the timed region
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:                    # ctypes releases the GIL: real parallelism
        done = sum(ex.map(work, jobs))
    t1 = time.perf_counter()
    return {"value": done * n_steps / (t1 - t0), "unit": "agent-steps/s", "cores": cores, "kind": "port", "seconds": t1 - t0,
            "sample": f"C oracle (oracle/sp_oracle.c, fp64, libm tanh), {done} agents x {n_steps} steps, "
                      f"{cores} threads; the TensorFlow reference itself cannot run (no TF in the image)"}


def _build_numpy_loops(seed0: int, count: int, n_steps: int):
    from oracle import sp_c, sp_numpy

    ic = sp_c.default_idhp_config()
    ref = float(np.deg2rad(5)) * reference_table(n_steps)
    loops = []
    for j in range(count):
        seed = seed0 + j
        rng = np.random.default_rng(seed)
        x0 = np.deg2rad(rng.uniform(-2, 2, size=(2, 1)))
        env = sp_numpy.ShortPeriodPlant({"x0": x0, "dt": 0.02, "t_end": n_steps * 0.02, "fault_time": 20,
                                         "fault_scenario": None, "reference": {"signal": [ref]}})
        w = sp_c.init_weights(1, seed)
        loops.append(sp_numpy.IDHPspLoop(env, ic, {k: v[0] for k, v in w.items()}))
    return loops


def _ref_worker(idx, per_core, k, w, ready, done):
    """One process of the reference arm: builds its agents (untimed), warms up, then -- between the two barriers the parent
    times -- runs `train(k)` of every agent it owns."""
    warm = _build_numpy_loops(10_000_000 + idx * 16, 2, max(w, 3))
    loops = _build_numpy_loops(100 + idx * per_core, per_core, k)
    for lp in warm:
        lp.train(max(w, 3))
    ready.wait(timeout=900)
    for lp in loops:
        lp.train(k)
    done.wait(timeout=900)


def run_reference(args) -> dict:
    """--impl reference: the reference-shaped CPU loop (numpy restatement of IDHPsp.train with one object graph per agent,
    one agent per task on a process pool like functions.py:131-139).  The TensorFlow original cannot be imported in this
    image; this loop produced the same bits as the verbatim agent on the TensorFlow stand-in.  Agents are
    constructed OUTSIDE the timed region and the sample is sized for >= 3 s of work.  The C oracle (a compiled port of the
    same path, all host threads) is timed in the same run and printed beside it, so the GPU/CPU ratio has a second,
    non-Python denominator."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    cores = os.cpu_count() or 1
    k = max(1, min(args.steps, 3000))
    w = min(max(args.warmup, 3), 50)
    # calibrate one agent on this core, then size the sample: >= 3 s of timed work, at most ~40 s
    probe = _build_numpy_loops(1, 1, 60)[0]
    t0 = time.perf_counter(); probe.train(60); rate1 = 60 / (time.perf_counter() - t0)
    per_core = int(min(max(1, np.ceil(4.0 * rate1 / k)), 4000))
    ctx = mp.get_context("fork")
    ready, done = ctx.Barrier(cores + 1), ctx.Barrier(cores + 1)
    procs = [ctx.Process(target=_ref_worker, args=(i, per_core, k, w, ready, done)) for i in range(cores)]
    for p in procs:
        p.start()
    ready.wait(timeout=900)          # a crashed worker breaks the barrier instead of hanging the run
    t0 = time.perf_counter()
    done.wait(timeout=900)
    t1 = time.perf_counter()
    for p in procs:
        p.join()
    agents = per_core * cores
    val = agents * k / (t1 - t0)
    sample = (f"numpy restatement of IDHPsp.train (oracle/sp_numpy.py), {agents} agents x {k} steps, {per_core} agents per "
              f"process on {cores} processes, agents constructed outside the timed region ({t1 - t0:.2f} s timed)")
    c_oracle = cpu_baseline_c_port(k, budget_s=6.0)
    return {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (t1 - t0) / k,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "linear short-period IDHP (BASELINE.json configs[1] hyper-parameters = idhp_sp.py), CPU sample of the "
                               "same workload", "agents": agents, "episode_steps": k, "seconds_timed": t1 - t0},
        "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "cpu_baseline_c_oracle": c_oracle,
        "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ------------------------------------------------------------------------------------------
def default_idhp_config() -> dict:
    """idhp_sp.py:150-173."""
    return {
        "gamma": 0.6, "multistep": 2, "gamma_rls": 1.0, "lambda_h": 0.576, "lambda_l": 0.296, "kappa": 1140,
        "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 3.0, "error_thresh": 1, "tau": 0.01, "in_dims": 1,
        "actor_config": {"layers": {4: "tanh", 1: "tanh"}, "eta_h": 3.55, "eta_l": 0.054, "elig": None},
        "critic_config": {"layers": {4: "tanh", 2: "linear"}, "eta_h": 0.338, "eta_l": 0.00, "elig": None},
        "rls_config": {"state_dim": 2, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6},
    }


def make_engine(policy, n, device, seed, n_table):
    import torch

    from rl4afcs_b200 import sp_engine

    eng = sp_engine.SpEngine(n, policy=policy, device=device)
    sp_engine.apply_idhp_config(eng, default_idhp_config(), dt=0.02)
    eng.set_hp("REF_AMP", float(np.deg2rad(5)))
    eng.set_hpi("FAULT_STEP", -1)
    eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(reference_table(n_table))
    g = torch.Generator(device=device)
    g.manual_seed(1234 + seed)
    x0 = (torch.rand((n, 2), generator=g, device=device, dtype=torch.float64) * 4.0 - 2.0) * (np.pi / 180.0)
    w = sp_engine.truncated_normal_weights(n, 99 + seed, 0.1, device)
    return eng, x0, w


def make_sweep_engine(n, device, rank, world, n_table):
    """BASELINE.json configs[3]: grid over eta_a_h in [2.5, 4.7] x eta_c_h in [0.45, 0.55] x
    rls_gamma in [0.99, 1.0] x reference amplitude in [1, 10] deg (the reference has no excitation signal; the amplitude of
    the tracked sine stands in), flattened row-major onto the global agent index, contiguous shard per rank."""
    import torch

    from rl4afcs_b200 import sp_engine

    total = n * world
    idx = np.arange(rank * n, (rank + 1) * n)
    dims = [64, 64, 32, 32]
    while np.prod(dims) > total:
        dims[int(np.argmax(dims))] //= 2
    i3 = idx % dims[3]; i2 = (idx // dims[3]) % dims[2]
    i1 = (idx // (dims[3] * dims[2])) % dims[1]; i0 = (idx // (dims[3] * dims[2] * dims[1])) % dims[0]
    lin = lambda i, d, lo, hi: lo + (hi - lo) * (i / max(d - 1, 1))  # noqa: E731
    eng = sp_engine.SpEngine(n, policy="mixed", device=device)
    sp_engine.apply_idhp_config(eng, default_idhp_config(), dt=0.02)
    eng.set_hp("ETA_A_H", lin(i0, dims[0], 2.5, 4.7)); eng.set_hp("ETA_C_H", lin(i1, dims[1], 0.45, 0.55))
    eng.set_hp("RLS_GAMMA", lin(i2, dims[2], 0.99, 1.0)); eng.set_hp("REF_AMP", np.deg2rad(lin(i3, dims[3], 1.0, 10.0)))
    eng.set_hpi("FAULT_STEP", -1); eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(reference_table(n_table))
    w = sp_engine.truncated_normal_weights(n, 7, 0.1, device)     # the same draw on every rank: grid points differ by hyper-parameters only
    x0 = torch.zeros((n, 2), dtype=torch.float64, device=device)
    return eng, x0, w, dims


NL_FAULT_NAMES = ["damp_elevator", "damp_aileron", "damp_rudder", "damp_all", "shift_cg", "slow_all", "saturate_elevator",
                  "saturate_aileron", "saturate_rudder"]              # envs/nonlinear/env.py:134-158


def make_faults_engine(n, device, rank, fault_window_steps=None):
    """BASELINE.json configs[4]: per-agent fault family sampled uniformly from the nine names of
    envs/nonlinear/env.py:134-158, fault time U(30, 70) s (compressed into the timed window when `fault_window_steps` is
    given, so that a short bench run exercises the faulted code paths), damping factor U(0.2, 0.5), c.g. shift U(-0.5, 0)."""
    import torch

    from rl4afcs_b200 import nl_engine

    rng = np.random.default_rng(1000 + rank)
    pick = rng.integers(0, len(NL_FAULT_NAMES), n)
    table = np.asarray([nl_engine.split_fault(nm) for nm in NL_FAULT_NAMES], dtype=np.int32)
    ds = table[pick]
    eng = nl_engine.NlEngine(n, policy="mixed", device=device)
    eng.set_hpi("FAULT_DAMP", ds[:, 0]); eng.set_hpi("FAULT_SAT", ds[:, 1])
    u = rng.uniform(0.0, 1.0, n)
    if fault_window_steps is None:
        fstep = ((30.0 + 40.0 * u) / 0.01).astype(np.int32)
    else:
        lo, hi = fault_window_steps
        fstep = (lo + (hi - lo) * u).astype(np.int32)
    eng.set_hpi("FAULT_STEP", fstep)
    eng.set_hp("DAMP_FACTOR", rng.uniform(0.2, 0.5, n)); eng.set_hp("CG_SHIFT", rng.uniform(-0.5, 0.0, n))
    eng.set_reference(nl_engine.theta_reference())
    g = torch.Generator(device=device); g.manual_seed(11 + rank)
    wd = lambda k: (torch.randn((n, k), generator=g, device=device).clamp_(-2, 2) * 0.1).double()  # noqa: E731
    eng.init(wd(40), wd(10), wd(40), wd(30))
    return eng, g, pick


def timed_region(fn, device, dist, world):
    """barrier + synchronize, CUDA events on the launch stream around fn(), synchronize + barrier, max over ranks."""
    import torch

    from rl4afcs_b200 import _lib

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.load().rl4_launch_count()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    l1 = _lib.load().rl4_launch_count()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, int(l1 - l0)


def timed_run(eng, x0, w, warmup, steps, dist, world):
    """init + W warm-up steps (untimed), then exactly K steps timed."""
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(warmup)
    return timed_region(lambda: eng.run(steps), eng.device, dist, world)


def bind_to_gpu_numa_node(device_index: int):
    """Best effort: run this process (and therefore first-touch its pinned staging buffers) on the CPUs of the GPU's NUMA
    node, so that eight ranks do not all stage through one memory controller.  Returns the node, or None when the
    topology is not visible / the cpuset forbids it."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def e2e_sp(policy, n, steps, device_index, seed, out_groups=("STATS", "WEIGHTS", "RLS")):
    """The host-buffer entry point a reference user would call (IDHPsp(...).train() for a batch): pinned host inputs ->
    H2D -> init + fused run -> D2H of what train() leaves behind for its caller (statistics, final weights, RLS model)."""
    import torch

    from rl4afcs_b200 import _lib, sp_engine
    from rl4afcs_b200._lib import OUT, SPE, SPI, SPN

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
    mask = sum(OUT[k] for k in out_groups)
    ctx = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(device_index, _lib.POLICY[policy], n, steps, ctypes.byref(ctx)), "rl4_ctx_create")
    io = _lib.SpHostIO(x0.data_ptr(), ws["w1a"].data_ptr(), ws["w2a"].data_ptr(), ws["w1c"].data_ptr(),
                       ws["w2c"].data_ptr(), ref.data_ptr(), out_env.data_ptr(), out_net.data_ptr(), out_ints.data_ptr(), mask, 0)
    try:
        _lib.check(L.rl4_sp_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, min(steps, 30), 0), "warm-up")
        torch.cuda.synchronize()
        l0 = L.rl4_launch_count()
        t0 = time.perf_counter()
        _lib.check(L.rl4_sp_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps, 0), "rl4_sp_episode_host")
        t1 = time.perf_counter()
        launches = int(L.rl4_launch_count() - l0)
    finally:
        L.rl4_ctx_destroy(ctx)
    env_rows = (4 if "STATE" in out_groups else 0) + (1 if "STATE" in out_groups else 0) + (18 if "RLS" in out_groups else 0) \
        + (2 if "STATS" in out_groups else 0) + (20 if "TRACES" in out_groups else 0)
    net_rows = (8 if "STATE" in out_groups else 0) + (20 if "WEIGHTS" in out_groups else 0) + (12 if "TARGET" in out_groups else 0)
    int_rows = 4 if ("STATS" in out_groups or "STATE" in out_groups) else 0
    h2d = 22 * n * 8 + steps * 8
    d2h = n * (env_rows * out_env.element_size() + net_rows * out_net.element_size() + int_rows * 4)
    div = int((out_ints[SPI["DIVERGED_STEP"]] >= 0).sum())
    return {"seconds": t1 - t0, "h2d": h2d, "d2h": d2h, "diverged": div, "launches": launches,
            "returned": "+".join(g_.lower() for g_ in out_groups)}


def e2e_nl(n, steps, device_index, seed, integ="ode5"):
    """IDHPnonlin(...).train() for a batch through rl4_nl_episode_host: pinned host weights / reference -> GPU, reset +
    prologue + the fused launches with the N(0,1) stream drawn on the device, statistics + weights + RLS model -> host."""
    import torch

    from rl4afcs_b200 import _lib, nl_engine
    from rl4afcs_b200._lib import NLE, NLI, NLN, OUT

    L = _lib.load()
    eng = nl_engine.NlEngine(1, policy="mixed", device=f"cuda:{device_index}")    # only to build the params struct
    eng.params.integrator = _lib.INTEGRATOR[integ]
    g = torch.Generator(); g.manual_seed(177 + seed)
    pin = lambda *shape, dtype=torch.float64: torch.empty(shape, dtype=dtype).pin_memory()  # noqa: E731
    ws = []
    for wd in (40, 10, 40, 30):
        t = pin(wd, n); t.copy_((torch.randn((wd, n), generator=g, dtype=torch.float32).clamp_(-2, 2) * 0.1).double())
        ws.append(t)
    th = nl_engine.theta_reference()
    reps = (steps + th.shape[0] - 1) // th.shape[0]
    ref = pin(steps); ref.copy_(torch.from_numpy(np.tile(th, reps)[:steps].copy()))
    out_env = pin(NLE["COUNT"], n); out_net = pin(NLN["COUNT"], n, dtype=torch.float32); out_ints = pin(NLI["COUNT"], n, dtype=torch.int32)
    mask = OUT["STATS"] | OUT["WEIGHTS"] | OUT["RLS"]
    ctx = ctypes.c_void_p()
    _lib.check(L.rl4_ctx_create(device_index, _lib.MIXED, n, steps, ctypes.byref(ctx)), "rl4_ctx_create")
    io = _lib.NlHostIO(ws[0].data_ptr(), ws[1].data_ptr(), ws[2].data_ptr(), ws[3].data_ptr(), ref.data_ptr(), None, 4242 + seed, 0,
                       out_env.data_ptr(), out_net.data_ptr(), out_ints.data_ptr(), mask, 0)
    try:
        _lib.check(L.rl4_nl_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, min(steps, 10)), "warm-up")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.rl4_nl_episode_host(ctx, ctypes.byref(eng.params), ctypes.byref(io), n, steps), "rl4_nl_episode_host")
        t1 = time.perf_counter()
    finally:
        L.rl4_ctx_destroy(ctx)
    env_rows = 32 + 7 + 2       # RLS (theta 12, cov 16, eps 3, eps_norm 1) + statistics (rse 2, nz, eta/lambda 4) + rse_flight 2
    h2d = 120 * n * 8 + steps * 8
    d2h = n * (env_rows * 8 + 120 * 4 + 4 * 4)
    return {"seconds": t1 - t0, "h2d": h2d, "d2h": d2h, "diverged": int((out_ints[NLI["DIVERGED_STEP"]] >= 0).sum())}


def nonlinear_workload(device, counters, n=1 << 18, steps=300, warmup=700, e2e=True, sm_mhz=1965.0) -> dict:
    """BASELINE.json configs[2]: nonlinear aircraft IDHP attitude tracking, 256K agents, dt = 0.01, reported next to
    the headline (never mixed into it).  The plant is the documented surrogate (reference plant: source-less binary).
    The timed window starts after 700 steps: past the 4 s warm-up of the learning rates and past the wave of early
    divergences, i.e. the regime the remaining 92 % of the 9000-step episode runs in."""
    import torch

    from rl4afcs_b200 import _lib, nl_engine

    out = {}
    prop = torch.cuda.get_device_properties(device)
    issue_peak = prop.multi_processor_count * 4 * 32 * sm_mhz * 1e6        # thread-instructions per second
    prof = counters.get("nl_run_kernel_mixed_ode5", {})
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
        rate = n * steps / (ms * 1e-3)
        res = {"value": rate, "unit": "agent-steps/s", "ms_per_step": ms / steps, "diverged": int(eng.stats()["diverged"].sum())}
        if policy == "mixed" and integ == "ode5":
            # binding resource of this kernel: instruction issue under latency (it is neither HBM- nor FP-pipe-bound): thread
            # instructions per agent-step (ncu, from_profile) x measured rate / issue-slot peak of the chip
            roof = {"bound": "issue", "unit": "Ginstr/s", "peak": issue_peak / 1e9,
                    "peak_source": "SMs x 4 schedulers x 32 lanes x SM clock (one warp instruction per scheduler per cycle)",
                    "survey_agent_flop_per_agent_step": SURVEY_NL_FLOP_PER_AGENT_STEP,
                    "achieved_survey_tflops": SURVEY_NL_FLOP_PER_AGENT_STEP * rate / 1e12,
                    "hbm_algorithmic_bytes_per_launch": 2 * n * (113 * 8 + 211 * 4 + 16) + steps * n * 4,
                    "from_profile": prof or None}
            if prof.get("instr_per_agent_step"):
                roof["achieved"] = prof["instr_per_agent_step"] * rate / 1e9
                roof["frac"] = prof["instr_per_agent_step"] * rate / issue_peak
            res["roofline"] = roof
        out[f"{policy}_{integ}"] = res
        del eng, nz
        torch.cuda.empty_cache()
    ret = {"workload": "nonlinear aircraft IDHP attitude tracking (BASELINE.json configs[2]), surrogate 6-DOF plant",
           "agents": n, "steps": steps, "warmup": warmup, "results": out}
    ret["dasmat"] = dasmat_workload(device, counters, issue_peak)
    if e2e:
        idx = torch.device(device).index or 0
        r = e2e_nl(n, steps, idx, 0)
        ret["e2e"] = {"value": n * steps / r["seconds"], "unit": "agent-steps/s", "seconds": r["seconds"],
                      "h2d_bytes_per_step": r["h2d"] / steps, "d2h_bytes_per_step": r["d2h"] / steps,
                      "note": "rl4_nl_episode_host (mixed, ode5): pinned host weights / reference -> GPU, reset + K fused steps from the "
                              "start of the episode with device-drawn noise, statistics + weights + RLS -> host"}
    return ret


def dasmat_workload(device, counters, issue_peak, n=151552, steps=12, warmup=4) -> dict:
    """The same task on the reference's OWN aircraft model (plant='dasmat': the `_citation` binary translated at build time,
    csrc/dasmat_plant.cu) -- exact plant, ~270x heavier than the surrogate (113 712 x86 instructions per step).  Reported
    beside the surrogate numbers; n = one resident wave of 1024 aircraft per SM (one 1024-thread CTA each)."""
    import torch

    from rl4afcs_b200 import _lib, nl_engine

    L = _lib.load()
    if not L.rl4_dasmat_available():
        return {"unavailable": "library built without the reference's plant binary"}
    prof = counters.get("dasmat_step_kernel", {})
    out = {"agents": n, "steps": steps, "warmup": warmup,
           "x86_instructions_per_plant_step": 113712, "note": "translated from envs/nonlinear/extended_input/_citation.cp39-win_amd64.pyd"}
    # (1) the plant alone: citation.step for every aircraft
    dz = nl_engine.DasmatPlant(L, torch.device(device), n, n)
    trim = [-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0]
    dz.reset(trim, 1001)
    g = torch.Generator(device=device); g.manual_seed(3)
    u = torch.tensor(trim, dtype=torch.float64, device=device).reshape(11, 1).repeat(1, n)
    u[0:3] += 0.01 * torch.randn((3, n), generator=g, device=device, dtype=torch.float64)
    u = u.contiguous()
    xo = torch.zeros((12, n), dtype=torch.float64, device=device)
    run = lambda k: _lib.check(L.rl4_dasmat_step(dz.image.data_ptr(), dz.state.data_ptr(), n, n, u.data_ptr(), n, k, xo.data_ptr(), n,  # noqa: E731
                                                 None, dz.err.data_ptr(), None), "rl4_dasmat_step")
    run(warmup); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(steps); e1.record(); torch.cuda.synchronize()
    dz.check()
    ms = e0.elapsed_time(e1)
    rate = n * steps / (ms * 1e-3)
    roof = {"bound": "issue", "unit": "Ginstr/s", "peak": issue_peak / 1e9, "from_profile": prof or None}
    if prof.get("instr_per_plant_step"):
        roof["achieved"] = prof["instr_per_plant_step"] * rate / 1e9
        roof["frac"] = prof["instr_per_plant_step"] * rate / issue_peak
    if prof.get("dram_bytes_per_plant_step"):
        # the second limiter: thread-local memory (the model's block signals and its x86 stack frames, ~11 KB touched per aircraft)
        # does not fit L1 / L2 for a resident wave, so its stores stream to HBM
        hbm_peak = 6559.7e9
        try:
            hbm_peak = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"]) * 1e9
        except Exception:  # noqa: BLE001 - driver-written file; the fallback is the value SURVEY.md quotes from it
            pass
        roof["hbm"] = {"achieved_gbps": prof["dram_bytes_per_plant_step"] * rate / 1e9, "peak_gbps": hbm_peak / 1e9,
                       "frac": prof["dram_bytes_per_plant_step"] * rate / hbm_peak}
    out["plant_only"] = {"value": rate, "unit": "plant-steps/s", "ms_per_step": ms / steps, "roofline": roof,
                         "native_binary_one_host_core_steps_per_s": 3.2e4,
                         "native_note": "the reference's .pyd executing natively on one core of the build container (it cannot run on this box)"}
    try:        # CPU yardstick measured HERE: the gcc build of the same translation (test infrastructure, oracle/pe_probe), one core
        from oracle.pe_probe import lifted
        if lifted.available():
            ac = lifted.Aircraft(); ac.initialize()
            ac.run(np.asarray(trim), 200)
            t0 = time.perf_counter(); ac.run(np.asarray(trim), 4000); dt = time.perf_counter() - t0
            out["plant_only"]["cpu_translation_one_core_steps_per_s"] = 4000 / dt
    except Exception as e:  # noqa: BLE001 - a missing checker library must not fail the bench
        out["plant_only"]["cpu_translation_error"] = str(e)[:120]
    del dz
    # (2) fused env + agent on it
    eng = nl_engine.NlEngine(n, policy="mixed", device=device, plant="dasmat")
    eng.set_reference(nl_engine.theta_reference())
    w = lambda k: (torch.randn((n, k), generator=g, device=device).clamp_(-2, 2) * 0.1).double()  # noqa: E731
    eng.init(w(40), w(10), w(40), w(30))
    nz = torch.randn((warmup + steps, n), generator=g, device=device)
    eng.run(warmup, nz[:warmup]); torch.cuda.synchronize()
    e0.record(); eng.run(steps, nz[warmup:]); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out["fused_mixed"] = {"value": n * steps / (ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms / steps,
                          "diverged": int(eng.stats()["diverged"].sum())}
    del eng, nz
    torch.cuda.empty_cache()
    return out


def sp_roofline(policy, n, K, ms, eng, counters, L):
    """Pipe roofline of sp_run_kernel: SURVEY 8(d) numerator (344 FLOP/agent-step) over the measured DFMA / FFMA peak."""
    is_double = policy != "fp32"
    peak = ctypes.c_double(0.0)
    from rl4afcs_b200 import _lib
    _lib.check(L.rl4_peak_fma(1 if is_double else 0, ctypes.byref(peak), None), "rl4_peak_fma")
    peak3 = ctypes.c_double(0.0)
    _lib.check(L.rl4_peak_fma(3 if is_double else 2, ctypes.byref(peak3), None), "rl4_peak_fma")
    rate = n * K / (ms * 1e-3)                                     # rank 0's kernel; per GPU
    achieved = SURVEY_FLOP_PER_AGENT_STEP * rate
    state_bytes = eng.env.element_size() * 45 + eng.net.element_size() * 40 + 16
    roof = {"bound": "fp64" if is_double else "fp32", "kernel": "sp_run_kernel", "achieved": achieved / 1e12, "peak": peak.value / 1e12,
            "unit": "TFLOP/s", "frac": achieved / peak.value, "traffic": None,
            "numerator": f"SURVEY.md 8(d): {SURVEY_FLOP_PER_AGENT_STEP} FLOP per agent-step (add / mul separately, FMA = 2); "
                         "13 tanh + 12 div + 1 sqrt reported separately",
            "flop_per_agent_step": SURVEY_FLOP_PER_AGENT_STEP,
            "special_ops_per_s": {k: v * rate for k, v in SURVEY_SPECIAL_PER_AGENT_STEP.items()},
            "peak_source": "rl4_peak_fma measured live on this GPU (MEASURED_PEAKS.json has no FP64/FP32 vector peak)",
            "fma_rate_distinct_operands_tflops": peak3.value / 1e12,
            "fma_rate_note": "the same micro-benchmark with three distinct register operands per FMA (the agent kernels' operand pattern); "
                             "context only, the fraction above uses the shared-operand peak",
            "hbm_algorithmic_bytes_per_launch": 2 * state_bytes * n,
            "hbm_bytes_per_agent_step": 2 * state_bytes / K}
    prof = counters.get(f"sp_run_kernel_{policy}")
    if prof:
        fp = dict(prof)
        # derived with THIS run's rate: how busy the binding pipe is in issue slots (1 FP64 instruction = 1 slot, DFMA peak = 2 FLOP/slot)
        if prof.get("fp64_instr_per_agent_step"):
            fp["fp64_issue_slot_frac"] = prof["fp64_instr_per_agent_step"] * rate / (peak.value / 2.0)
        if prof.get("expanded_flop_per_agent_step"):
            roof["expanded"] = {"flop_per_agent_step": prof["expanded_flop_per_agent_step"],
                                "achieved": prof["expanded_flop_per_agent_step"] * rate / 1e12,
                                "frac": prof["expanded_flop_per_agent_step"] * rate / peak.value,
                                "note": "the 344 FLOP plus the arithmetic inside the software tanh / IEEE divisions / square roots "
                                        "(the reference gets those from libm / TensorFlow); NOT the headline fraction"}
        if prof.get("dram_bytes_per_agent_launch"):
            roof["traffic"] = prof["dram_bytes_per_agent_launch"] * n
        roof["from_profile"] = fp
    return roof


def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist

    from rl4afcs_b200 import _lib
    from rl4afcs_b200 import dist as rdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
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
    K, W = args.steps, max(args.warmup, 0)
    counters = load_counters()
    sampler = ClockSampler(local)
    out = None

    if args.config == "faults":
        n = args.agents or (1 << 20)
        eng, g, pick = make_faults_engine(n, device, rank, fault_window_steps=(0, W + K))
        noise = lambda c: torch.randn((c, n), generator=g, device=device, dtype=torch.float32)  # noqa: E731
        if W:
            eng.run(W, noise(W))
        nz = noise(K)
        if rank == 0:
            sampler.start()
        ms, launches = timed_region(lambda: eng.run(K, nz), device, dist, world)
        clocks = sampler.stop() if rank == 0 else {}
        # the one collective of this workload: per-agent statistics (4 numbers per agent) all-gathered over NCCL
        st = eng.stats_planes()
        S = _lib.NLS
        per_agent = torch.stack([st[S["RSE_WARMUP"]] + st[S["RSE_FLIGHT"]], st[S["NZ_PEAK"]], st[S["DIVERGED"]],
                                 torch.as_tensor(pick, device=device, dtype=torch.float64)], dim=1).float().contiguous()
        tg, _ = timed_region(lambda: rdist.gather_per_agent(per_agent, world, n_total=n * world), device, dist, world)
        allst = rdist.gather_per_agent(per_agent, world, n_total=n * world)
        if rank == 0:
            alive = allst[:, 2] == 0
            by_fault = {nm: {"agents": int((allst[:, 3] == i).sum()), "diverged": int(((allst[:, 3] == i) & ~alive).sum())}
                        for i, nm in enumerate(NL_FAULT_NAMES)}
            out = {"metric": METRIC, "value": world * n * K / (ms * 1e-3), "unit": "agent-steps/s", "n_gpus": world, "steps": K,
                   "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                   "dtype": "f32 nets + f64 plant/RLS", "data": "synthetic",
                   "config": {"workload": "Monte-Carlo fault study, nonlinear plant with sampled actuator / c.g. faults "
                                          "(BASELINE.json configs[4]: 8M agents over 8 GPUs = 2^20 per GPU), surrogate 6-DOF plant",
                              "agents_per_gpu": n, "policy": "mixed", "integrator": "ode5", "fault_step_window": [0, W + K],
                              "l2": "state planes (%.0f MB/GPU) exceed L2" % ((eng.env.numel() * 8 + eng.net.numel() * 4) / 1e6)},
                   "clocks": clocks, "gpu_launches": launches, "e2e": None,
                   "gather": {"collective": "all_gather_into_tensor (NCCL)", "bytes_total": int(allst.numel() * 4), "ms": tg},
                   "stats": {"agents": int(allst.shape[0]), "diverged": int((~alive).sum()),
                             "mean_cumulative_RSE_theta": float(allst[alive, 0].double().mean()) if bool(alive.any()) else None,
                             "peak_nz": float(allst[alive, 1].max()) if bool(alive.any()) else None, "by_fault": by_fault},
                   "impl": "ours"}
    else:
        sweep = args.config == "sweep"
        policy = args.policy or ("mixed" if sweep else "fp64")
        n = args.agents or ((1 << 19) if sweep else (1 << 20))
        if sweep:
            eng, x0, w, dims = make_sweep_engine(n, device, rank, world, W + K)
        else:
            eng, x0, w = make_engine(policy, n, device, rank, W + K)
        if rank == 0:
            sampler.start()
        ms, launches = timed_run(eng, x0, w, W, K, dist, world)
        # episode statistics: the only cross-GPU exchange of this path (one all-gather at episode end)
        stats = rdist.gather_episode_summary(eng, world)
        clocks = sampler.stop() if rank == 0 else {}
        value = world * n * K / (ms * 1e-3)

        # e2e through the host-buffer C-ABI call (per rank, max over ranks)
        e2e = None
        if not args.no_e2e and not sweep:
            if world > 1:
                dist.barrier()
            r = e2e_sp(policy, n, K, local, rank)
            t = r["seconds"]
            if world > 1:
                tt = torch.tensor([t], device=device, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t = float(tt.item())
            e2e = {"value": world * n * K / t, "unit": "agent-steps/s", "h2d_bytes_per_step": r["h2d"] / K,
                   "d2h_bytes_per_step": r["d2h"] / K, "seconds": t, "returned": r["returned"], "numa_node": numa,
                   "note": "rl4_sp_episode_host: pinned host x0/weights/ref -> GPU, init + K fused steps, statistics + final weights + "
                           "RLS model -> host (what IDHPsp.train() leaves for its caller)"}
        if rank == 0:
            roof = sp_roofline(policy, n, K, ms, eng, counters, L)
            extras = {}
            if e2e is not None and world == 1 and K < 3000:
                # the copy-hidden regime: one whole 3000-step episode through the same host-buffer call
                r = e2e_sp(policy, n, 3000, local, rank)
                extras["e2e_full_episode"] = {"value": n * 3000 / r["seconds"], "unit": "agent-steps/s", "steps": 3000,
                                              "seconds": r["seconds"], "h2d_bytes": r["h2d"], "d2h_bytes": r["d2h"]}
                r = e2e_sp(policy, n, K, local, rank, out_groups=("STATS",))
                extras["e2e_stats_only"] = {"value": n * K / r["seconds"], "unit": "agent-steps/s", "seconds": r["seconds"],
                                            "d2h_bytes_per_step": r["d2h"] / K}
            variants = {}
            if not args.no_variants and world == 1 and not sweep:
                for pol in ("mixed", "fp32", "fp64"):
                    if pol == policy:
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
            if not args.no_nonlinear and world == 1 and not sweep:
                nonlinear = nonlinear_workload(device, counters, sm_mhz=float(clocks.get("sm_max_mhz") or 1965.0))
            workload = ("hyper-parameter sweep grid: actor / critic learning rate x RLS forgetting factor x reference amplitude "
                        "(BASELINE.json configs[3]: 4M agents over 8 GPUs = 2^19 per GPU), linear short-period IDHP" if sweep else
                        "linear short-period IDHP, independent agents with randomised ICs/weights "
                        "(BASELINE.json configs[1]; idhp_sp.py hyper-parameters)")
            cfg = {"workload": workload, "agents_per_gpu": n, "policy": policy, "tanh": "t13",
                   "l2": "state planes (%.0f MB/GPU) exceed L2; one persistent launch for the K steps"
                         % ((eng.env.numel() * eng.env.element_size() + eng.net.numel() * eng.net.element_size()) / 1e6)}
            if sweep:
                cfg["grid"] = dims
            out = {
                "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp64": "f64", "fp32": "f32", "mixed": "f32 nets + f64 env/RLS"}[policy],
                "data": "synthetic", "config": cfg,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
                "variants": variants, "nonlinear": nonlinear, "stats": stats, "impl": "ours", **extras,
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
