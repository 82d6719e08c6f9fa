"""HBM roofline of the step-API kernels (one call each, state streamed from / to HBM).

    python scripts/bench_step_api.py [--agents N] [--policy mixed|fp64|fp32]

achieved GB/s = algorithmic bytes (fields read + written) / CUDA-event time; peak = the measured
STREAM-copy figure of MEASURED_PEAKS.json (6559.7 GB/s) when present, else the 6650 GB/s fallback of the profiling guide."""
import argparse
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl4afcs_b200 import _lib, sp_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=1 << 24)
ap.add_argument("--policy", default="mixed")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
peak = 6650.0
src = "fallback (B200_PROFILING.md)"
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]; src = "MEASURED_PEAKS.json"

n = a.agents
eng = sp_engine.SpEngine(n, policy=a.policy)
L = eng.lib
tn, te = eng.tn, eng.te
sn, se = torch.empty(0, dtype=tn).element_size(), torch.empty(0, dtype=te).element_size()
eng.set_hp("KAPPA", 1140.0); eng.set_hp("REF_AMP", 1.0); eng.set_hp("RLS_GAMMA", 1.0); eng.set_hp("RLS_COV0", 1e6)
eng.set_hpi("FAULT_STEP", -1)
eng.set_reference(np.sin(np.arange(64) * 0.01))
g = torch.Generator(device="cuda"); g.manual_seed(0)
rnd = lambda rows, dt: (torch.rand((rows, n), generator=g, device="cuda", dtype=torch.float32) * 0.02 - 0.01).to(dt)  # noqa: E731
x = rnd(2, te); act = rnd(1, tn); r = torch.empty(n, dtype=te, device="cuda"); e = torch.empty_like(r); rg = torch.empty_like(r)
theta = rnd(6, te); cov = torch.zeros((9, n), dtype=te, device="cuda"); cov[0] = cov[4] = cov[8] = 1e6
dx0 = rnd(2, te); da0 = rnd(1, te); dx1 = rnd(2, te); eps = torch.empty((2, n), dtype=te, device="cuda"); en = torch.empty(n, dtype=te, device="cuda")
z = rnd(1, tn); w1 = rnd(4, tn); w2c = rnd(8, tn); w2a = rnd(4, tn); Ec = torch.zeros((12, n), dtype=te, device="cuda")
Ea = torch.zeros((8, n), dtype=te, device="cuda"); lam = torch.empty((2, n), dtype=tn, device="cuda")
aout = torch.empty(n, dtype=tn, device="cuda"); dadz = torch.empty(n, dtype=tn, device="cuda")
P = ctypes.byref(eng.params)
pid = eng.policy_id
calls = {
    "sp_env_step": (lambda: L.rl4_sp_env_step(pid, P, eng.ref_base.data_ptr(), 3, x.data_ptr(), act.data_ptr(), r.data_ptr(), e.data_ptr(), rg.data_ptr(), n, n, None),
                    (2 * 2 + 3) * se + sn),
    "sp_rls_update": (lambda: L.rl4_sp_rls_update(pid, P, theta.data_ptr(), cov.data_ptr(), dx0.data_ptr(), da0.data_ptr(), dx1.data_ptr(), eps.data_ptr(), en.data_ptr(), n, n, None),
                      (15 * 2 + 5 + 3) * se),
    "sp_critic_forward": (lambda: L.rl4_sp_critic_forward(pid, z.data_ptr(), w1.data_ptr(), w2c.data_ptr(), Ec.data_ptr(), lam.data_ptr(), 0.3, 0, n, n, None),
                          (1 + 4 + 8 + 2) * sn + 12 * 2 * se),
    "sp_actor_forward": (lambda: L.rl4_sp_actor_forward(pid, z.data_ptr(), w1.data_ptr(), w2a.data_ptr(), Ea.data_ptr(), aout.data_ptr(), dadz.data_ptr(), 0.3, 0, n, n, None),
                         (1 + 4 + 4 + 2) * sn + 8 * 2 * se),
}
out = {"policy": a.policy, "agents": n, "peak_gbs": peak, "peak_source": src, "kernels": {}}
for name, (fn, bytes_per_agent) in calls.items():
    for _ in range(3):
        _lib.check(fn(), name)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    gbs = bytes_per_agent * n / (ms * 1e-3) / 1e9
    out["kernels"][name] = {"ms": ms, "bytes_per_agent": bytes_per_agent, "achieved_gbs": gbs, "frac": gbs / peak}
print(json.dumps(out))
