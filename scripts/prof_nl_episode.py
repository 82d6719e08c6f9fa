"""Per-chunk timing of a full nonlinear episode (where in the 90 s flight does the fused kernel spend its time)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rl4afcs_b200 import _lib, nl_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=1 << 18)
ap.add_argument("--steps", type=int, default=9000)
ap.add_argument("--chunk", type=int, default=1000)
ap.add_argument("--integrator", default="ode5")
ap.add_argument("--warmup-steps", type=int, default=None, help="override int(warmup_time/dt) = 400")
a = ap.parse_args()
eng = nl_engine.NlEngine(a.agents, policy="mixed")
eng.params.integrator = _lib.INTEGRATOR[a.integrator]
eng.set_reference(nl_engine.theta_reference())
if a.warmup_steps is not None:
    eng.set_hpi("WARMUP_STEPS", a.warmup_steps)
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = lambda k: (torch.randn((a.agents, k), generator=g, device="cuda").clamp_(-2, 2) * 0.1).double()  # noqa: E731
eng.init(w(40), w(10), w(40), w(30))
k = 0
while k < a.steps:
    c = min(a.chunk, a.steps - k)
    nz = torch.randn((c, a.agents), generator=g, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(c, nz); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    div = int((eng.int_field("DIVERGED_STEP") >= 0).sum())
    eta = float(eng.env_field("ETA_A")[0].median())
    print(f"steps {k:5d}-{k + c:5d}: {ms:8.2f} ms  {a.agents * c / ms * 1e3 / 1e9:6.3f} G/s  eta_a {eta:7.3f}  diverged {div}")
    k += c
chk = eng.env[:, : a.agents].nan_to_num(nan=1.0, posinf=2.0, neginf=3.0).double().sum().item()
print(f"checksum of the final env plane: {chk!r}")
