"""Timing / ncu driver for the fused nonlinear kernel."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl4afcs_b200 import nl_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--policy", default="mixed")
ap.add_argument("--agents", type=int, default=1 << 18)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--warmup", type=int, default=50)
ap.add_argument("--integrator", default="ode5")
a = ap.parse_args()
eng = nl_engine.NlEngine(a.agents, policy=a.policy)
from rl4afcs_b200 import _lib  # noqa: E402
eng.params.integrator = _lib.INTEGRATOR[a.integrator]
eng.set_reference(nl_engine.theta_reference())
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = lambda k: (torch.randn((a.agents, k), generator=g, device="cuda").clamp_(-2, 2) * 0.1).double()  # noqa: E731
eng.init(w(40), w(10), w(40), w(30))
nz = torch.randn((max(a.steps, a.warmup), a.agents), generator=g, device="cuda")
eng.run(a.warmup, nz[: a.warmup])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.run(a.steps, nz[: a.steps]); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"nonlinear {a.policy} {a.integrator}: {a.agents} agents x {a.steps} steps in {ms:.2f} ms -> {a.agents * a.steps / ms * 1e3 / 1e9:.3f} G agent-steps/s")
