"""Small end-to-end run of every kernel for compute-sanitizer (memcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from rl4afcs_b200 import _lib, nl_engine, sp_engine  # noqa: E402

for pol in ("fp64", "mixed", "fp32"):
    eng, x0, w = bench.make_engine(pol, 333, "cuda:0", 0, 80)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(30, log_level=_lib.LOG_FULL, log_agents=100)
    eng.set_hpi("ELIG_A", np.random.default_rng(0).integers(0, 3, 333).astype(np.int32)); eng.set_hpi("ELIG_C", 1)
    eng.run(30, log_level=_lib.LOG_BASIC, log_agents=333, log_every=3)
    eng.run(10)
for pol in ("mixed", "fp64"):
    e = nl_engine.NlEngine(201, policy=pol)
    e.set_reference(nl_engine.theta_reference())
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    w = lambda k: (torch.randn((201, k), generator=g, device="cuda") * 0.1).double()  # noqa: E731
    e.init(w(40), w(10), w(40), w(30))
    e.run(25, torch.randn((25, 201), generator=g, device="cuda"), log_agents=50)
    e.run(25, torch.randn((25, 201), generator=g, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: done")
