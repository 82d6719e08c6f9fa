"""How many nonlinear agents learn to track: IDHPnonlin.train() for a batch (idhp_nonlin.py hyper-parameters, random initial
weights / noise) on the calibrated surrogate plant or on the reference's own aircraft model (--plant dasmat), statistics of
the pitch error over the last 10 s of the flight.

    python scripts/nl_learning_stats.py [--agents 4096] [--steps 3000] [--plant surrogate|dasmat]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl4afcs_b200 import nl_engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=4096)
ap.add_argument("--steps", type=int, default=3000)
ap.add_argument("--plant", default="surrogate")
a = ap.parse_args()
n = a.agents
eng = nl_engine.NlEngine(n, policy="mixed", plant=a.plant)
eng.set_reference(nl_engine.theta_reference())
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = lambda k: (torch.randn((n, k), generator=g, device="cuda").clamp_(-2, 2) * 0.1).double()  # noqa: E731
eng.init(w(40), w(10), w(40), w(30))
last = min(1000, a.steps)
done = 0
while done < a.steps - last:                       # chunks of 1000 steps (the noise plane is per chunk)
    k = min(1000, a.steps - last - done)
    eng.run(k, torch.randn((k, n), generator=g, device="cuda"))
    done += k
rse0 = eng.env_field("RSE", 2)[0].clone()
eng.run(last, torch.randn((last, n), generator=g, device="cuda"))
st = eng.stats()
err_deg = torch.rad2deg((eng.env_field("RSE", 2)[0] - rse0) / last)
alive = ~st["diverged"]
e = err_deg[alive].cpu().numpy()
print(json.dumps({"plant": a.plant, "agents": n, "steps": a.steps, "diverged": int((~alive).sum()),
                  "mean_abs_theta_error_last_10s_deg": {"median": float(np.median(e)), "p25": float(np.percentile(e, 25)),
                                                         "p75": float(np.percentile(e, 75)), "p90": float(np.percentile(e, 90))},
                  "fraction_below_1deg": float((e < 1.0).mean()), "fraction_below_3deg": float((e < 3.0).mean()),
                  "peak_nz_g_median": float(st["nz_peak"][alive].median())}))
