"""Small driver for ncu captures of the fused short-period kernel (one policy, one launch)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--policy", default="fp64")
ap.add_argument("--agents", type=int, default=1 << 18)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--warmup", type=int, default=20)
a = ap.parse_args()
eng, x0, w = bench.make_engine(a.policy, a.agents, "cuda:0", 0, a.warmup + a.steps)
ms, launches = bench.timed_run(eng, x0, w, a.warmup, a.steps, None, 1)
print(f"{a.policy}: {a.agents} agents x {a.steps} steps in {ms:.3f} ms -> {a.agents * a.steps / ms / 1e6:.1f} M agent-steps/s/1e3 ({launches} launch)")
