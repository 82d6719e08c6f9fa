"""Replays the golden trajectories of the reference's plant binary (tests/golden/citation_*.npz) through the translated
plant on the GPU and times it.   python scripts/dasmat_check.py [--agents N] [--steps K]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl4afcs_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--agents", type=int, default=65536)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--scenarios", default="trim,elevator_doublet,aileron_rudder,shift_cg")
args = ap.parse_args()
L = _lib.load()
dev = torch.device("cuda:0")
assert L.rl4_dasmat_available(), _lib.last_error() if hasattr(_lib, "last_error") else "no dasmat"
W = L.rl4_dasmat_state_words()
img = torch.zeros(L.rl4_dasmat_image_bytes(), dtype=torch.uint8, device=dev)
t0 = time.time()
_lib.check(L.rl4_dasmat_initialize(img.data_ptr(), None), "init")
torch.cuda.synchronize()
print(f"initialize: {time.time() - t0:.3f} s, state words {W}")
err = torch.zeros(1, dtype=torch.int32, device=dev)
res = {}
for name in args.scenarios.split(","):
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"citation_{name}.npz"))
    u, x = g["u"], g["x"]
    n = 32
    st = torch.zeros(W, n, dtype=torch.int64, device=dev)
    _lib.check(L.rl4_dasmat_reset(img.data_ptr(), st.data_ptr(), n, n, None), "reset")
    ud = torch.tensor(u.T.copy(), device=dev)          # [11][N]
    out = torch.zeros(u.shape[0], 12, n, dtype=torch.float64, device=dev)
    ucol = torch.zeros(11, n, dtype=torch.float64, device=dev)
    t0 = time.time()
    for k in range(u.shape[0]):
        ucol.copy_(ud[:, k:k + 1].expand(11, n))
        _lib.check(L.rl4_dasmat_step(img.data_ptr(), st.data_ptr(), n, n, ucol.data_ptr(), n, 1, out[k].data_ptr(), n, None, err.data_ptr(), None), "step")
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    assert int(err.item()) == 0, f"device error bits {int(err.item()):#x}"
    assert np.array_equal(o[:, :, 0], o[:, :, n - 1]), "lanes differ"
    d = np.abs(o[:, :, 0] - x)
    scale = np.maximum(np.abs(x).max(axis=0), 1e-3)
    rel = (d / scale).max(axis=0)
    res[name] = {"steps": int(u.shape[0]), "bit_identical_rows": int((o[:, :, 0] == x).all(axis=1).sum()), "max_rel_err_per_state": rel.tolist(),
                 "seconds": time.time() - t0}
    print(name, "rows bit-identical:", res[name]["bit_identical_rows"], "/", u.shape[0], "max rel err %.3g" % rel.max())
# throughput
n = args.agents
st = torch.zeros(W, n, dtype=torch.int64, device=dev)
_lib.check(L.rl4_dasmat_reset(img.data_ptr(), st.data_ptr(), n, n, None), "reset")
u0 = torch.tensor([-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0], dtype=torch.float64, device=dev)
ucol = (u0[:, None] + 0.01 * torch.randn(11, n, dtype=torch.float64, device=dev) * torch.tensor([1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0], device=dev)[:, None]).contiguous()
out = torch.zeros(12, n, dtype=torch.float64, device=dev)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.rl4_dasmat_step(img.data_ptr(), st.data_ptr(), n, n, ucol.data_ptr(), n, args.steps, out.data_ptr(), n, None, err.data_ptr(), None), "step")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{n} aircraft x {args.steps} steps: {ms:.1f} ms -> {n * args.steps / ms * 1e3:.3e} plant steps/s; err bits {int(err.item())}")
res["throughput"] = {"agents": n, "steps": args.steps, "ms": ms, "plant_steps_per_s": n * args.steps / ms * 1e3}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dasmat_check.json", "w"), indent=1)
