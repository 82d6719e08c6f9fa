"""Stress of the translated plant: aircraft driven far outside the flight envelope (hard-over surfaces, throttle chops, c.g. shifts)
until states blow up -- no hang, no out-of-bounds access, error bits reported; the first aircraft are cross-checked against the
CPU build of the same translation while their states are finite.   python scripts/dasmat_stress.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl4afcs_b200 import _lib  # noqa: E402

L = _lib.load()
dev = torch.device("cuda:0")
n, seg, n_seg = 2048, 50, 60
img = torch.zeros(L.rl4_dasmat_image_bytes(), dtype=torch.uint8, device=dev)
_lib.check(L.rl4_dasmat_initialize(img.data_ptr(), None), "init")
st = torch.zeros((L.rl4_dasmat_state_words(), n), dtype=torch.int64, device=dev)
_lib.check(L.rl4_dasmat_reset(img.data_ptr(), st.data_ptr(), n, n, None), "reset")
err = torch.zeros(1, dtype=torch.int32, device=dev)
rng = np.random.default_rng(0)
trim = np.array([-0.02855, 0, 0, 0, 0, 0, 0, 0, 0.55, 0.55, 0])
crafts = []
try:
    from oracle.pe_probe import lifted
    if lifted.available():
        crafts = [lifted.Aircraft() for _ in range(6)]
        for c in crafts:
            c.initialize()
except Exception as e:  # noqa: BLE001
    print("no CPU translation:", e)
worst, n_nan_hist = 0.0, []
scale = np.array([0.1, 0.1, 0.1, 90, 0.06, 0.05, 0.1, 0.06, 0.1, 2000, 1000, 100.0])
for s in range(n_seg):
    u = np.tile(trim, (n, 1))
    u[:, 0] += rng.uniform(-0.5, 0.5, n); u[:, 1] += rng.uniform(-0.6, 0.6, n); u[:, 2] += rng.uniform(-0.4, 0.4, n)
    u[:, 6] = rng.choice([0.0, 0.3, 1.0], n); u[:, 7] = rng.choice([0.0, 1.0], n)
    u[:, 8:10] = rng.uniform(0.0, 1.0, (n, 2)); u[:, 10] = rng.uniform(-1.0, 1.0, n)
    ud = torch.tensor(u.T.copy(), device=dev)
    out_all = torch.zeros((seg, 12, n), dtype=torch.float64, device=dev)
    _lib.check(L.rl4_dasmat_step(img.data_ptr(), st.data_ptr(), n, n, ud.data_ptr(), n, seg, None, n, out_all.data_ptr(), err.data_ptr(), None), "step")
    torch.cuda.synchronize()
    got = out_all.cpu().numpy()
    n_nan_hist.append(int(np.isnan(got[-1]).any(axis=0).sum()))
    for i, c in enumerate(crafts):
        ref = c.run(u[i], seg)
        ok = np.isfinite(ref).all(axis=1) & np.isfinite(got[:, :, i]).all(axis=1) & (np.abs(ref).max(axis=1) < 1e6)
        if ok.any():
            worst = max(worst, float((np.abs(got[ok, :, i] - ref[ok]) / np.maximum(scale, np.abs(ref[ok]))).max()))
print("aircraft with NaN states per segment:", n_nan_hist[::6], "final", n_nan_hist[-1], "of", n)
print("device error bits:", int(err.item()), "(2 = an access outside the model's memory was attempted and redirected)")
print("max relative difference to the CPU translation while finite: %.3g" % worst)
