"""Diagnostics: who diverges and when (nonlinear path), plus input checksums, to compare GPU boxes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl4afcs_b200 import nl_engine  # noqa: E402

n = 1 << 18
p = torch.cuda.get_device_properties(0)
print("device", p.name, "SMs", p.multi_processor_count, "torch", torch.__version__, "driver", torch.version.cuda)
eng = nl_engine.NlEngine(n, policy="mixed")
eng.set_reference(nl_engine.theta_reference())
g = torch.Generator(device="cuda"); g.manual_seed(0)
w = [(torch.randn((n, k), generator=g, device="cuda").clamp_(-2, 2) * 0.1).double() for k in (40, 10, 40, 30)]
print("weight checksums", [float(x.sum()) for x in w], [float(x.abs().max()) for x in w])
eng.init(*w)
nz = torch.randn((600, n), generator=g, device="cuda")
print("noise checksum", float(nz.double().sum()), float(nz.abs().max()))
eng.run(600, nz)
d = eng.int_field("DIVERGED_STEP").cpu().numpy()
bad = np.flatnonzero(d >= 0)
print("diverged", bad.size, "steps hist", np.histogram(d[bad], bins=[0, 10, 50, 100, 200, 300, 400, 500, 600])[0].tolist())
print("first bad indices", bad[:20].tolist())
per_cta = np.bincount(bad // 128, minlength=n // 128)
print("CTAs with >= 1 bad", int((per_cta > 0).sum()), "max per CTA", int(per_cta.max()), "lane hist", np.bincount(bad % 32, minlength=32).tolist())
x = eng.env_field("XFULL", 12).cpu().numpy()
print("x_full of first bad agents:\n", x[:, bad[:3]].T)
