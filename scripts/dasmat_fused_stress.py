"""Stress of the fused env + agent kernel on the translated plant: absurd learning rates, unit-variance initial weights and a c.g. shift
for everybody drive most agents to divergence -- the CTA-wide vote that disables the in-model barrier, the freeze of diverged agents
and the plant itself must survive it (no hang, no error bits).   python scripts/dasmat_fused_stress.py"""
import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from rl4afcs_b200 import nl_engine, _lib  # noqa: E402,F401
from oracle import nl_c  # noqa: E402
n=3000
eng=nl_engine.NlEngine(n, policy="mixed", plant="dasmat")
eng.set_hpi("FAULT_STEP", 200); eng.set_hpi("FAULT_DAMP", 5)     # c.g. shift for everybody at 2 s
eng.set_hp("ETA_A_H", np.where(np.arange(n)%3==0, 400.0, 25.0)); eng.set_hp("ETA_C_H", np.where(np.arange(n)%5==0, 50.0, 2.0))
eng.set_reference(nl_engine.theta_reference())
w=nl_c.init_weights(n,3, sigma=1.0)
eng.init(w["W1a"],w["W2a"],w["W1c"],w["W2c"])
g=torch.Generator(device="cuda").manual_seed(0)
for k in range(3):
    nz=torch.randn((500,n),generator=g,device="cuda")
    eng.run(500,nz)
    torch.cuda.synchronize()
    print(k, 'diverged', int(eng.stats()["diverged"].sum()), 'of', n, flush=True)
