import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import nl_c, sp_c
from tests import test_gpu_nl_parity as T, _util_nl
from rl4afcs_b200 import _lib
sp_c.build(); nl_c.lib()
n, steps = 64, 600
noise = np.random.default_rng(2).standard_normal((steps, n)).astype(np.float32)
def run(chunks, log):
    eng, st, cfg, th = T._setup(nl_c, n, 'mixed', seed=5)
    k=0
    for c in chunks:
        eng.run(c, noise[k:k+c], log_agents=(n if log else 0)); k+=c
    return eng
A=run([600],False); B=run([1,7,92,500],False); C=run([600],True); D=run([300,300],False)
for name,X in (('chunked',B),('logged',C),('2chunks',D)):
    de=(A.env!=X.env)&~(torch.isnan(A.env)&torch.isnan(X.env)); dn=(A.net!=X.net)&~(torch.isnan(A.net)&torch.isnan(X.net))
    print(name,'env fields differing:',sorted(set(de.nonzero()[:,0].tolist())),'net fields:',sorted(set(dn.nonzero()[:,0].tolist()))[:20], 'agents', len(set(de.nonzero()[:,1].tolist())))
    if de.any():
        f,a=de.nonzero()[0].tolist(); print('  example field',f,'agent',a,A.env[f,a].item(),X.env[f,a].item())
