"""State conversion between the nonlinear oracle's AoS records and the NlEngine SoA planes."""
import numpy as np

from rl4afcs_b200._lib import NLE, NLI, NLN

ENV_MAP = [("x_full", "XFULL", 12), ("x_act", "XACT", 3), ("x_lon", "XLON", 3), ("x_prev_lon", "XPREVLON", 3),
           ("theta", "THETA", 12), ("cov", "COV", 16), ("eps", "EPS", 3), ("rse", "RSE", 2), ("Ea", "EA", 50),
           ("rse_flight", "RSE_FLIGHT", 2)]
ENV_SCALARS = [("eps_norm", "EPS_NORM"), ("nz_peak", "NZ_PEAK"), ("eta_a", "ETA_A"), ("eta_c", "ETA_C"),
               ("lambdaa", "LAMBDAA"), ("gl", "GL")]
NET_MAP = [("s", "S", 4), ("s_prev", "SPREV", 4), ("W1a", "W1A", 40), ("W2a", "W2A", 10), ("W1c", "W1C", 40),
           ("W2c", "W2C", 30), ("W1t", "W1T", 40), ("W2t", "W2T", 30), ("M_prev", "MPREV", 9)]
NET_SCALARS = [("a", "A"), ("a_prev", "APREV"), ("lr_a", "LR_A"), ("lr_c", "LR_C")]


def engine_to_oracle(eng, nl_c):
    env = eng.env[:, : eng.n].cpu().numpy()
    net = eng.net[:, : eng.n].double().cpu().numpy()
    ints = eng.ints[:, : eng.n].cpu().numpy()
    st = np.zeros(eng.n, dtype=nl_c.STATE_DTYPE)
    for f, k, w in ENV_MAP:
        st[f] = env[NLE[k]:NLE[k] + w].T
    for f, k in ENV_SCALARS:
        st[f] = env[NLE[k]]
    st["cgrad_prev"][:, 2] = env[NLE["CGRAD_PREV"]]
    for f, k, w in NET_MAP:
        st[f] = net[NLN[k]:NLN[k] + w].T
    for f, k in NET_SCALARS:
        st[f] = net[NLN[k]]
    st["cooldown"] = ints[NLI["COOLDOWN"]]
    st["diverged_step"] = ints[NLI["DIVERGED_STEP"]]
    st["stepp"] = ints[NLI["STEPP"]]
    st["pyfloat_mask"] = ints[NLI["PYFLOAT_MASK"]]
    return st


def oracle_to_engine(st, eng):
    import torch

    n = eng.n
    env = np.zeros((NLE["COUNT"], n))
    net = np.zeros((NLN["COUNT"], n))
    ints = np.zeros((NLI["COUNT"], n), dtype=np.int32)
    for f, k, w in ENV_MAP:
        env[NLE[k]:NLE[k] + w] = st[f].T
    for f, k in ENV_SCALARS:
        env[NLE[k]] = st[f]
    env[NLE["CGRAD_PREV"]] = st["cgrad_prev"][:, 2]
    for f, k, w in NET_MAP:
        net[NLN[k]:NLN[k] + w] = st[f].T
    for f, k in NET_SCALARS:
        net[NLN[k]] = st[f]
    ints[NLI["COOLDOWN"]] = st["cooldown"]
    ints[NLI["DIVERGED_STEP"]] = st["diverged_step"]
    ints[NLI["STEPP"]] = st["stepp"]
    ints[NLI["PYFLOAT_MASK"]] = st["pyfloat_mask"]
    eng.env[:, :n] = torch.as_tensor(env).to(eng.device)
    eng.net[:, :n] = torch.as_tensor(net).to(eng.net.dtype).to(eng.device)
    eng.ints[:, :n] = torch.as_tensor(ints).to(eng.device)


def max_rel(a, b, floor):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b) / np.maximum(np.abs(b), floor)
    d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
    return float(np.nanmax(d)) if d.size else 0.0
