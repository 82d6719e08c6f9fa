"""Helpers shared by the parity tests: move state between the oracle's AoS records and the
engine's SoA planes, and compare them."""
import numpy as np

from rl4afcs_b200 import _lib
from rl4afcs_b200._lib import SPE, SPF, SPI, SPN


def engine_state_to_oracle(eng, oracle_mod, gamma_lambda_of):
    """Engine planes -> array of oracle STATE records.  gamma_lambda_of(low: bool array) -> gl."""
    env = eng.env[:, : eng.n].double().cpu().numpy()
    net = eng.net[:, : eng.n].double().cpu().numpy()
    ints = eng.ints[:, : eng.n].cpu().numpy()
    st = np.zeros(eng.n, dtype=oracle_mod.STATE_DTYPE)
    st["x"] = env[SPE["X"]:SPE["X"] + 2].T
    st["x_prev"] = env[SPE["XPREV"]:SPE["XPREV"] + 2].T
    st["theta"] = env[SPE["THETA"]:SPE["THETA"] + 6].T
    st["cov"] = env[SPE["COV"]:SPE["COV"] + 9].T
    st["cgrad_prev"][:, 0] = env[SPE["CGRAD_PREV"]]
    st["eps"] = env[SPE["EPS"]:SPE["EPS"] + 2].T
    st["eps_norm"] = env[SPE["EPS_NORM"]]
    st["sum_c"] = env[SPE["SUM_C"]]
    st["sum_abs_e"] = env[SPE["SUM_ABS_E"]]
    st["Ea"] = env[SPE["EA"]:SPE["EA"] + 8].T
    h = env[SPE["EC_H"]:SPE["EC_H"] + 4].T
    st["Ec"][:, 0:4] = h
    st["Ec"][:, 8:12] = env[SPE["EC_W1R0"]:SPE["EC_W1R0"] + 4].T
    st["Ec"][:, 16:20] = h
    st["Ec"][:, 20:24] = env[SPE["EC_W1R1"]:SPE["EC_W1R1"] + 4].T
    st["a"] = net[SPN["A"]]
    st["a_prev"] = net[SPN["APREV"]]
    for nm, width in (("W1a", 4), ("W2a", 4), ("W1c", 4), ("W2c", 8), ("W1t", 4), ("W2t", 8)):
        st[nm] = net[SPN[nm.upper()]:SPN[nm.upper()] + width].T
    st["M_prev"] = net[SPN["MPREV"]:SPN["MPREV"] + 4].T
    st["eta_a"] = net[SPN["ETA_A"]]
    st["eta_c"] = net[SPN["ETA_C"]]
    flags = ints[SPI["FLAGS"]]
    st["cooldown"] = ints[SPI["COOLDOWN"]]
    st["changed"] = (flags & SPF["CHANGED"]) != 0
    st["lr_init"] = (flags & SPF["LR_INIT"]) != 0
    st["x_nan"] = (flags & SPF["X_NAN"]) != 0
    st["diverged_step"] = ints[SPI["DIVERGED_STEP"]]
    st["conv_step"] = ints[SPI["CONV_STEP"]]
    gl = gamma_lambda_of((flags & SPF["LAMBDA_LOW"]) != 0)
    st["gl_a"] = gl
    st["gl_c"] = gl
    return st


def oracle_state_to_engine(st, eng, gl_low=None):
    """Oracle STATE records -> engine planes (teacher forcing).  gl_low = lambda_l*gamma."""
    import torch

    n = eng.n
    env = np.zeros((SPE["COUNT"], n))
    net = np.zeros((SPN["COUNT"], n))
    ints = np.zeros((SPI["COUNT"], n), dtype=np.int32)
    env[SPE["X"]:SPE["X"] + 2] = st["x"].T
    env[SPE["XPREV"]:SPE["XPREV"] + 2] = st["x_prev"].T
    env[SPE["THETA"]:SPE["THETA"] + 6] = st["theta"].T
    env[SPE["COV"]:SPE["COV"] + 9] = st["cov"].T
    env[SPE["CGRAD_PREV"]] = st["cgrad_prev"][:, 0]
    env[SPE["EPS"]:SPE["EPS"] + 2] = st["eps"].T
    env[SPE["EPS_NORM"]] = st["eps_norm"]
    env[SPE["SUM_C"]] = st["sum_c"]
    env[SPE["SUM_ABS_E"]] = st["sum_abs_e"]
    env[SPE["EA"]:SPE["EA"] + 8] = st["Ea"].T
    env[SPE["EC_H"]:SPE["EC_H"] + 4] = st["Ec"][:, 0:4].T
    env[SPE["EC_W1R0"]:SPE["EC_W1R0"] + 4] = st["Ec"][:, 8:12].T
    env[SPE["EC_W1R1"]:SPE["EC_W1R1"] + 4] = st["Ec"][:, 20:24].T
    net[SPN["A"]] = st["a"]
    net[SPN["APREV"]] = st["a_prev"]
    for nm, width in (("W1a", 4), ("W2a", 4), ("W1c", 4), ("W2c", 8), ("W1t", 4), ("W2t", 8)):
        net[SPN[nm.upper()]:SPN[nm.upper()] + width] = st[nm].T
    net[SPN["MPREV"]:SPN["MPREV"] + 4] = st["M_prev"].T
    net[SPN["ETA_A"]] = st["eta_a"]
    net[SPN["ETA_C"]] = st["eta_c"]
    ints[SPI["COOLDOWN"]] = st["cooldown"]
    lam_low = (st["gl_a"] == gl_low) if gl_low is not None else np.zeros(n, dtype=bool)
    ints[SPI["FLAGS"]] = (st["changed"] * SPF["CHANGED"] + st["lr_init"] * SPF["LR_INIT"] + st["x_nan"] * SPF["X_NAN"]
                          + lam_low * SPF["LAMBDA_LOW"])
    ints[SPI["DIVERGED_STEP"]] = st["diverged_step"]
    ints[SPI["CONV_STEP"]] = st["conv_step"]
    eng.env[:, :n] = torch.as_tensor(env).to(eng.env.dtype).to(eng.device)
    eng.net[:, :n] = torch.as_tensor(net).to(eng.net.dtype).to(eng.device)
    eng.ints[:, :n] = torch.as_tensor(ints).to(eng.device)


STATE_FLOAT_FIELDS = ("x", "x_prev", "a", "a_prev", "W1a", "W2a", "W1c", "W2c", "W1t", "W2t", "Ea", "Ec", "theta",
                      "cov", "M_prev", "eta_a", "eta_c", "gl_a", "gl_c", "eps", "eps_norm", "sum_c", "sum_abs_e")
STATE_INT_FIELDS = ("cooldown", "changed", "lr_init", "diverged_step", "conv_step", "x_nan")


def state_mismatches(a, b, fields=None):
    """{field: (n_mismatching_agents, max_abs_diff)} for fields that are not bit-identical."""
    out = {}
    for f in (fields or STATE_FLOAT_FIELDS + STATE_INT_FIELDS):
        if f == "cgrad_prev":
            continue
        x, y = a[f], b[f]
        same = (x == y) | (np.isnan(x.astype(float)) & np.isnan(y.astype(float)))
        if not same.all():
            bad = ~same
            rows = bad.reshape(bad.shape[0], -1).any(axis=1)
            with np.errstate(invalid="ignore"):
                d = np.nanmax(np.abs(x.astype(float) - y.astype(float))[bad]) if bad.any() else 0.0
            out[f] = (int(rows.sum()), float(d))
    return out
