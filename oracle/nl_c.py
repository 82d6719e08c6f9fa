"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the nonlinear-path C oracle (oracle/nl_oracle.c),
which restates envs/nonlinear/env.py:60-311 and objects.py:283-437,1006-1564 around the documented
surrogate plant (include/rl4_citation_surrogate.h; the reference's plant is a source-less binary)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import sp_c

PLANT_FIELDS = ["m", "S", "c", "b", "Ixx", "Iyy", "Izz", "Ixz", "g", "CL0", "CLa", "CLq", "CLde", "CLflap", "al_stall",
                "CD0", "CDk", "CDgear", "CDflap", "CDstall", "Cm0", "Cma", "Cmq", "Cmde", "Cmflap", "Cmstall",
                "CYb", "CYp", "CYr", "CYda", "CYdr", "Clb", "Clp", "Clr", "Clda", "Cldr",
                "Cnb", "Cnp", "Cnr", "Cnda", "Cndr", "Tstatic", "TV", "Vref", "xcg_gain",
                "inv_m", "inv_Iyy", "inv_gam", "inv_al_stall", "inv_c", "inv_b"]
PLANT_DTYPE = np.dtype([(f, "f8") for f in PLANT_FIELDS] + [("zeta_per_m", "f8"), ("rho_poly", "f8", (21,)),
                        ("lapse_poly", "f8", (21,))], align=True)

CFG_DTYPE = np.dtype([
    ("plant", PLANT_DTYPE), ("trim_input", "f8", (11,)), ("dt", "f8"),
    ("gamma", "f8"), ("gamma_sq", "f8"), ("tau", "f8"), ("lambda_h", "f8"), ("lambda_l", "f8"), ("lr_decay", "f8"),
    ("eta_a_h", "f8"), ("eta_a_l", "f8"), ("eta_c_h", "f8"), ("eta_c_l", "f8"),
    ("rls_gamma", "f8"), ("rls_cov0", "f8"), ("Q_sym", "f8"), ("lambda_t", "f8"), ("lambda_s", "f8"),
    ("noise_std", "f8", (4,)), ("omega0", "f8"), ("omega_slow", "f8"), ("rate_limit", "f8"),
    ("limit_deg", "f8", (3,)), ("damp_factor", "f8"), ("cg_shift", "f8"), ("sat_limit", "f8", (3,)),
    ("multistep", "i4"), ("warmup_steps", "i4"), ("cooldown_steps", "i4"), ("fault_step", "i4"),
    ("elig_a", "i4"), ("fault_damp", "i4"), ("fault_sat", "i4"), ("integrator", "i4"), ("flight_step", "i4"), ("numpy2", "i4"),
], align=True)

STATE_DTYPE = np.dtype([
    ("x_full", "f8", (12,)), ("x_act", "f8", (3,)), ("s", "f8", (4,)), ("s_prev", "f8", (4,)),
    ("a", "f8"), ("a_prev", "f8"), ("x_lon", "f8", (3,)), ("x_prev_lon", "f8", (3,)),
    ("W1a", "f8", (40,)), ("W2a", "f8", (10,)), ("W1c", "f8", (40,)), ("W2c", "f8", (30,)),
    ("W1t", "f8", (40,)), ("W2t", "f8", (30,)), ("Ea", "f8", (50,)),
    ("theta", "f8", (12,)), ("cov", "f8", (16,)), ("cgrad_prev", "f8", (3,)), ("M_prev", "f8", (9,)),
    ("eta_a", "f8"), ("eta_c", "f8"), ("lambdaa", "f8"), ("lr_a", "f8"), ("lr_c", "f8"), ("gl", "f8"),
    ("eps", "f8", (3,)), ("eps_norm", "f8"), ("rse", "f8", (2,)), ("nz_peak", "f8"), ("rse_flight", "f8", (2,)),
    ("cooldown", "i4"), ("diverged_step", "i4"), ("stepp", "i4"), ("pyfloat_mask", "i4"),
], align=True)

LOG_DTYPE = np.dtype([
    ("x_full", "f8", (12,)), ("s_next", "f8", (4,)), ("a_next", "f8"), ("reward", "f8"), ("e_theta", "f8"),
    ("lam", "f8", (3,)), ("lam_t", "f8", (3,)), ("td", "f8", (3,)), ("dads", "f8", (4,)), ("M", "f8", (9,)),
    ("loss_grad", "f8"), ("a_random", "f8"), ("surf", "f8", (3,)), ("model_input", "f8", (11,)),
    ("eta_a", "f8"), ("rse_step", "f8", (2,)), ("yref_theta", "f8"), ("W1a", "f8", (40,)), ("W2a", "f8", (10,)),
    ("W1c", "f8", (40,)), ("W2c", "f8", (30,)), ("a_grad", "f8", (50,)), ("c_grad", "f8", (70,)), ("theta", "f8", (12,)),
    ("cov", "f8", (16,)), ("eps", "f8", (3,)), ("eps_norm", "f8"), ("wa_norm", "f8"), ("wc_norm", "f8"),
], align=True)

FAULT_DAMP = {None: 0, "none": 0, "damp_elevator": 1, "damp_aileron": 2, "damp_rudder": 3, "damp_all": 4,
              "shift_cg": 5, "slow_all": 6}
FAULT_SAT = {None: 0, "none": 0, "saturate_elevator": 1, "saturate_aileron": 2, "saturate_rudder": 3}
INTEGRATOR = {"rk4": 0, "ode5": 1}

_bound = False


def lib():
    global _bound
    L = sp_c.lib()
    if not _bound:
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.orc_nl_default_cfg.argtypes = [vp]
        L.orc_nl_init.argtypes = [ctypes.c_int, vp, ctypes.c_int, vp, vp, vp, vp, vp, i64]
        L.orc_nl_run.argtypes = [ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int,
                                 vp, i64, vp, i64]
        assert L.orc_nl_sizeof_cfg() == CFG_DTYPE.itemsize, (L.orc_nl_sizeof_cfg(), CFG_DTYPE.itemsize)
        assert L.orc_nl_sizeof_state() == STATE_DTYPE.itemsize, (L.orc_nl_sizeof_state(), STATE_DTYPE.itemsize)
        assert L.orc_nl_sizeof_logrow() == LOG_DTYPE.itemsize, (L.orc_nl_sizeof_logrow(), LOG_DTYPE.itemsize)
        _bound = True
    return L


_ext_keep = []


def set_external_plant(n_agents: int, variant: str = "extended_input"):
    """Runs the oracle's env on the reference's OWN aircraft model (the translated binary, oracle/pe_probe/lifted.py), one
    instance per agent, instead of the surrogate header; ``set_external_plant(0)`` switches back.  Single-threaded."""
    L = lib()
    L.orc_nl_set_external_plant.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    _ext_keep.clear()
    if not n_agents:
        L.orc_nl_set_external_plant(None, None, None)
        return None
    from oracle.pe_probe import lifted
    P = lifted.lib(variant)
    crafts = [lifted.Aircraft(variant) for _ in range(int(n_agents))]
    arr = (ctypes.c_void_p * len(crafts))(*[c.h.value for c in crafts])
    _ext_keep.extend([crafts, arr, P])
    L.orc_nl_set_external_plant(ctypes.cast(P.cit_lifted_step, ctypes.c_void_p), ctypes.cast(P.cit_lifted_initialize, ctypes.c_void_p), arr)
    return crafts


def split_fault(name):
    """The reference matches fault names by substring in an if/elif chain
    (envs/nonlinear/env.py:134-158), e.g. 'damp_elevator_and_saturate_elevator' (idhp_nonlin.py:75)."""
    name = name or "none"
    damp = 0
    for key in ("damp_elevator", "damp_aileron", "damp_rudder", "damp_all", "shift_cg", "slow_all"):
        if key in name:
            damp = FAULT_DAMP[key]
            break
    sat = 0
    for key in ("saturate_elevator", "saturate_aileron", "saturate_rudder"):
        if key in name:
            sat = FAULT_SAT[key]
            break
    return damp, sat


def make_cfg(n=1, *, fault=None, fault_time=60, integrator="ode5", elig_a="accumulating", multistep=0, **over):
    cfg = np.zeros(n, dtype=CFG_DTYPE)
    one = np.zeros(1, dtype=CFG_DTYPE)
    lib().orc_nl_default_cfg(one.ctypes.data_as(ctypes.c_void_p))
    cfg[:] = one[0]
    damp, sat = split_fault(fault)
    cfg["fault_damp"], cfg["fault_sat"] = damp, sat
    cfg["fault_step"] = -1 if (damp == 0 and sat == 0) else int(fault_time / 0.01)
    cfg["integrator"] = INTEGRATOR[integrator]
    cfg["elig_a"] = sp_c.ELIG[elig_a]
    cfg["multistep"] = 1 if multistep else 0
    for k, v in over.items():
        cfg[k] = v
        if k == "gamma":
            cfg["gamma_sq"] = np.asarray(v) ** 2
    return cfg


def theta_reference(t_end=90, dt=0.01):
    """idhp_nonlin.py:36-48."""
    n = int(t_end / dt)
    th = 0.0576 + np.zeros(n)
    th[:4500] += np.deg2rad(5) * np.sin(2 * np.pi * np.linspace(0, 45, 4500) / 15) * (np.linspace(2.0, 0.8, 4500))
    th[:4500] += np.deg2rad(4) * np.sin(2 * np.pi * np.linspace(0, 45, 4500) / 30) * (np.linspace(2.0, 0.8, 4500))
    th[5500:6500] += np.deg2rad(1.5) * np.linspace(0, 10, 1000)
    th[6500:7500] += np.deg2rad(15)
    th[7500:8500] += np.deg2rad(1.5) * np.linspace(10, 0, 1000)
    return th


def init_weights(n, seed, sigma=0.1):
    rng = np.random.default_rng(seed)
    f = lambda w: sp_c.truncated_normal(rng, (n, w), sigma).astype(np.float32).astype(np.float64)  # noqa: E731
    return {"W1a": f(40), "W2a": f(10), "W1c": f(40), "W2c": f(30)}


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def init_states(policy, cfg, w, n):
    st = np.zeros(n, dtype=STATE_DTYPE)
    arrs = [np.ascontiguousarray(w[k], dtype=np.float64) for k in ("W1a", "W2a", "W1c", "W2c")]
    rc = lib().orc_nl_init(sp_c.POLICY[policy], _ptr(cfg), 0 if cfg.shape[0] == 1 else 1, *[_ptr(a) for a in arrs], _ptr(st), n)
    if rc:
        raise RuntimeError(f"orc_nl_init failed: {rc}")
    return st


def run(policy, cfg, theta_ref, noise, states, k0, n_steps, *, tanh="libm", n_log=0):
    """noise: float32 (n_steps, n_agents)."""
    n = states.shape[0]
    theta_ref = np.ascontiguousarray(theta_ref, dtype=np.float64)
    noise = np.ascontiguousarray(noise, dtype=np.float32)
    assert noise.shape == (n_steps, n) and theta_ref.shape[0] >= k0 + n_steps
    log = np.zeros((n_log, n_steps), dtype=LOG_DTYPE) if n_log else None
    rc = lib().orc_nl_run(sp_c.POLICY[policy], sp_c.TANH[tanh], _ptr(cfg), 0 if cfg.shape[0] == 1 else 1, _ptr(theta_ref),
                          _ptr(noise), k0, n_steps, _ptr(states), n, _ptr(log) if n_log else None, n_log)
    if rc:
        raise RuntimeError(f"orc_nl_run failed: {rc}")
    return log


def rls_update(gamma, theta, cov, dx0, da0, dx1):
    """Step-level RLS.update, n = 3, m = 1 (objects.py:492-543); theta (n,12) and cov (n,16) are updated in place."""
    L = lib()
    L.orc_nl_rls_update.argtypes = [ctypes.c_double] + [ctypes.c_void_p] * 7 + [ctypes.c_int64]
    n = theta.shape[0]
    dx0 = np.ascontiguousarray(np.broadcast_to(dx0, (n, 3)), dtype=np.float64)
    dx1 = np.ascontiguousarray(np.broadcast_to(dx1, (n, 3)), dtype=np.float64)
    da0 = np.ascontiguousarray(np.broadcast_to(np.asarray(da0, dtype=np.float64).reshape(-1), (n,)), dtype=np.float64)
    eps, en = np.zeros((n, 3)), np.zeros(n)
    L.orc_nl_rls_update(float(gamma), _ptr(theta), _ptr(cov), _ptr(dx0), _ptr(da0), _ptr(dx1), _ptr(eps), _ptr(en), n)
    return eps, en


def env_step(cfg, theta_ref_k, act, x_full, x_act, stepp):
    """One Ce500NonLinear.step without the agent (envs/nonlinear/env.py:182-256) for ONE aircraft; x_full (12,) and
    x_act (3,) are updated in place (x_full is the plant's carried state).  Returns dict(surf, u, e, reward, x_obs) with
    x_obs = what model.step returned, i.e. the state before this step (output-then-update)."""
    L = lib()
    L.orc_nl_env_step.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                  ctypes.c_int32] + [ctypes.c_void_p] * 5
    act = np.ascontiguousarray(act, dtype=np.float64)
    surf, u, e, r, xo = np.zeros(3), np.zeros(11), np.zeros(3), np.zeros(1), np.zeros(12)
    L.orc_nl_env_step(_ptr(cfg), float(theta_ref_k), _ptr(act), _ptr(x_full), _ptr(x_act), int(stepp), _ptr(surf), _ptr(u),
                      _ptr(e), _ptr(r), _ptr(xo))
    return dict(surf=surf, u=u, e=e, reward=float(r[0]), x_obs=xo)
