"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/sp_oracle.c).

The C file restates the reference's short-period IDHP path (envs/linear/env.py:156-220,
objects.py:39-281,439-549,551-1004).  This module adds what the reference does in
Python before the loop: the plant matrices from the stability derivatives
(envs/linear/env.py:66-119,127-154), the config dict parsing (idhp_sp.py:45-52,150-173)
and the default reference signal (idhp_sp.py:41-44,174).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

POLICY = {"fp64": 0, "fp32": 1, "mixed": 2}
ELIG = {None: 0, "none": 0, "accumulating": 1, "replacing": 2}
TANH = {"libm": 0, "t13": 1}

CFG_DTYPE = np.dtype(
    [
        ("A", "f8", (4,)), ("B", "f8", (2,)), ("A_fault", "f8", (4,)), ("B_fault", "f8", (2,)),
        ("dt", "f8"), ("gamma", "f8"), ("gamma_sq", "f8"), ("tau", "f8"), ("kappa", "f8"),
        ("lambda_h", "f8"), ("lambda_l", "f8"),
        ("eta_a_h", "f8"), ("eta_a_l", "f8"), ("eta_c_h", "f8"), ("eta_c_l", "f8"),
        ("rls_gamma", "f8"), ("rls_cov0", "f8"), ("error_thresh_deg", "f8"), ("ref_amp", "f8"),
        ("multistep", "i4"), ("warmup_steps", "i4"), ("cooldown_steps", "i4"), ("fault_step", "i4"),
        ("elig_a", "i4"), ("elig_c", "i4"), ("q3_alias", "i4"), ("q7_numpy1", "i4"), ("tracked_q", "i4"), ("pad", "i4"),
    ],
    align=True,
)

STATE_DTYPE = np.dtype(
    [
        ("x", "f8", (2,)), ("x_prev", "f8", (2,)), ("a", "f8"), ("a_prev", "f8"),
        ("W1a", "f8", (4,)), ("W2a", "f8", (4,)), ("W1c", "f8", (4,)), ("W2c", "f8", (8,)),
        ("W1t", "f8", (4,)), ("W2t", "f8", (8,)),
        ("Ea", "f8", (8,)), ("Ec", "f8", (24,)),
        ("theta", "f8", (6,)), ("cov", "f8", (9,)),
        ("cgrad_prev", "f8", (2,)), ("M_prev", "f8", (4,)),
        ("eta_a", "f8"), ("eta_c", "f8"), ("gl_a", "f8"), ("gl_c", "f8"),
        ("eps", "f8", (2,)), ("eps_norm", "f8"), ("sum_c", "f8"), ("sum_abs_e", "f8"),
        ("cooldown", "i4"), ("changed", "i4"), ("lr_init", "i4"),
        ("diverged_step", "i4"), ("conv_step", "i4"), ("x_nan", "i4"),
    ],
    align=True,
)

LOG_DTYPE = np.dtype(
    [
        ("t", "f8"), ("x", "f8", (2,)), ("a", "f8"), ("s", "f8", (2,)), ("c", "f8"), ("ref", "f8"),
        ("a_w1", "f8", (4,)), ("a_w2", "f8", (4,)), ("c_w1", "f8", (4,)), ("c_w2", "f8", (8,)),
        ("a_e", "f8", (8,)), ("c_e", "f8", (24,)),
        ("a_all_grad", "f8", (8,)), ("c_all_grad", "f8", (12,)),
        ("a_grad_norm", "f8"), ("c_grad_norm", "f8"),
        ("params", "f8", (6,)), ("cov", "f8", (9,)), ("eps_norm", "f8"), ("eps_abs", "f8", (2,)),
        ("e", "f8"), ("lam", "f8", (2,)), ("lam_t", "f8", (2,)), ("td", "f8", (2,)),
        ("dadz", "f8"), ("M", "f8", (4,)), ("loss_grad", "f8"),
    ],
    align=True,
)

_lib = None


def build(force: bool = False) -> None:
    """Compile the C oracle into oracle/_ref/ (a no-op when the binaries are current)."""
    out = os.path.join(_HERE, "_ref", "liboracle_sp.so")
    if force or not os.path.isfile(out):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    else:
        subprocess.run(["make", "-C", _HERE], check=False, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    fl = line.split(":")[1].split()
                    return "fma" in fl and "avx2" in fl
    except OSError:
        pass
    return False


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    name = "liboracle_sp_fma.so" if _cpu_has_fma() else "liboracle_sp.so"
    path = os.path.join(_HERE, "_ref", name)
    if not os.path.isfile(path):
        build()
    L = ctypes.CDLL(path)
    L.orc_sp_run.restype = ctypes.c_int
    L.orc_sp_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                             ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                             ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]
    L.orc_sp_init.restype = ctypes.c_int
    L.orc_sp_init.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_int64]
    L.orc_sp_env_step.restype = ctypes.c_int
    L.orc_sp_env_step.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5 + [ctypes.c_int64]
    L.orc_sp_rls_update.restype = ctypes.c_int
    L.orc_sp_rls_update.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 7 + [ctypes.c_int64]
    L.orc_tanh_t13_f64_array.restype = None
    L.orc_tanh_t13_f64_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    L.orc_tanh_t13_f32_array.restype = None
    L.orc_tanh_t13_f32_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    L.orc_tanh_t13_f64.restype = ctypes.c_double
    L.orc_tanh_t13_f64.argtypes = [ctypes.c_double]
    L.orc_tanh_t13_f32.restype = ctypes.c_float
    L.orc_tanh_t13_f32.argtypes = [ctypes.c_float]
    assert L.orc_sizeof_cfg() == CFG_DTYPE.itemsize, (L.orc_sizeof_cfg(), CFG_DTYPE.itemsize)
    assert L.orc_sizeof_state() == STATE_DTYPE.itemsize, (L.orc_sizeof_state(), STATE_DTYPE.itemsize)
    assert L.orc_sizeof_logrow() == LOG_DTYPE.itemsize, (L.orc_sizeof_logrow(), LOG_DTYPE.itemsize)
    _lib = L
    return L


# ----------------------------------------------------------------------------------
# plant matrices, restating envs/linear/env.py:66-119 and :127-154 operation by operation
# (Python floats, including the `**2` spellings, so the doubles equal the reference's).
# ----------------------------------------------------------------------------------
def ce500_coeffs() -> dict:
    c = 2.022
    return dict(V=59.9, m=4.5478e3, c=c, S=24.2, mu_c=102.7, K2_Y=0.980, x_cg=0.3 * c,
                C_Za=-5.16, C_Zadot=-1.43, C_Zq=-3.86, C_Zde=-0.6238,
                C_ma=-0.43, C_madot=-3.7, C_mq=-7.04, C_mde=-1.553)


def ce500_A(k: dict) -> np.ndarray:
    Vc = k["V"] / k["c"]
    u_cK2Y = k["mu_c"] * k["K2_Y"]
    z_a = Vc * k["C_Za"] / (2 * k["mu_c"] - k["C_Zadot"])
    z_q = (2 * k["mu_c"] + k["C_Zq"]) / (2 * k["mu_c"] - k["C_Zadot"])
    m_a = Vc ** 2 * (k["C_ma"] + k["C_Za"] * k["C_madot"] / (2 * k["mu_c"] - k["C_Zadot"])) / (2 * u_cK2Y)
    m_q = Vc * (k["C_mq"] + k["C_madot"] * (2 * k["mu_c"] + k["C_Zq"]) / (2 * k["mu_c"] - k["C_Zadot"])) / (2 * u_cK2Y)
    return np.array([[z_a, z_q], [m_a, m_q]])


def ce500_B(k: dict) -> np.ndarray:
    Vc = k["V"] / k["c"]
    muc_czadot = 2 * k["mu_c"] - k["C_Zadot"]
    z_de = Vc * (k["C_Zde"] / muc_czadot)
    m_de = Vc ** 2 * (k["C_mde"] + k["C_Zde"] * k["C_madot"] / muc_czadot) / (2 * k["mu_c"] * k["K2_Y"])
    return np.array([[z_de], [m_de]])


def ce500_fault(k: dict, A: np.ndarray, B: np.ndarray, case):
    """Plant after ``_engage_fault`` (envs/linear/env.py:127-154)."""
    A, B = A.copy(), B.copy()
    if case == "invert_elevator":
        B *= -1
    elif case == "damp_elevator":
        B *= 0.5
    elif case == "shift_cg":
        k = dict(k)
        shift = -0.5
        k["C_mq"] += -(k["C_Zq"] + k["C_ma"]) * shift / k["c"] + k["C_Za"] * (shift / k["c"]) ** 2
        k["C_ma"] -= k["C_Za"] * shift / k["c"]
        k["C_Zq"] -= k["C_Za"] * shift / k["c"]
        k["C_madot"] -= k["C_Za"] * shift / k["c"]
        A = ce500_A(k)
    return A, B


def default_reference(t_end=60, dt=0.02, T=10):
    """sin(2*pi*t/T) on linspace(0, t_end, N) and the amplitude (idhp_sp.py:41-44,174)."""
    n = int(t_end / dt)
    t = np.linspace(0, t_end, n)
    return np.sin(2 * np.pi * t / T), float(np.deg2rad(5))


def default_idhp_config() -> dict:
    """idhp_sp.py:150-173."""
    return {
        "gamma": 0.6, "multistep": 2, "gamma_rls": 1.0, "lambda_h": 0.576, "lambda_l": 0.296,
        "kappa": 1140, "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 3.0, "error_thresh": 1,
        "tau": 0.01, "in_dims": 1,
        "actor_config": {"layers": {4: "tanh", 1: "tanh"}, "eta_h": 3.55, "eta_l": 0.054, "elig": None},
        "critic_config": {"layers": {4: "tanh", 2: "linear"}, "eta_h": 0.338, "eta_l": 0.00, "elig": None},
        "rls_config": {"state_dim": 2, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6},
    }


def make_cfg(idhp_config=None, *, dt=0.02, fault_time=20, fault_scenario=None, ref_amp=None,
             q3_alias=1, q7_numpy1=0, tracked_q=0, n=1) -> np.ndarray:
    """Build ``n`` identical orc_sp_cfg records from reference-style config dicts."""
    ic = default_idhp_config() if idhp_config is None else idhp_config
    k = ce500_coeffs()
    A, B = ce500_A(k), ce500_B(k)
    Af, Bf = ce500_fault(k, A, B, fault_scenario)
    cfg = np.zeros(n, dtype=CFG_DTYPE)
    cfg["A"] = A.reshape(-1)
    cfg["B"] = B.reshape(-1)
    cfg["A_fault"] = Af.reshape(-1)
    cfg["B_fault"] = Bf.reshape(-1)
    cfg["dt"] = dt
    cfg["gamma"] = ic["gamma"]
    cfg["gamma_sq"] = ic["gamma"] ** 2
    cfg["tau"] = ic["tau"]
    cfg["kappa"] = ic["kappa"]
    cfg["lambda_h"], cfg["lambda_l"] = ic["lambda_h"], ic["lambda_l"]
    cfg["eta_a_h"], cfg["eta_a_l"] = ic["actor_config"]["eta_h"], ic["actor_config"]["eta_l"]
    cfg["eta_c_h"], cfg["eta_c_l"] = ic["critic_config"]["eta_h"], ic["critic_config"]["eta_l"]
    cfg["rls_gamma"] = ic["rls_config"]["rls_gamma"]
    cfg["rls_cov0"] = ic["rls_config"]["rls_cov"]
    cfg["error_thresh_deg"] = ic["error_thresh"]
    cfg["ref_amp"] = float(np.deg2rad(5)) if ref_amp is None else ref_amp
    cfg["multistep"] = 1 if ic["multistep"] > 0 else 0
    cfg["warmup_steps"] = int(ic["warmup_time"] / dt)
    cfg["cooldown_steps"] = int(ic["cooldown_time"] / dt)
    cfg["fault_step"] = -1 if fault_scenario is None else int(fault_time / dt)
    cfg["elig_a"] = ELIG[ic["actor_config"]["elig"]]
    cfg["elig_c"] = ELIG[ic["critic_config"]["elig"]]
    cfg["q3_alias"] = q3_alias
    cfg["q7_numpy1"] = q7_numpy1
    cfg["tracked_q"] = tracked_q
    return cfg


def truncated_normal(rng: np.random.Generator, shape, sigma: float) -> np.ndarray:
    """TruncatedNormal(0, sigma) re-drawn beyond 2 sigma (tf.keras initializer semantics,
    objects.py:74); the TF stream itself is not reproducible, weights are explicit inputs."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out * sigma


def init_weights(n: int, seed: int, sigma: float = 0.1) -> dict:
    rng = np.random.default_rng(seed)
    return {
        "W1a": truncated_normal(rng, (n, 4), sigma).astype(np.float32).astype(np.float64),
        "W2a": truncated_normal(rng, (n, 4), sigma).astype(np.float32).astype(np.float64),
        "W1c": truncated_normal(rng, (n, 4), sigma).astype(np.float32).astype(np.float64),
        "W2c": truncated_normal(rng, (n, 8), sigma).astype(np.float32).astype(np.float64),
    }


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def init_states(policy: str, cfg: np.ndarray, x0: np.ndarray, w: dict) -> np.ndarray:
    n = x0.shape[0]
    x0 = np.ascontiguousarray(x0, dtype=np.float64).reshape(n, 2)
    arrs = [np.ascontiguousarray(w[k], dtype=np.float64) for k in ("W1a", "W2a", "W1c", "W2c")]
    st = np.zeros(n, dtype=STATE_DTYPE)
    stride = 0 if cfg.shape[0] == 1 else 1
    if stride:
        assert cfg.shape[0] == n
    rc = lib().orc_sp_init(POLICY[policy], _ptr(cfg), stride, _ptr(x0), *[_ptr(a) for a in arrs], _ptr(st), n)
    if rc != 0:
        raise RuntimeError(f"orc_sp_init failed: {rc}")
    return st


def run(policy: str, cfg: np.ndarray, ref_base: np.ndarray, states: np.ndarray, k0: int, n_steps: int,
        *, tanh: str = "libm", n_log: int = 0):
    """Advance ``states`` in place by ``n_steps``; returns the log (n_log, n_steps) or None."""
    ref_base = np.ascontiguousarray(ref_base, dtype=np.float64)
    assert ref_base.shape[0] >= k0 + n_steps
    assert states.dtype == STATE_DTYPE and states.flags.c_contiguous
    n = states.shape[0]
    stride = 0 if cfg.shape[0] == 1 else 1
    if stride:
        assert cfg.shape[0] == n
    log = np.zeros((n_log, n_steps), dtype=LOG_DTYPE) if n_log else None
    rc = lib().orc_sp_run(POLICY[policy], TANH[tanh], _ptr(cfg), stride, _ptr(ref_base), k0, n_steps,
                          _ptr(states), n, _ptr(log) if n_log else None, n_log)
    if rc != 0:
        raise RuntimeError(f"orc_sp_run failed: {rc}")
    return log


def env_step(policy: str, cfg: np.ndarray, ref_base, stepp: int, x: np.ndarray, action_deg: np.ndarray):
    """Step-API env step for n agents: x (n,2) is advanced in place; returns reward, e, reward_grad[0]."""
    n = x.shape[0]
    assert x.dtype == np.float64 and x.flags.c_contiguous
    ref_base = np.ascontiguousarray(ref_base, dtype=np.float64)
    act = np.ascontiguousarray(action_deg, dtype=np.float64).reshape(n)
    r, e, g = np.empty(n), np.empty(n), np.empty(n)
    rc = lib().orc_sp_env_step(POLICY[policy], _ptr(cfg), 0 if cfg.shape[0] == 1 else 1, _ptr(ref_base), stepp,
                               _ptr(x), _ptr(act), _ptr(r), _ptr(e), _ptr(g), n)
    if rc != 0:
        raise RuntimeError(f"orc_sp_env_step failed: {rc}")
    return r, e, g


def rls_update(policy: str, cfg: np.ndarray, theta: np.ndarray, cov: np.ndarray, dx0, da0, dx1):
    """Step-API RLS update: theta (n,6), cov (n,9) in place; returns eps (n,2), eps_norm (n,)."""
    n = theta.shape[0]
    assert theta.dtype == np.float64 and cov.dtype == np.float64 and theta.flags.c_contiguous and cov.flags.c_contiguous
    dx0 = np.ascontiguousarray(dx0, dtype=np.float64).reshape(n, 2)
    dx1 = np.ascontiguousarray(dx1, dtype=np.float64).reshape(n, 2)
    da0 = np.ascontiguousarray(da0, dtype=np.float64).reshape(n)
    eps, en = np.empty((n, 2)), np.empty(n)
    rc = lib().orc_sp_rls_update(POLICY[policy], _ptr(cfg), 0 if cfg.shape[0] == 1 else 1, _ptr(theta), _ptr(cov),
                                 _ptr(dx0), _ptr(da0), _ptr(dx1), _ptr(eps), _ptr(en), n)
    if rc != 0:
        raise RuntimeError(f"orc_sp_rls_update failed: {rc}")
    return eps, en


def tanh_t13(x: np.ndarray) -> np.ndarray:
    L = lib()
    x = np.ascontiguousarray(x)
    if x.dtype == np.float32:
        y = np.empty_like(x)
        L.orc_tanh_t13_f32_array(_ptr(x), _ptr(y), x.size)
        return y
    x = x.astype(np.float64, copy=False)
    y = np.empty_like(x)
    L.orc_tanh_t13_f64_array(_ptr(x), _ptr(y), x.size)
    return y


def episode_stats(states: np.ndarray, cfg: np.ndarray, ref_base, n_steps: int) -> dict:
    """Per-agent episode statistics from the oracle's final states: ``sum_c`` = sum(c)/kappa (functions.py:53),
    ``converged_time`` (utils.py:350-369), ``diverged`` (functions.py:162; objects.py:991), ``mean_abs_e`` and
    ``nmae`` = mean|e| / (max ref - min ref) with ref = ref_amp * ref_base[:n_steps] (an addition of the new repo:
    BASELINE.json names the statistic, the reference has none -- SURVEY.md section 5.5)."""
    base = np.asarray(ref_base, dtype=np.float64)[: max(int(n_steps), 1)]
    kappa, amp = cfg["kappa"].astype(np.float64), cfg["ref_amp"].astype(np.float64)
    mean_abs_e = states["sum_abs_e"] / float(max(int(n_steps), 1))
    rng = np.abs(amp * base.max() - amp * base.min())
    with np.errstate(invalid="ignore", divide="ignore"):
        return {"sum_c": states["sum_c"] / kappa, "converged_time": states["conv_step"].astype(np.float64) * cfg["dt"],
                "diverged": (states["diverged_step"] >= 0) | (states["x_nan"] != 0),
                "mean_abs_e": mean_abs_e, "nmae": mean_abs_e / rng}
