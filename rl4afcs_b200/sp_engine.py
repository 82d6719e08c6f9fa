"""Device-side engine of the short-period path: SoA state planes (torch CUDA tensors used
only as device memory), the shared / per-agent hyper-parameters, and the calls into the
C-ABI.  The reference-facing classes in ``objects.py`` / ``envs/linear/env.py`` are thin
views over this.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import HP, HPI, SPE, SPI, SPN


# ---- plant of envs/linear/env.py:66-119 and its three fault variants (:127-154) ----------
def ce500_coefficients() -> dict:
    """PH-LAB stability / control derivatives (envs/linear/env.py:66-87)."""
    chord = 2.022
    return {"V": 59.9, "m": 4.5478e3, "c": chord, "S": 24.2, "mu_c": 102.7, "K2_Y": 0.980, "x_cg": 0.3 * chord,
            "C_Za": -5.16, "C_Zadot": -1.43, "C_Zq": -3.86, "C_Zde": -0.6238,
            "C_ma": -0.43, "C_madot": -3.7, "C_mq": -7.04, "C_mde": -1.553}


def state_matrix(d: dict) -> np.ndarray:
    """A of the (alpha, q) short-period model, same operation order as envs/linear/env.py:89-108."""
    Vc = d["V"] / d["c"]
    ucK = d["mu_c"] * d["K2_Y"]
    den = 2 * d["mu_c"] - d["C_Zadot"]
    z_a = Vc * d["C_Za"] / den
    z_q = (2 * d["mu_c"] + d["C_Zq"]) / den
    m_a = Vc ** 2 * (d["C_ma"] + d["C_Za"] * d["C_madot"] / den) / (2 * ucK)
    m_q = Vc * (d["C_mq"] + d["C_madot"] * (2 * d["mu_c"] + d["C_Zq"]) / den) / (2 * ucK)
    return np.array([[z_a, z_q], [m_a, m_q]])


def input_matrix(d: dict) -> np.ndarray:
    """B (elevator), envs/linear/env.py:110-119."""
    Vc = d["V"] / d["c"]
    den = 2 * d["mu_c"] - d["C_Zadot"]
    z_de = Vc * (d["C_Zde"] / den)
    m_de = Vc ** 2 * (d["C_mde"] + d["C_Zde"] * d["C_madot"] / den) / (2 * d["mu_c"] * d["K2_Y"])
    return np.array([[z_de], [m_de]])


def plant_variants():
    """(A, B) for RL4_FAULT_{NONE, INVERT_ELEVATOR, DAMP_ELEVATOR, SHIFT_CG}."""
    d = ce500_coefficients()
    A, B = state_matrix(d), input_matrix(d)
    out = [(A, B), (A, B * -1), (A, B * 0.5)]
    s = dict(d)
    shift = -0.5                                                   # envs/linear/env.py:141-148
    s["C_mq"] += -(s["C_Zq"] + s["C_ma"]) * shift / s["c"] + s["C_Za"] * (shift / s["c"]) ** 2
    s["C_ma"] -= s["C_Za"] * shift / s["c"]
    s["C_Zq"] -= s["C_Za"] * shift / s["c"]
    s["C_madot"] -= s["C_Za"] * shift / s["c"]
    out.append((state_matrix(s), B))
    return out


_TORCH_DT = {"f8": torch.float64, "f4": torch.float32}


def _collapse(value):
    """A per-agent array whose entries are all equal is the shared scalar (the kernels then read it from the constant bank
    and the trace-free / uniform instantiations stay available)."""
    if np.ndim(value) == 0:
        return value
    a = np.asarray(value)
    if a.size and (a == a.flat[0]).all():
        return a.flat[0]
    return value


def policy_dtypes(policy: str):
    """(network dtype, env dtype) of a dtype policy."""
    return {"fp64": (torch.float64, torch.float64), "fp32": (torch.float32, torch.float32),
            "mixed": (torch.float32, torch.float64)}[policy]


class SpEngine:
    """State + parameters of a batch of short-period IDHP agents on one GPU."""

    def __init__(self, n_agents: int, *, policy: str = "mixed", device="cuda", dt: float = 0.02):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.Rl4Error("rl4afcs_b200 runs on B200 GPUs only (no CPU fallback)")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        _lib.check(self.lib.rl4_device_check(idx), "rl4_device_check")
        self.n = int(n_agents)
        self.policy = policy
        self.policy_id = _lib.POLICY[policy]
        self.tn, self.te = policy_dtypes(policy)
        self.stride = max(self.n, 1)
        self.env = torch.zeros((SPE["COUNT"], self.stride), dtype=self.te, device=self.device)
        self.net = torch.zeros((SPN["COUNT"], self.stride), dtype=self.tn, device=self.device)
        self.ints = torch.zeros((SPI["COUNT"], self.stride), dtype=torch.int32, device=self.device)
        self.params = _lib.SpParams()
        self.params.dt = dt
        for v, (A, B) in enumerate(plant_variants()):
            for j in range(4):
                self.params.A[v][j] = float(A.reshape(-1)[j])
            for j in range(2):
                self.params.B[v][j] = float(B.reshape(-1)[j])
        self.params.q3_alias = 1
        self.params.q7_numpy1 = 0     # NEP 50 compare (the mode observed against the verbatim agent); 1 = numpy 1.x, unverified
        self._keep = {}          # per-agent override tensors kept alive
        self.use_traces = False
        self.ref_base = None
        self.k = 0

    # ---- hyper-parameters ------------------------------------------------------------
    def set_hp(self, name: str, value) -> None:
        """Shared scalar or per-agent array (length n_agents) for a float hyper-parameter."""
        j = HP[name]
        value = _collapse(value)
        if np.ndim(value) == 0:
            self.params.hp[j] = float(value)
            self.params.hp_agent[j] = None
            self._keep.pop(("hp", j), None)
        else:
            t = torch.as_tensor(np.asarray(value, dtype=np.float64)).to(self.device).contiguous()
            assert t.numel() == self.n, f"{name}: expected {self.n} values"
            self._keep[("hp", j)] = t
            self.params.hp[j] = float(t[0])
            self.params.hp_agent[j] = t.data_ptr()

    def set_hpi(self, name: str, value) -> None:
        j = HPI[name]
        value = _collapse(value)
        if np.ndim(value) == 0:
            self.params.hpi[j] = int(value)
            self.params.hpi_agent[j] = None
            self._keep.pop(("hpi", j), None)
        else:
            t = torch.as_tensor(np.asarray(value, dtype=np.int32)).to(self.device).contiguous()
            assert t.numel() == self.n, f"{name}: expected {self.n} values"
            self._keep[("hpi", j)] = t
            self.params.hpi[j] = int(t[0])
            self.params.hpi_agent[j] = t.data_ptr()
        if name in ("ELIG_A", "ELIG_C"):
            self.use_traces = self._any_traces()

    def _any_traces(self) -> bool:
        for nm in ("ELIG_A", "ELIG_C"):
            j = HPI[nm]
            t = self._keep.get(("hpi", j))
            if t is not None:
                if bool((t != 0).any()):
                    return True
            elif self.params.hpi[j] != 0:
                return True
        return False

    def set_reference(self, ref_base) -> None:
        self._ref_host = np.ascontiguousarray(np.asarray(ref_base, dtype=np.float64))
        self.ref_base = torch.as_tensor(self._ref_host).to(self.device).contiguous()

    # ---- state views -------------------------------------------------------------------
    def state_struct(self) -> _lib.SpState:
        return _lib.SpState(self.env.data_ptr(), self.net.data_ptr(), self.ints.data_ptr(), self.stride)

    def env_field(self, name: str, count: int = 1) -> torch.Tensor:
        o = SPE[name]
        return self.env[o:o + count, : self.n]

    def net_field(self, name: str, count: int = 1) -> torch.Tensor:
        o = SPN[name]
        return self.net[o:o + count, : self.n]

    def int_field(self, name: str) -> torch.Tensor:
        return self.ints[SPI[name], : self.n]

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- calls -------------------------------------------------------------------------
    def init(self, x0, W1a, W2a, W1c, W2c) -> None:
        """x0 (n,2); W1a, W2a, W1c (n,4); W2c (n,8) -- any float arrays / tensors."""
        def plane(v, width):
            t = torch.as_tensor(np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v)
            t = t.to(device=self.device, dtype=torch.float64).reshape(self.n, width)
            return t.t().contiguous()
        if self.n == 0:
            self.k = 0
            return
        with torch.cuda.device(self.device):
            px0, pa1, pa2, pc1, pc2 = plane(x0, 2), plane(W1a, 4), plane(W2a, 4), plane(W1c, 4), plane(W2c, 8)
            rc = self.lib.rl4_sp_init(self.policy_id, ctypes.byref(self.params), px0.data_ptr(), pa1.data_ptr(),
                                      pa2.data_ptr(), pc1.data_ptr(), pc2.data_ptr(), self.n,
                                      self.state_struct(), self.n, self._stream())
            _lib.check(rc, "rl4_sp_init")
        self.k = 0

    def run(self, n_steps: int, *, log_level: int = 0, log_agents: int = 0, log_every: int = 1):
        """Advance every agent by n_steps from the current step counter.  Returns the log
        tensor (rows, fields, log_agents) in float64, or None."""
        assert self.ref_base is not None, "set_reference() first"
        assert self.ref_base.numel() >= self.k + n_steps, "reference table too short"
        log_t = None
        lg = _lib.SpLog(None, 0, 1, 0)
        if log_level:
            log_agents = min(int(log_agents), self.n)
            nf = _lib.LF["COUNT"] if log_level == _lib.LOG_FULL else _lib.LB["COUNT"]
            rows = (n_steps + log_every - 1) // log_every
            log_t = torch.zeros((rows, nf, max(log_agents, 1)), dtype=torch.float64, device=self.device)
            lg = _lib.SpLog(log_t.data_ptr(), log_level, log_every, log_agents)
        with torch.cuda.device(self.device):
            rc = self.lib.rl4_sp_run(self.policy_id, ctypes.byref(self.params), self.ref_base.data_ptr(),
                                     self.k, n_steps, self.state_struct(), self.n, int(self.use_traces), lg,
                                     self._stream())
            _lib.check(rc, "rl4_sp_run")
        self.k += n_steps
        return log_t

    # ---- episode statistics (functions.py:39-60; utils.py:350-369) --------------------------
    def stats_planes(self, n_steps=None) -> torch.Tensor:
        """(SPS.COUNT, n) float64 per-agent statistics computed by ``rl4_sp_agent_stats`` (rows: _lib.SPS)."""
        n_steps = self.k if n_steps is None else int(n_steps)
        out = torch.empty((_lib.SPS["COUNT"], self.stride), dtype=torch.float64, device=self.device)
        ref = getattr(self, "_ref_host", None)
        seg = ref[: max(n_steps, 1)] if ref is not None and ref.size else np.zeros(1)
        with torch.cuda.device(self.device):
            rc = self.lib.rl4_sp_agent_stats(self.policy_id, ctypes.byref(self.params), self.state_struct(), self.n, n_steps,
                                             float(seg.min()), float(seg.max()), out.data_ptr(), self.stride, self._stream())
            _lib.check(rc, "rl4_sp_agent_stats")
        return out[:, : self.n]

    def stats(self, n_steps=None) -> dict:
        """Per-agent episode statistics: ``sum_c`` (functions.py:53), ``converged_time`` (utils.py:350-369), ``diverged``
        (functions.py:162), ``mean_abs_e`` and ``nmae`` = mean|e| / (max ref - min ref) -- the normalised tracking error
        BASELINE.json names (the reference has no such statistic; it is an addition of this repo)."""
        S = _lib.SPS
        pl = self.stats_planes(n_steps)
        return {"sum_c": pl[S["SUM_C"]], "converged_time": pl[S["CONV_TIME"]], "diverged": pl[S["DIVERGED"]] != 0,
                "unsteady": pl[S["UNSTEADY"]] != 0, "mean_abs_e": pl[S["MEAN_ABS_E"]], "nmae": pl[S["NMAE"]]}


def default_reference(t_end=60, dt=0.02, period=10):
    """sin(2 pi t / T) on linspace(0, t_end, N) and the 5 deg amplitude (idhp_sp.py:41-44,174)."""
    n = int(t_end / dt)
    t = np.linspace(0, t_end, n)
    return np.sin(2 * np.pi * t / period), float(np.deg2rad(5))


def apply_idhp_config(eng: SpEngine, cfg: dict, *, dt: float) -> None:
    """Map a reference-style ``idhp_config`` (idhp_sp.py:150-173) onto the engine.  Every
    numeric entry may be a scalar or a per-agent array."""
    def per(v, f):
        return f(v) if np.ndim(v) == 0 else np.asarray([f(x) for x in np.asarray(v).ravel()])

    eng.set_hp("GAMMA", cfg["gamma"])
    eng.set_hp("GAMMA_SQ", per(cfg["gamma"], lambda g: g ** 2))                 # objects.py:887
    eng.set_hp("TAU", cfg["tau"])
    eng.set_hp("KAPPA", cfg["kappa"])
    eng.set_hp("LAMBDA_H", cfg["lambda_h"])
    eng.set_hp("LAMBDA_L", cfg["lambda_l"])
    eng.set_hp("ETA_A_H", cfg["actor_config"]["eta_h"])
    eng.set_hp("ETA_A_L", cfg["actor_config"]["eta_l"])
    eng.set_hp("ETA_C_H", cfg["critic_config"]["eta_h"])
    eng.set_hp("ETA_C_L", cfg["critic_config"]["eta_l"])
    eng.set_hp("RLS_GAMMA", cfg["rls_config"]["rls_gamma"])
    eng.set_hp("RLS_COV0", cfg["rls_config"]["rls_cov"])
    eng.set_hp("ERROR_THRESH_DEG", cfg["error_thresh"])
    eng.set_hpi("MULTISTEP", per(cfg["multistep"], lambda m: 1 if m > 0 else 0))      # objects.py:560
    eng.set_hpi("WARMUP_STEPS", per(cfg["warmup_time"], lambda t: int(t / dt)))       # objects.py:793
    eng.set_hpi("COOLDOWN_STEPS", per(cfg["cooldown_time"], lambda t: int(t / dt)))   # objects.py:569

    def elig(v):
        if isinstance(v, (list, tuple, np.ndarray)):
            return np.asarray([_lib.ELIG[x] for x in v], dtype=np.int32)
        return _lib.ELIG[v]

    eng.set_hpi("ELIG_A", elig(cfg["actor_config"]["elig"]))
    eng.set_hpi("ELIG_C", elig(cfg["critic_config"]["elig"]))


def truncated_normal_weights(n: int, seed: int, sigma, device, *, repeat: int = 1) -> dict:
    """TruncatedNormal(0, sigma), re-drawn beyond 2 sigma (the keras initializer of
    objects.py:74), from torch's Philox generator on the device.  TensorFlow's own stream is not
    reproducible outside TF, so weights are always explicit data here.

    ``sigma`` may be a per-agent array (length n * repeat).  ``repeat`` > 1 draws the standard-normal numbers for ``n``
    agents once and tiles them ``repeat`` times before scaling: a Monte-Carlo over ``repeat`` hyper-parameter sets x ``n``
    seeds in which seed s starts every set from the same draw, scaled by that set's sigma (functions.py:80,97: the
    reference re-uses seeds 0..seeds-1 for every config)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    if np.ndim(sigma) == 0:
        sig = float(sigma)
    else:
        sig = torch.as_tensor(np.asarray(sigma, dtype=np.float64)).to(device=device, dtype=torch.float32).reshape(-1, 1)
        assert sig.shape[0] == n * repeat, f"sigma: expected {n * repeat} values"

    def draw(width):
        out = torch.randn((n, width), generator=g, device=device, dtype=torch.float32)
        bad = out.abs() > 2.0
        while bool(bad.any()):
            out = torch.where(bad, torch.randn((n, width), generator=g, device=device, dtype=torch.float32), out)
            bad = out.abs() > 2.0
        if repeat > 1:
            out = out.repeat(repeat, 1)
        return (out * sig).to(torch.float64)

    return {"W1a": draw(4), "W2a": draw(4), "W1c": draw(4), "W2c": draw(8)}


def deg2rad(x):
    return x * (math.pi / 180.0)
