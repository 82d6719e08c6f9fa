"""Device-side engine of the nonlinear path: SoA state planes, parameters, calls into the C-ABI
(rl4_nl_init / rl4_nl_run / rl4_nl_env_step)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import NHP, NHPI, NLE, NLI, NLN
from .sp_engine import _collapse


def theta_reference(t_end=90, dt=0.01):
    """Pitch-attitude reference of idhp_nonlin.py:36-48 (trim 0.0576 rad + tapered sine sum for 45 s +
    ramp / hold 15 deg / ramp)."""
    n = int(t_end / dt)
    th = 0.0576 + np.zeros(n)
    th[:4500] += np.deg2rad(5) * np.sin(2 * np.pi * np.linspace(0, 45, 4500) / 15) * (np.linspace(2.0, 0.8, 4500))
    th[:4500] += np.deg2rad(4) * np.sin(2 * np.pi * np.linspace(0, 45, 4500) / 30) * (np.linspace(2.0, 0.8, 4500))
    th[5500:6500] += np.deg2rad(1.5) * np.linspace(0, 10, 1000)
    th[6500:7500] += np.deg2rad(15)
    th[7500:8500] += np.deg2rad(1.5) * np.linspace(10, 0, 1000)
    return th


def split_fault(name):
    """Substring matching of envs/nonlinear/env.py:134-158 (if / elif chains), e.g.
    'damp_elevator_and_saturate_elevator' (idhp_nonlin.py:75) -> (damp kind, saturation kind)."""
    name = name or "none"
    damp = sat = 0
    for key in ("damp_elevator", "damp_aileron", "damp_rudder", "damp_all", "shift_cg", "slow_all"):
        if key in name:
            damp = _lib.NL_DAMP[key]
            break
    for key in ("saturate_elevator", "saturate_aileron", "saturate_rudder"):
        if key in name:
            sat = _lib.NL_SAT[key]
            break
    return damp, sat


_DASMAT_IMAGE = {}      # device index -> the model's memory image after initialize() (shared by every engine on that device)
_DASMAT_TRIM = {}       # (device index, trim input, steps) -> (state words of one trimmed aircraft, what the last reset call returned)


class DasmatPlant:
    """The reference's own aircraft model on the GPU (`_citation.initialize/step`, envs/nonlinear/citation.py:62-69), translated
    from its binary at build time (csrc/dasmat_plant.cu).  Holds the per-aircraft model memory of one batch."""

    def __init__(self, lib, device, n, stride):
        if not lib.rl4_dasmat_available():
            raise _lib.Rl4Error("plant='dasmat': this librl4afcs_b200.so was built without the reference's plant binary "
                                "(csrc/_gen/ is produced by tools/lift_plant.py where /root/reference exists)")
        self.lib, self.device, self.n, self.stride = lib, device, n, stride
        self.words = lib.rl4_dasmat_state_words()
        self.word_x, self.word_engine = lib.rl4_dasmat_word_x(), lib.rl4_dasmat_word_engine()
        if device.index not in _DASMAT_IMAGE:
            img = torch.zeros(lib.rl4_dasmat_image_bytes(), dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                _lib.check(lib.rl4_dasmat_initialize(img.data_ptr(), self._stream()), "rl4_dasmat_initialize")
            _DASMAT_IMAGE[device.index] = img
        self.image = _DASMAT_IMAGE[device.index]
        self.state = torch.zeros((self.words, stride), dtype=torch.int64, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reset(self, trim_input, n_steps):
        """Ce500NonLinear.reset (envs/nonlinear/env.py:278-291): initialize(), then `n_steps` calls of step(trim input).  Every
        aircraft starts identically, so one warp flies the settling phase once per process and the result is broadcast.
        Returns what the last call returned (12 doubles on the host)."""
        key = (self.device.index, tuple(float(v) for v in trim_input), int(n_steps))
        if key not in _DASMAT_TRIM:
            nb = 32
            st = torch.zeros((self.words, nb), dtype=torch.int64, device=self.device)
            u = torch.tensor(key[1], dtype=torch.float64, device=self.device).reshape(11, 1).repeat(1, nb).contiguous()
            out = torch.zeros((12, nb), dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.rl4_dasmat_reset(self.image.data_ptr(), st.data_ptr(), nb, nb, self._stream()), "rl4_dasmat_reset")
                _lib.check(self.lib.rl4_dasmat_step(self.image.data_ptr(), st.data_ptr(), nb, nb, u.data_ptr(), nb, int(n_steps),
                                                    out.data_ptr(), nb, None, self.err.data_ptr(), self._stream()), "rl4_dasmat_step")
            self.check()
            _DASMAT_TRIM[key] = (st, out[:, 0].cpu().numpy().copy())
        st, x_obs = _DASMAT_TRIM[key]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rl4_dasmat_broadcast(st.data_ptr(), st.shape[1], 0, self.state.data_ptr(), self.stride, self.n,
                                                     self._stream()), "rl4_dasmat_broadcast")
        return x_obs

    def check(self):
        e = int(self.err.item())
        if e:
            self.err.zero_()
            raise _lib.Rl4Error(f"dasmat plant: the translated model raised error bits {e:#x} (1 untranslated path, 2 wild access, "
                                "4 store into the shared image)")

    @property
    def x(self) -> torch.Tensor:
        """(n, 12) the continuous airframe states the model carries [p q r V alpha beta phi theta psi h xe ye]"""
        return self.state[self.word_x:self.word_x + 12, : self.n].view(torch.float64).t()

    @property
    def engine(self) -> torch.Tensor:
        """(n, 4) the engine states (two per engine)"""
        return self.state[self.word_engine:self.word_engine + 4, : self.n].view(torch.float64).t()


class NlEngine:
    def __init__(self, n_agents: int, *, policy: str = "mixed", device="cuda", plant: str = "surrogate"):
        if policy not in ("mixed", "fp64"):
            raise _lib.Rl4Error("the nonlinear path supports the 'mixed' and 'fp64' policies")
        if plant not in ("surrogate", "dasmat"):
            raise _lib.Rl4Error("plant must be 'surrogate' (calibrated stand-in, fast) or 'dasmat' (the reference's own model)")
        if plant == "dasmat" and policy != "mixed":
            raise _lib.Rl4Error("plant='dasmat' runs with the 'mixed' policy (float32 networks, the reference's own arithmetic)")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.Rl4Error("rl4afcs_b200 runs on B200 GPUs only (no CPU fallback)")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        _lib.check(self.lib.rl4_device_check(idx), "rl4_device_check")
        self.n = int(n_agents)
        self.stride = max(self.n, 1)
        self.policy = policy
        self.policy_id = _lib.POLICY[policy]
        self.tn = torch.float32 if policy == "mixed" else torch.float64
        self.env = torch.zeros((NLE["COUNT"], self.stride), dtype=torch.float64, device=self.device)
        self.net = torch.zeros((NLN["COUNT"], self.stride), dtype=self.tn, device=self.device)
        self.ints = torch.zeros((NLI["COUNT"], self.stride), dtype=torch.int32, device=self.device)
        self.params = _lib.NlParams()
        _lib.check(self.lib.rl4_nl_default_params(ctypes.byref(self.params)), "rl4_nl_default_params")
        self._keep = {}
        self.theta_ref = None
        self.k = 0
        self.plant = plant
        self.dasmat = DasmatPlant(self.lib, self.device, self.n, self.stride) if plant == "dasmat" else None
        self.trim_obs = None

    def set_hp(self, name, value):
        j = NHP[name]
        value = _collapse(value)
        if np.ndim(value) == 0:
            self.params.hp[j] = float(value); self.params.hp_agent[j] = None; self._keep.pop(("hp", j), None)
        else:
            t = torch.as_tensor(np.asarray(value, dtype=np.float64)).to(self.device).contiguous()
            assert t.numel() == self.n
            self._keep[("hp", j)] = t; self.params.hp[j] = float(t[0]); self.params.hp_agent[j] = t.data_ptr()

    def set_hpi(self, name, value):
        j = NHPI[name]
        value = _collapse(value)
        if np.ndim(value) == 0:
            self.params.hpi[j] = int(value); self.params.hpi_agent[j] = None; self._keep.pop(("hpi", j), None)
        else:
            t = torch.as_tensor(np.asarray(value, dtype=np.int32)).to(self.device).contiguous()
            assert t.numel() == self.n
            self._keep[("hpi", j)] = t; self.params.hpi[j] = int(t[0]); self.params.hpi_agent[j] = t.data_ptr()

    def set_reference(self, theta_ref):
        self.theta_ref = torch.as_tensor(np.asarray(theta_ref, dtype=np.float64)).to(self.device).contiguous()

    def state_struct(self):
        return _lib.NlState(self.env.data_ptr(), self.net.data_ptr(), self.ints.data_ptr(), self.stride)

    def env_field(self, name, count=1):
        o = NLE[name]
        return self.env[o:o + count, : self.n]

    def net_field(self, name, count=1):
        o = NLN[name]
        return self.net[o:o + count, : self.n]

    def int_field(self, name):
        return self.ints[NLI[name], : self.n]

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def init(self, W1a, W2a, W1c, W2c):
        """W1a (n,40) [(4,10) row-major], W2a (n,10), W1c (n,40), W2c (n,30) [(10,3) row-major]."""
        def plane(v, w):
            t = torch.as_tensor(np.asarray(v, dtype=np.float64) if not torch.is_tensor(v) else v)
            return t.to(device=self.device, dtype=torch.float64).reshape(self.n, w).t().contiguous()
        with torch.cuda.device(self.device):
            a1, a2, c1, c2 = plane(W1a, 40), plane(W2a, 10), plane(W1c, 40), plane(W2c, 30)
            rc = self.lib.rl4_nl_init(self.policy_id, ctypes.byref(self.params), a1.data_ptr(), a2.data_ptr(), c1.data_ptr(),
                                      c2.data_ptr(), self.n, self.state_struct(), self.n, self._stream())
            _lib.check(rc, "rl4_nl_init")
            if self.dasmat is not None:
                # reset on the reference's own model: initialize() + 1001 trim calls (envs/nonlinear/env.py:288-291)
                n_trim = int(10.0 / self.params.dt) + 1
                self.trim_obs = self.dasmat.reset([self.params.trim_input[j] for j in range(11)], n_trim)
                self.env_field("XFULL", 12).copy_(torch.as_tensor(self.trim_obs, device=self.device).reshape(12, 1).expand(12, self.n))
        self.k = 0

    def run(self, n_steps, noise, *, log_agents=0, log_every=1, log_level=1):
        """noise: float32 (n_steps, n_agents) N(0,1) draws (objects.py:1375).  Returns the log
        (rows, fields, log_agents) or None; log_level 1 = compact (_lib.NLL), 2 = full (_lib.NLF), 3 = the MC_test_hparam row (_lib.NLM)."""
        assert self.theta_ref is not None and self.theta_ref.numel() >= self.k + n_steps
        noise = torch.as_tensor(noise, device=self.device).to(torch.float32).contiguous()
        assert noise.shape == (n_steps, self.n)
        lg = _lib.SpLog(None, 0, 1, 0)
        log_t = None
        if log_agents:
            log_agents = min(int(log_agents), self.n)
            rows = (n_steps + log_every - 1) // log_every
            nf = {1: _lib.NLL, 2: _lib.NLF, 3: _lib.NLM}[int(log_level)]["COUNT"]
            log_t = torch.empty((rows, nf, log_agents), dtype=torch.float64, device=self.device)
            lg = _lib.SpLog(log_t.data_ptr(), int(log_level), log_every, log_agents)
        with torch.cuda.device(self.device):
            if self.dasmat is not None:
                dz = self.dasmat
                rc = self.lib.rl4_nl_run_dasmat(self.policy_id, ctypes.byref(self.params), self.theta_ref.data_ptr(), noise.data_ptr(),
                                                self.n, self.k, n_steps, self.state_struct(), self.n, lg, dz.image.data_ptr(),
                                                dz.state.data_ptr(), dz.stride, dz.err.data_ptr(), self._stream())
                _lib.check(rc, "rl4_nl_run_dasmat")
                dz.check()
            else:
                rc = self.lib.rl4_nl_run(self.policy_id, ctypes.byref(self.params), self.theta_ref.data_ptr(), noise.data_ptr(),
                                         self.n, self.k, n_steps, self.state_struct(), self.n, lg, self._stream())
                _lib.check(rc, "rl4_nl_run")
        self.k += n_steps
        return log_t

    def stats_planes(self) -> torch.Tensor:
        """(NLS.COUNT, n) float64 per-agent statistics computed by ``rl4_nl_agent_stats`` (rows: _lib.NLS)."""
        out = torch.empty((_lib.NLS["COUNT"], self.stride), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rl4_nl_agent_stats(self.state_struct(), self.n, out.data_ptr(), self.stride, self._stream()),
                       "rl4_nl_agent_stats")
        return out[:, : self.n]

    def stats(self):
        return {"rse": self.env_field("RSE", 2).t().clone(), "rse_flight": self.env_field("RSE_FLIGHT", 2).t().clone(),
                "nz_peak": self.env_field("NZ_PEAK")[0].clone(),
                "diverged": self.int_field("DIVERGED_STEP") >= 0}
