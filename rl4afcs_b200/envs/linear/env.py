"""Batched, GPU-resident ``Ce500ShortPeriod`` -- drop-in for envs/linear/env.py:7-264 of
wingos80/RL4AFCS with a leading batch dimension.

Same constructor config keys (``x0, dt, t_end, fault_time, fault_scenario, reference{tracked_state,
signal}``), same attributes (``A, B, C, D, x0, x, dt, t_end, fault_time, fault_scenario, t, kappa,
stepp, tracked_state, state_reference, x_hist, y_hist, yref_hist``), same ``reset(seed) ->
(obs, reward, terminated, truncated, info)`` / ``step(action)`` 5-tuples, same info keys
(``yref, t, x, e, reward_grad``) and the same quirks: ``obs`` / ``info['x']`` ALIAS the env's own
state tensor, ``e`` uses the pre-integration state (Q12), ``reward_grad`` always sits
in the alpha slot (Q4), the fault engages when ``stepp == int(fault_time/dt)`` (Q11).

Differences, all forced by batching: tensors carry a leading batch axis -- ``x`` is ``(B, 2, 1)``,
``reward`` / ``e`` are ``(B,)``, ``reward_grad`` is ``(B, 1, 2)``, ``A`` / ``B`` are per-agent when
agents have different fault scenarios; ``fault_scenario`` and ``x0`` may be per-agent.
The arithmetic runs in ``rl4_sp_env_step`` (csrc/sp_kernels.cu); there is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ... import _lib, sp_engine


class Ce500ShortPeriod:
    def __init__(self, env_config, render_mode=None, *, batch: int = 1, device="cuda", dtype: str = "mixed",
                 record_history=None):
        self.batch = int(batch)
        self.dtype_policy = dtype
        self.dt = env_config["dt"]
        self.t_end = env_config["t_end"]
        self.fault_time = env_config["fault_time"]
        self.fault_scenario = env_config.get("fault_scenario")
        self._engine = sp_engine.SpEngine(self.batch, policy=dtype, device=device, dt=self.dt)
        self.device = self._engine.device
        self._define_coeffs()
        self._variants = sp_engine.plant_variants()
        self.A = self._A()
        self.B = self._B()
        self.C = np.array([[1, 0], [0, 1]])
        self.D = np.array([[0], [0]])
        x0 = env_config["x0"]
        x0 = x0.detach().cpu().numpy() if torch.is_tensor(x0) else np.asarray(x0, dtype=np.float64)
        if x0.size == 2:
            x0 = np.broadcast_to(x0.reshape(1, 2), (self.batch, 2))
        self.x0 = np.ascontiguousarray(x0.reshape(self.batch, 2), dtype=np.float64)
        self.t = 0
        self.kappa = 28                                   # envs/linear/env.py:49 (overwritten by the agent)
        self.stepp = 0
        self.action = 0
        self.tracked_state = env_config["reference"]["tracked_state"][0]
        self.state_reference = np.asarray(env_config["reference"]["signal"][0], dtype=np.float64)
        self.record_history = (self.batch <= 4096) if record_history is None else bool(record_history)
        self._asserts()
        self._set_fault_params()
        self._engine.set_reference(self.state_reference)
        self._engine.set_hp("REF_AMP", 1.0)                # the signal array already carries its amplitude
        self._engine.set_hpi("TRACKED_Q", 1 if self.tracked_state == "q" else 0)      # envs/linear/env.py:180-184
        self._write_x(self.x0)
        self.x_hist, self.y_hist, self.yref_hist = [], [], []

    # ---- reference-shaped helpers -------------------------------------------------------
    def _asserts(self):
        assert self.x0.shape[1] == 2, f"State vector x0 must have the same size as the state matrix A, got {self.x0.shape}"
        assert self.tracked_state in ["alpha", "q"], f"Tracked state must be either 'alpha' or 'q', got {self.tracked_state}"

    def _define_coeffs(self):
        for k, v in sp_engine.ce500_coefficients().items():
            setattr(self, k, v)

    def _A(self):
        return sp_engine.state_matrix(sp_engine.ce500_coefficients())

    def _B(self):
        return sp_engine.input_matrix(sp_engine.ce500_coefficients())

    def _get_c_grad(self, error_scalar):
        g = torch.zeros((self.batch, 1, 2), dtype=self._engine.te, device=self.device)
        g[:, 0, 0 if self.tracked_state == "alpha" else 1] = -2 * error_scalar
        k = self.kappa
        if np.ndim(k):
            k = torch.as_tensor(np.asarray(k, dtype=np.float64), device=self.device).to(g.dtype).reshape(-1, 1, 1)
        return k * g

    def _fault_kinds(self):
        fs = self.fault_scenario
        if isinstance(fs, (list, tuple, np.ndarray)):
            assert len(fs) == self.batch
            return np.asarray([_lib.FAULT[f] for f in fs], dtype=np.int32)
        return _lib.FAULT[fs]

    def _set_fault_params(self):
        kinds = self._fault_kinds()
        self._engine.set_hpi("FAULT_KIND", kinds)
        ft = self.fault_time
        if np.ndim(ft) == 0:
            step = int(ft / self.dt)                       # envs/linear/env.py:128
            self._engine.set_hpi("FAULT_STEP", step if (np.ndim(kinds) or kinds != 0) else -1)
        else:
            self._engine.set_hpi("FAULT_STEP", np.asarray([int(t / self.dt) for t in ft], dtype=np.int32))

    def _engage_fault(self):
        """Mirror of envs/linear/env.py:127-154 for the host-visible A / B attributes; the kernel
        selects the plant variant itself from (stepp, fault_step, fault_kind)."""
        if np.ndim(self.fault_time) == 0 and np.ndim(self._fault_kinds()) == 0:
            if self.stepp == int(self.fault_time / self.dt):
                A, B = self._variants[int(self._fault_kinds())]
                self.A, self.B = A.copy(), B.copy()

    def _write_x(self, x_np):
        eng = self._engine
        eng.env_field("X", 2).copy_(torch.as_tensor(x_np.T.copy()).to(eng.te))

    # ---- state views ----------------------------------------------------------------------
    @property
    def x(self) -> torch.Tensor:
        """(B, 2, 1) view of the state plane (in-place integration target, like ``self.x +=``)."""
        return self._engine.env_field("X", 2).t().unsqueeze(-1)

    # ---- gymnasium-style API ----------------------------------------------------------------
    def reset(self, seed=None):
        self._write_x(self.x0)
        self.t = 0
        self.stepp = 0
        self._define_coeffs()
        self.A, self.B = self._A(), self._B()
        self.x_hist, self.y_hist, self.yref_hist = [], [], []
        self._engine.set_hp("KAPPA", self.kappa)
        self._engine.k = 0
        obs = self.x
        info = {"yref": self.state_reference[self.stepp], "t": 0, "x": self.x, "e": 0,
                "reward_grad": self._get_c_grad(0)}
        return obs, 0, False, False, info

    def step(self, action):
        """action: the elevator command in DEGREES (the caller's ``20*a``), shape (B,1,1) / (B,1) / (B,),
        in the policy's network dtype (float32 in the reference's mix)."""
        eng = self._engine
        act = torch.as_tensor(action, device=self.device).to(eng.tn).reshape(self.batch).contiguous()
        self.action = act
        eng.set_hp("KAPPA", self.kappa)
        xplane = eng.env_field("X", 2)
        if self.record_history:
            y = self.x.clone()
        reward = torch.empty(self.batch, dtype=eng.te, device=self.device)
        e = torch.empty_like(reward)
        rg0 = torch.empty_like(reward)
        with torch.cuda.device(self.device):
            rc = eng.lib.rl4_sp_env_step(eng.policy_id, ctypes.byref(eng.params), eng.ref_base.data_ptr(), self.stepp,
                                         xplane.data_ptr(), act.data_ptr(), reward.data_ptr(), e.data_ptr(),
                                         rg0.data_ptr(), eng.stride, self.batch, eng._stream())
            _lib.check(rc, "rl4_sp_env_step")
        ref = self.state_reference[self.stepp]
        reward_grad = torch.zeros((self.batch, 1, 2), dtype=eng.te, device=self.device)
        reward_grad[:, 0, 0] = rg0                          # envs/linear/env.py:189 (always the alpha slot)
        self.t += self.dt
        self.stepp += 1
        self._engage_fault()
        terminated = self.t >= self.t_end
        info = {"yref": ref, "t": self.t, "x": self.x, "e": e, "reward_grad": reward_grad}
        if self.record_history:
            self.x_hist.append(self.x)                      # the reference appends the live array too
            self.y_hist.append(y)
        self.yref_hist.append(ref)
        return self.x, reward, terminated, False, info

    def render(self, mode="human"):
        pass

    def close(self):
        pass
