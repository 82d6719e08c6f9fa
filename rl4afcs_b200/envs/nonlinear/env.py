"""Batched, GPU-resident ``Ce500NonLinear`` -- drop-in for envs/nonlinear/env.py:11-319 of wingos80/RL4AFCS
with a leading batch dimension.

Same config keys (``state_dim, action_dim, trim_input, trim_state, dt, t_end, total_steps, fault_time,
fault_scenario, reference{tracked_state, signal}``), same ``reset(seed) -> (MDP_state, reward, None, None, info)``
and ``step(action) -> (MDP_state, reward, None, False, info)`` with the reference's info keys (``nans, s, yref,
action_commanded, action_effective, rates, t, x_full, x, e, RSE, reward_grad``).  The wrapper logic (action scaling,
rate-limited first-order actuators, fault / saturation injection, rewards, MDP state) follows the reference line by
line inside ``rl4_nl_env_step``.  The aircraft itself (``_citation.step``, envs/nonlinear/env.py:210) is selectable:

* ``plant="dasmat"`` (default): the reference's OWN model -- its ``_citation`` Windows binary translated to C at build time and
  compiled for the GPU (csrc/dasmat_plant.cu; every step within ~1e-15 relative of the binary).
  ``reset()`` runs the model's own ``initialize()`` and the 1001 trim calls of envs/nonlinear/env.py:288-291.  Needs a library
  built where the reference tree exists; otherwise the constructor raises and asks for an explicit choice;
* ``plant="surrogate"``: the stand-in of ``include/rl4_citation_surrogate.h``, calibrated against the reference's binary --
  a different (similar) aircraft, ~120x lighter: the plant of the throughput configurations.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ... import _lib, nl_engine


class Ce500NonLinear:
    def __init__(self, env_config, render_mode=None, *, batch: int = 1, device="cuda", dtype: str = "mixed",
                 integrator: str = "ode5", plant: str = "dasmat"):
        self.batch = int(batch)
        self.fault_scenario = env_config["fault_scenario"]
        self.initialized = False
        self.dt = env_config["dt"]
        self.t_end = env_config["t_end"]
        self.total_steps = env_config["total_steps"]
        self.fault_time = env_config["fault_time"]
        self.trim_state = np.asarray(env_config["trim_state"], dtype=np.float64)
        self.trim_input = np.asarray(env_config["trim_input"], dtype=np.float64)
        self.mdp_s_dim = env_config["state_dim"]
        self.t = 0
        self.kappa = [1, 1, 1]
        self.stepp = 0
        self.action = None
        self.tracked_state = env_config["reference"]["tracked_state"]
        self.state_reference = env_config["reference"]["signal"]
        # plant: "dasmat" (default) = the reference's own model translated from its binary -- what the reference flies;
        # "surrogate" = the calibrated stand-in (fast, ~120x lighter), an explicit choice because it is a different aircraft
        if plant == "dasmat" and not _lib.load().rl4_dasmat_available():
            raise _lib.Rl4Error("Ce500NonLinear: this librl4afcs_b200.so was built without the reference's plant binary, so the default "
                                "plant='dasmat' (the reference's own aircraft model) is not available; pass plant='surrogate' for the "
                                "calibrated stand-in")
        if plant == "dasmat" and integrator != "ode5":
            raise _lib.Rl4Error("plant='dasmat' integrates with the model's own fixed-step ode5; integrator='rk4' exists for plant='surrogate'")
        self.plant = plant
        self._engine = nl_engine.NlEngine(self.batch, policy=dtype, device=device, plant=plant)
        self.device = self._engine.device
        eng = self._engine
        eng.params.dt = self.dt
        eng.params.integrator = _lib.INTEGRATOR[integrator]
        for i in range(11):
            eng.params.trim_input[i] = float(self.trim_input[i])
        self._set_saturations()
        self._set_surfaces_dynamics()
        self._set_weight_matrices(self.kappa)
        self._apply_fault_config()
        eng.set_reference(np.asarray(self.state_reference[1], dtype=np.float64))      # theta reference; phi / psi are zero
        for j in (0, 2):
            assert not np.any(np.asarray(self.state_reference[j])), "only the pitch channel carries a reference (idhp_nonlin.py:116-117)"

    # ---- reference-shaped helpers ----
    def _set_weight_matrices(self, kappa):
        self.Q_sym = kappa[1]
        self.Q_asym = np.diag([kappa[0], kappa[2]])
        self._engine.set_hp("Q_SYM", self.Q_sym)

    def _reset_surfaces_states(self):
        self._engine.env_field("XACT", 3).zero_()

    def _set_surfaces_dynamics(self, omega_0=13, reset=True):
        self.act_omega_0 = omega_0
        self._engine.params.omega0 = float(omega_0)
        if reset:
            self._reset_surfaces_states()

    def _set_saturations(self):
        self.limits = {"de": np.array([-15, 15]), "da": np.array([-37, 37]), "dr": np.array([-22, 22])}
        for i, k in enumerate(("de", "da", "dr")):
            self._engine.params.limit_deg[i] = float(self.limits[k][1])

    def _apply_fault_config(self):
        fs, eng = self.fault_scenario, self._engine
        if isinstance(fs, (list, tuple, np.ndarray)):
            ds = np.asarray([nl_engine.split_fault(f) for f in fs], dtype=np.int32)
            eng.set_hpi("FAULT_DAMP", ds[:, 0]); eng.set_hpi("FAULT_SAT", ds[:, 1])
            any_fault = True
        else:
            d, s = nl_engine.split_fault(fs)
            eng.set_hpi("FAULT_DAMP", d); eng.set_hpi("FAULT_SAT", s)
            any_fault = bool(d or s)
        if np.ndim(self.fault_time) == 0:
            eng.set_hpi("FAULT_STEP", int(self.fault_time / self.dt) if any_fault else -1)      # env.py:132
        else:
            eng.set_hpi("FAULT_STEP", np.asarray([int(t / self.dt) for t in self.fault_time], dtype=np.int32))

    @property
    def state(self) -> torch.Tensor:
        """(B, 12) p q r V alpha beta phi theta psi h xe ye as ``model.step`` last RETURNED it (envs/nonlinear/env.py:210,291).
        The reference's plant returns the state before the step it then takes, so this is one sample behind the state
        the plant carries (``plant_state``)."""
        return self._obs

    @property
    def plant_state(self) -> torch.Tensor:
        """(B, 12) the state the plant carries into its next step."""
        if self._engine.dasmat is not None:
            return self._engine.dasmat.x
        return self._engine.env_field("XFULL", 12).t()

    # ---- gymnasium-style API ----
    def reset(self, seed=None):
        eng = self._engine
        z = torch.zeros((self.batch, 1), dtype=torch.float64, device=self.device)
        eng.init(z.expand(self.batch, 40), z.expand(self.batch, 10), z.expand(self.batch, 40), z.expand(self.batch, 30))
        self.initialized = True
        self.stepp = 0
        self.t = 0
        if eng.dasmat is not None:
            obs = eng.trim_obs
        else:
            obs = (ctypes.c_double * 12)()
            _lib.check(eng.lib.rl4_nl_trim_state(ctypes.byref(eng.params), None, obs), "rl4_nl_trim_state")
        self._obs = torch.tensor(list(obs), dtype=torch.float64, device=self.device).repeat(self.batch, 1)   # env.py:291
        MDP_state = torch.zeros((self.batch, self.mdp_s_dim), dtype=torch.float64, device=self.device)
        info = {"nans": False, "s": MDP_state, "yref": np.zeros(3), "action": np.zeros(3), "rates": np.zeros(3), "t": self.t,
                "x_full": self.state, "x": [torch.zeros((self.batch, 3, 1), dtype=torch.float64, device=self.device),
                                            torch.zeros((self.batch, 4, 1), dtype=torch.float64, device=self.device)],
                "e": torch.zeros((self.batch, 3), dtype=torch.float64, device=self.device), "RSE": [0, 0],
                "reward_grad": [np.zeros(2), np.zeros(4)]}
        return MDP_state, torch.zeros((self.batch, 2, 1, 1), dtype=torch.float64, device=self.device), None, None, info

    def step(self, action):
        """action: normalised surface commands in [-1, 1], shape (B, 3) (or (3,) broadcast).  Returns the reference's
        tuple (envs/nonlinear/env.py:182-256): MDP_state (B,4), reward (B,2,1,1) = [longitudinal, lateral], None, False,
        info."""
        eng = self._engine
        act = torch.as_tensor(action, device=self.device, dtype=torch.float64)
        if act.ndim == 1:
            act = act.reshape(1, 3).expand(self.batch, 3)
        act_p = act.t().contiguous()
        B, dev = self.batch, self.device
        mdp = torch.empty((4, B), dtype=torch.float64, device=dev)
        reward_lon = torch.empty(B, dtype=torch.float64, device=dev)
        e_th = torch.empty_like(reward_lon)
        surf = torch.empty((3, B), dtype=torch.float64, device=dev)
        eff = torch.empty((3, B), dtype=torch.float64, device=dev)
        xobs = torch.empty((12, B), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            args = (ctypes.byref(eng.params), eng.theta_ref.data_ptr(), self.stepp,
                    eng.env_field("XFULL", 12).data_ptr(), eng.env_field("XACT", 3).data_ptr(),
                    act_p.data_ptr(), mdp.data_ptr(), reward_lon.data_ptr(), e_th.data_ptr(),
                    surf.data_ptr(), eff.data_ptr(), xobs.data_ptr(), eng.stride, B)
            if eng.dasmat is not None:
                dz = eng.dasmat
                rc = eng.lib.rl4_nl_env_step_dasmat(*args, dz.image.data_ptr(), dz.state.data_ptr(), dz.stride, dz.err.data_ptr(), eng._stream())
                _lib.check(rc, "rl4_nl_env_step_dasmat")
                dz.check()
            else:
                rc = eng.lib.rl4_nl_env_step(*args, eng._stream())
                _lib.check(rc, "rl4_nl_env_step")
        self._obs = xobs.t()                                                            # self.state = model.step(input) (env.py:210)
        ref = [r[self.stepp] for r in self.state_reference]
        self.stepp += 1
        self.t += self.dt
        st = self.state
        err = torch.stack([st[:, 6] - ref[0], e_th, st[:, 8] - ref[2]], dim=1)                # env.py:215 (state - ref)
        rg_lon = torch.zeros((B, 1, 3), dtype=torch.float64, device=dev)
        rg_lon[:, 0, 2] = -self.Q_sym * e_th                                            # env.py:219-220
        k0, k2 = float(self.Q_asym[0, 0]), float(self.Q_asym[1, 1])
        reward_lat = (-0.5 * err[:, 0]) * k0 * err[:, 0] + (-0.5 * err[:, 2]) * k2 * err[:, 2]   # env.py:224 (lateral MDP, unused by IDHPnonlin)
        rg_lat = torch.zeros((B, 1, 4), dtype=torch.float64, device=dev)
        rg_lat[:, 0, 2] = -(k0 * err[:, 0])                                             # env.py:225-226, 254
        rg_lat[:, 0, 3] = -(k2 * err[:, 2])
        MDP_state = mdp.t()
        reward = torch.stack([reward_lon, reward_lat], dim=1).reshape(B, 2, 1, 1)
        info = {"nans": bool(torch.isnan(st).any()), "s": MDP_state, "yref": ref,
                "action_commanded": surf.t(), "action_effective": eff.t(), "rates": st[:, 6:9], "t": self.t, "x_full": st,
                "x": [st[:, [4, 7, 1]].unsqueeze(-1), st[:, [6, 5, 0, 2]].unsqueeze(-1)], "e": err,
                "RSE": [torch.sqrt(e_th * e_th), torch.sqrt(err[:, 0] ** 2 + err[:, 2] ** 2)],
                "reward_grad": [rg_lon, rg_lat]}
        return MDP_state, reward, None, False, info

    def render(self, mode="human"):
        pass

    def close(self):
        self.initialized = False
