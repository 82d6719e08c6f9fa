"""Batched, GPU-resident agent objects -- drop-in for objects.py of wingos80/RL4AFCS on the
short-period path: ``RLS`` (objects.py:439-549), ``Actor`` / ``Critic`` (objects.py:142-281) and
``IDHPsp`` (objects.py:551-1004).  Same class, method and attribute names; every array gains a
leading batch axis.  All arithmetic runs in the CUDA kernels behind include/rl4afcs_b200.h; the
objects are views over the SoA state planes of ``sp_engine.SpEngine``.

``IDHPsp.train()`` is the hot path: ONE fused persistent kernel launch runs the whole episode
(env step + critic / target / actor + RLS + adaptation + statistics) for every agent.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, sp_engine
from ._lib import LB, LF, SPE, SPN


# ------------------------------------------------------------------------------------------
class RLS:
    """Exponentially weighted recursive least squares on increments (objects.py:439-549)."""

    def __new__(cls, config, *, batch: int = 1, device="cuda", dtype: str = "mixed", _engine=None):
        if (config["state_dim"], config["action_dim"]) == (3, 1) and _engine is None:      # the nonlinear task's dimensions
            from . import nl_engine
            view = _RLSView(nl_engine.NlEngine(batch, policy=dtype, device=device), config)
            view._reset()
            return view
        return super().__new__(cls)

    def __init__(self, config, *, batch: int = 1, device="cuda", dtype: str = "mixed", _engine=None) -> None:
        self.state_dim = config["state_dim"]
        self.action_dim = config["action_dim"]
        assert (self.state_dim, self.action_dim) == (2, 1), "RLS supports the reference's two shapes: (2, 1) and (3, 1)"
        self.gamma = config["rls_gamma"]
        self.init_cov = config["rls_cov"]
        self._eng = _engine if _engine is not None else sp_engine.SpEngine(batch, policy=dtype, device=device)
        self.batch = self._eng.n
        self._eng.set_hp("RLS_GAMMA", self.gamma)
        self._eng.set_hp("RLS_COV0", self.init_cov)
        if _engine is None:
            self._reset()
        self.eps_norm_hist = [0, 0]
        self.eps_hist = [np.zeros(self.state_dim)]
        self.covs = []

    def _reset(self):
        """Theta <- 0, Cov <- init_cov * I (objects.py:477-482)."""
        e = self._eng
        e.env_field("THETA", 6).zero_()
        cov = e.env_field("COV", 9)
        cov.zero_()
        c0 = torch.as_tensor(np.asarray(self.init_cov, dtype=np.float64)).to(e.device).to(e.te)
        for d in (0, 4, 8):
            cov[d] = c0

    @property
    def params(self) -> torch.Tensor:
        """(B, 3, 2): [F.T; G.T] (objects.py:459-461)."""
        return self._eng.env_field("THETA", 6).t().reshape(self.batch, 3, 2)

    @property
    def Cov(self) -> torch.Tensor:
        return self._eng.env_field("COV", 9).t().reshape(self.batch, 3, 3)

    @property
    def F(self) -> torch.Tensor:
        return self.params[:, : self.state_dim, :].transpose(1, 2).clone()

    @property
    def G(self) -> torch.Tensor:
        return self.params[:, self.state_dim:, :].transpose(1, 2).clone()

    @property
    def epsilon(self) -> torch.Tensor:
        return self._eng.env_field("EPS", 2).t().unsqueeze(-1)

    @property
    def eps_norm(self) -> torch.Tensor:
        return self._eng.env_field("EPS_NORM")[0]

    def update(self, dx_t, da_t, dx_t1):
        """dx_t (B,2,1), da_t (B,1,1), dx_t1 (B,2,1) -> updates params / Cov in place (objects.py:492-543)."""
        e = self._eng
        te = e.te

        def plane(v, w):
            return torch.as_tensor(v, device=e.device).to(te).reshape(self.batch, w).t().contiguous()

        dx0, da0, dx1 = plane(dx_t, 2), plane(da_t, 1), plane(dx_t1, 2)
        e.set_hp("RLS_GAMMA", self.gamma)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_sp_rls_update(e.policy_id, ctypes.byref(e.params), e.env_field("THETA", 6).data_ptr(),
                                         e.env_field("COV", 9).data_ptr(), dx0.data_ptr(), da0.data_ptr(), dx1.data_ptr(),
                                         e.env_field("EPS", 2).data_ptr(), e.env_field("EPS_NORM").data_ptr(),
                                         e.stride, self.batch, e._stream())
            _lib.check(rc, "rl4_sp_rls_update")
        self.eps_hist.append(self.epsilon.clone())
        self.eps_norm_hist.append(self.eps_norm.clone())
        self.covs.append(self.Cov.clone())


# ------------------------------------------------------------------------------------------
class _Net:
    """Common part of Actor / Critic: bias-free in-h-out MLP whose weights live in the net plane
    (objects.py:39-140).  ``trainable_weights`` = [W1 (B,in,h), W2 (B,h,out)]."""

    _w1, _w2, _n_out = None, None, None

    def __init__(self, in_dim, layers, identity_init=False, std_init=0.01, seed=1, eligibility=None, *,
                 batch: int = 1, device="cuda", dtype: str = "mixed", _engine=None, _fields=None):
        sizes = list(layers.keys())
        acts = list(layers.values())
        assert in_dim == 1 and sizes[0] == 4 and acts[0] == "tanh", "short-period nets are 1-4-k (objects.py:160,235)"
        assert sizes[1] == self._n_out
        self.input_dim = in_dim
        self.layers_dict = layers
        self.n_layers = len(layers)
        self.eligibility = eligibility
        self.gamma_lambda = 0.0
        self._eng = _engine if _engine is not None else sp_engine.SpEngine(batch, policy=dtype, device=device)
        self.batch = self._eng.n
        if _fields is not None:
            self._w1, self._w2 = _fields
        if _engine is None and not identity_init:
            w = sp_engine.truncated_normal_weights(self.batch, seed, std_init, self._eng.device)
            key1, key2 = ("W1a", "W2a") if self._n_out == 1 else ("W1c", "W2c")
            self.set_weights([w[key1].reshape(self.batch, 1, 4), w[key2].reshape(self.batch, 4, self._n_out)])
        self.xi = [None, None]
        self.ai = [None, None]

    @property
    def trainable_weights(self):
        e = self._eng
        W1 = e.net_field(self._w1, 4).t().reshape(self.batch, 1, 4)
        W2 = e.net_field(self._w2, 4 * self._n_out).t().reshape(self.batch, 4, self._n_out)
        return [W1, W2]

    def set_weights(self, weights):
        W1, W2 = self.trainable_weights
        W1.copy_(torch.as_tensor(weights[0], device=W1.device).to(W1.dtype).reshape(W1.shape))
        W2.copy_(torch.as_tensor(weights[1], device=W2.device).to(W2.dtype).reshape(W2.shape))

    def get_weights(self):
        return [w.clone() for w in self.trainable_weights]

    def soft_update(self, source_weights, tau):
        """target <- (1 - tau) * target + tau * source, three separately rounded float ops per weight
        (objects.py:207-215), computed by ``rl4_soft_update``."""
        _soft_update(self._eng, [(self._w1, 4), (self._w2, 4 * self._n_out)], source_weights, tau, self.batch)


class Critic(_Net):
    """1-4-2 critic estimating the value-function gradient lambda (objects.py:142-215)."""
    _w1, _w2, _n_out = "W1C", "W2C", 2

    @property
    def E(self) -> torch.Tensor:
        """(B, 2, 12) Jacobian trace in the reference's layout (objects.py:160-166)."""
        e = self._eng
        out = torch.zeros((self.batch, 2, 12), dtype=e.te, device=e.device)
        h = e.env_field("EC_H", 4).t()
        out[:, 0, 0:4] = h
        out[:, 1, 4:8] = h
        out[:, 0, 8:12] = e.env_field("EC_W1R0", 4).t()
        out[:, 1, 8:12] = e.env_field("EC_W1R1", 4).t()
        return out

    def __call__(self, s):
        e = self._eng
        z = torch.as_tensor(s, device=e.device).to(e.tn).reshape(self.batch).contiguous()
        out = torch.empty((2, e.stride), dtype=e.tn, device=e.device)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_sp_critic_forward(e.policy_id, z.data_ptr(), e.net_field(self._w1, 4).data_ptr(),
                                             e.net_field(self._w2, 8).data_ptr(), e.env_field("EC_H", 12).data_ptr(),
                                             out.data_ptr(), float(self.gamma_lambda), _lib.ELIG[self.eligibility],
                                             e.stride, self.batch, e._stream())
            _lib.check(rc, "rl4_sp_critic_forward")
        return out.t().reshape(self.batch, 1, 2)

    call = __call__

    def get_weight_update(self, td_error):
        """td_error (B,1,2) -> [W1_update (B,1,4), W2_update (B,4,2)] (objects.py:195-205)."""
        e = self._eng
        td = torch.as_tensor(td_error, device=e.device).to(e.tn).reshape(self.batch, 2).t().contiguous()
        out = torch.empty((12, e.stride), dtype=e.tn, device=e.device)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_sp_critic_weight_update(e.policy_id, td.data_ptr(), e.env_field("EC_H", 12).data_ptr(),
                                                   out.data_ptr(), e.stride, self.batch, e._stream())
            _lib.check(rc, "rl4_sp_critic_weight_update")
        return [out[0:4].t().reshape(self.batch, 1, 4), out[4:12].t().reshape(self.batch, 4, 2)]


class Actor(_Net):
    """1-4-1 actor (objects.py:217-281)."""
    _w1, _w2, _n_out = "W1A", "W2A", 1

    @property
    def E(self) -> torch.Tensor:
        return self._eng.env_field("EA", 8).t().reshape(self.batch, 1, 8)

    def __call__(self, s, return_input_gradient: bool = False):
        e = self._eng
        z = torch.as_tensor(s, device=e.device).to(e.tn).reshape(self.batch).contiguous()
        out = torch.empty(self.batch, dtype=e.tn, device=e.device)
        dadz = torch.empty(self.batch, dtype=e.tn, device=e.device)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_sp_actor_forward(e.policy_id, z.data_ptr(), e.net_field(self._w1, 4).data_ptr(),
                                            e.net_field(self._w2, 4).data_ptr(), e.env_field("EA", 8).data_ptr(),
                                            out.data_ptr(), dadz.data_ptr(), float(self.gamma_lambda),
                                            _lib.ELIG[self.eligibility], e.stride, self.batch, e._stream())
            _lib.check(rc, "rl4_sp_actor_forward")
        self.input_gradient = dadz.reshape(self.batch, 1, 1)      # tape.gradient(a, nn_in) (objects.py:876-878)
        a = out.reshape(self.batch, 1, 1)
        return (a, self.input_gradient) if return_input_gradient else a

    call = __call__

    def get_weight_update(self, loss):
        """loss (B,1,1) -> [W1_update (B,1,4), W2_update (B,4,1)] = loss * E (objects.py:261-271)."""
        g = _actor_weight_update(self._eng, loss, 8, self.batch)          # (8, B): rows = E's columns
        return [g[4:8].t().reshape(self.batch, 1, 4), g[0:4].t().reshape(self.batch, 4, 1)]


# ------------------------------------------------------------------------------------------
class IDHPsp:
    """Incremental dual heuristic programming on the short-period model, one agent per batch entry
    (objects.py:551-1004).

    ``IDHPsp(env, config, verbose=True, seed=1)`` as in the reference; ``env`` is a batched
    ``rl4afcs_b200.envs.linear.env.Ce500ShortPeriod`` and fixes batch size, device and dtype policy.
    Every numeric entry of ``config`` may be a per-agent array (hyper-parameter sweeps).  Extra
    keyword arguments: ``weights`` (dict W1a (B,4), W2a (B,4), W1c (B,4), W2c (B,8); default:
    TruncatedNormal(sigma) drawn on the device from ``seed``), ``log`` ('full' | 'basic' | None),
    ``log_agents`` (how many leading agents are logged), ``log_every``; ``numpy2`` selects how `_adapt_check`'s
    `eta_a != self.eta_a` (objects.py:819) compares a float32 with a python float: True (default) = NEP 50 (numpy >= 2), the
    behaviour OBSERVED by running the verbatim agent (every golden fixture); False = numpy 1.x value-based promotion (the
    comparison is True at k = 1 and a cooldown starts) -- derived from the promotion rules, never observed: unverified.
    """

    def __init__(self, env, config, verbose=True, seed=1, *, weights=None, log="full", log_agents=None,
                 log_every: int = 1, ref_amp=None, numpy2: bool = True) -> None:
        self.seed = seed
        env._engine.params.q7_numpy1 = 0 if numpy2 else 1   # Q7: float32-vs-python-float compare at k = 1 (numpy 1.x) or NEP 50
        self.gamma = config["gamma"]
        self.tau = config["tau"]
        self.ms = config["multistep"]
        self.warmup_time = config["warmup_time"]
        self.error_thresh = config["error_thresh"]
        self.changed = False
        self.cooldown_1 = 0
        self.env = env
        self.env.kappa = config["kappa"]
        self.config = config
        self._eng = env._engine
        self.batch = self._eng.n
        self.verbose = verbose
        self.cooldown_timeit = int(np.max(config["cooldown_time"]) / env.dt)
        sp_engine.apply_idhp_config(self._eng, config, dt=env.dt)
        if ref_amp is not None:
            self._eng.set_hp("REF_AMP", ref_amp)
        self._setup_networks(config, seed, weights)
        self.log_level = {None: _lib.LOG_NONE, "none": _lib.LOG_NONE, "basic": _lib.LOG_BASIC, "full": _lib.LOG_FULL}[log]
        self.log_agents = min(self.batch, 64) if log_agents is None else min(int(log_agents), self.batch)
        self.log_every = int(log_every)

    def _setup_networks(self, config, seed, weights):
        eng = self._eng
        kw = dict(_engine=eng)
        self.hidden_dim = list(config["actor_config"]["layers"].keys())[0]
        self.actor = Actor(config["in_dims"], config["actor_config"]["layers"], False, config["sigma"], seed,
                           _first(config["actor_config"]["elig"]), **kw)
        self.critic = Critic(config["in_dims"], config["critic_config"]["layers"], False, config["sigma"], seed,
                             _first(config["critic_config"]["elig"]), **kw)
        self.target_critic = Critic(config["in_dims"], config["critic_config"]["layers"], False, config["sigma"], seed,
                                    _first(config["critic_config"]["elig"]), _fields=("W1T", "W2T"), **kw)
        self.lambda_h, self.lambda_l = config["lambda_h"], config["lambda_l"]
        gl = float(np.ravel(self.gamma)[0] * np.ravel(self.lambda_h)[0])
        self.actor.gamma_lambda = self.critic.gamma_lambda = self.target_critic.gamma_lambda = gl
        self.eta_a_h = config["actor_config"]["eta_h"]
        self.eta_c_h = config["critic_config"]["eta_h"]
        self.eta_a_l = config["actor_config"]["eta_l"]
        self.eta_c_l = config["critic_config"]["eta_l"]
        self.model = RLS(config["rls_config"], _engine=eng)
        self.n, self.m = self.model.state_dim, self.model.action_dim
        if weights is None:
            weights = sp_engine.truncated_normal_weights(self.batch, seed, config["sigma"], eng.device)   # scalar or per-agent sigma
        self._init_weights = weights

    # ---- the hot path ---------------------------------------------------------------------
    def train(self, n_steps=None):
        """Runs the whole episode in one fused kernel launch (objects.py:921-1004)."""
        env, eng = self.env, self._eng
        steps = int(env.t_end / env.dt) if n_steps is None else int(n_steps)
        env.reset(seed=self.seed)
        w = self._init_weights
        eng.init(torch.as_tensor(env.x0), w["W1a"], w["W2a"], w["W1c"], w["W2c"])
        self._setup_optimizers()
        log = eng.run(steps, log_level=self.log_level, log_agents=self.log_agents, log_every=self.log_every)
        env.stepp = steps
        env.t = env.dt * steps
        env.yref_hist = list(env.state_reference[:steps])
        self._steps = steps
        self._store_logs(log, steps)
        return self

    def _setup_optimizers(self):
        self.eta_a = self.eta_a_h
        self.eta_c = self.eta_c_h

    def _store_logs(self, log, steps):
        """Fills the history attributes of IDHPsp._reset_logs/_log (objects.py:617-726); arrays are
        (logged agents, rows, ...) torch tensors on the device."""
        if log is None:
            return
        lg = log.permute(2, 0, 1)                     # (agents, rows, fields)
        rows = lg.shape[1]
        dt = self.env.dt
        self.t_hist = torch.arange(rows, dtype=torch.float64) * (dt * self.log_every)
        self.x_hist = lg[:, :, LB["X"]:LB["X"] + 2]
        self.a_hist = lg[:, :, LB["A"]:LB["A"] + 1]
        self.s_hist = self.x_hist.to(self._eng.tn).to(torch.float64)
        self.c_hist = lg[:, :, LB["C"]]
        self.ref_hist = lg[:, :, LB["REF"]]
        self.e_hist = lg[:, :, LB["E"]]
        if self.log_level == _lib.LOG_FULL:
            self.a_weights_hist1 = lg[:, :, LF["AW1"]:LF["AW1"] + 4]
            self.a_weights_hist2 = lg[:, :, LF["AW2"]:LF["AW2"] + 4]
            self.c_weights_hist1 = lg[:, :, LF["CW1"]:LF["CW1"] + 4]
            self.c_weights_hist2 = lg[:, :, LF["CW2"]:LF["CW2"] + 8]
            self.a_e_hist = lg[:, :, LF["AE"]:LF["AE"] + 8]
            ce = torch.zeros(lg.shape[0], rows, 24, dtype=lg.dtype, device=lg.device)
            ce[:, :, 0:4] = lg[:, :, LF["CE"]:LF["CE"] + 4]
            ce[:, :, 16:20] = lg[:, :, LF["CE"]:LF["CE"] + 4]
            ce[:, :, 8:12] = lg[:, :, LF["CE"] + 4:LF["CE"] + 8]
            ce[:, :, 20:24] = lg[:, :, LF["CE"] + 8:LF["CE"] + 12]
            self.c_e_hist = ce
            self.a_all_grad_hist = lg[:, :, LF["AGRAD"]:LF["AGRAD"] + 8]
            self.c_all_grad_hist = lg[:, :, LF["CGRAD"]:LF["CGRAD"] + 12]
            # objects.py:706-707: np.linalg.norm of the W1 gradient as the network dtype holds it (a float32 norm in the mix)
            tn = self._eng.tn
            self.a_grad_hist = torch.linalg.vector_norm(self.a_all_grad_hist[:, :, 0:4].to(tn), dim=-1).to(lg.dtype)
            self.c_grad_hist = torch.linalg.vector_norm(self.c_all_grad_hist[:, :, 0:4].to(tn), dim=-1).to(lg.dtype)
            self.params_hist = lg[:, :, LF["PARAMS"]:LF["PARAMS"] + 6]
            self.cov_hist = lg[:, :, LF["COV"]:LF["COV"] + 9]
            self.eps_norm_hist = lg[:, :, LF["EPS_NORM"]]
            self.eps_hist = lg[:, :, LF["EPS_ABS"]:LF["EPS_ABS"] + 2]

    # ---- episode statistics (functions.py:39-60) ------------------------------------------------
    def stats(self) -> dict:
        return self._eng.stats(self._steps)

    @property
    def diverged(self) -> torch.Tensor:
        return self._eng.stats(self._steps)["diverged"]


# ------------------------------------------------------------------------------------------
class _BigNetView:
    """Read / write view of a 4-10-k net of the nonlinear agent (Critic_big / Actor_big,
    objects.py:283-437): ``trainable_weights`` = [W1 (B,4,10), W2 (B,10,k)]."""

    def __init__(self, engine, w1, w2, n_out, eligibility):
        self._eng, self._w1, self._w2, self._n_out = engine, w1, w2, n_out
        self.eligibility = eligibility
        self.batch = engine.n

    @property
    def trainable_weights(self):
        e = self._eng
        return [e.net_field(self._w1, 40).t().reshape(self.batch, 4, 10),
                e.net_field(self._w2, 10 * self._n_out).t().reshape(self.batch, 10, self._n_out)]

    def get_weights(self):
        return [w.clone() for w in self.trainable_weights]

    @property
    def gamma_lambda(self):
        return self._eng.env_field("GL")[0]


class Critic_big(_BigNetView):
    """4-10-3 critic of the nonlinear task; its gradient comes from autodiff (objects.py:304-305,1365)."""

    def __call__(self, s):
        """s (B,4) or (B,1,4) -> lambda (B,1,3) (objects.py:294-339; no trace is formed for this net)."""
        e = self._eng
        sp = torch.as_tensor(s, device=e.device).to(e.tn).reshape(self.batch, 4).t().contiguous()
        out = torch.empty((3, e.stride), dtype=e.tn, device=e.device)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_nl_critic_forward(e.policy_id, sp.data_ptr(), e.net_field(self._w1, 40).data_ptr(),
                                             e.net_field(self._w2, 30).data_ptr(), out.data_ptr(), e.stride, self.batch, e._stream())
            _lib.check(rc, "rl4_nl_critic_forward")
        return out[:, : self.batch].t().reshape(self.batch, 1, 3)

    call = __call__

    def soft_update(self, source_weights, tau):
        """target <- (1 - tau) target + tau source (objects.py:353-361), separately rounded, by ``rl4_soft_update``."""
        _soft_update(self._eng, [(self._w1, 40), (self._w2, 10 * self._n_out)], source_weights, tau, self.batch)


class Actor_big(_BigNetView):
    """4-10-1 actor with the hand-written Jacobian trace E (1,50) (objects.py:385-392)."""

    @property
    def E(self) -> torch.Tensor:
        return self._eng.env_field("EA", 50).t().reshape(self.batch, 1, 50)

    def __call__(self, s, trace=True, return_input_gradient: bool = False):
        """s (B,4) -> a (B,1,1); with ``trace`` the Jacobian trace E is updated (objects.py:374-407)."""
        e = self._eng
        sp = torch.as_tensor(s, device=e.device).to(e.tn).reshape(self.batch, 4).t().contiguous()
        out = torch.empty(e.stride, dtype=e.tn, device=e.device)
        dads = torch.empty((4, e.stride), dtype=e.tn, device=e.device)
        gl = float(self.gamma_lambda[0]) if torch.is_tensor(self.gamma_lambda) else float(self.gamma_lambda)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_nl_actor_forward(e.policy_id, sp.data_ptr(), e.net_field(self._w1, 40).data_ptr(),
                                            e.net_field(self._w2, 10).data_ptr(), e.env_field("EA", 50).data_ptr(), out.data_ptr(),
                                            dads.data_ptr(), gl, _lib.ELIG[self.eligibility], 1 if trace else 0, e.stride,
                                            self.batch, e._stream())
            _lib.check(rc, "rl4_nl_actor_forward")
        self.input_gradient = dads[:, : self.batch].t().clone()                 # tape.gradient(a, s) (objects.py:1323)
        a = out[: self.batch].reshape(self.batch, 1, 1)
        return (a, self.input_gradient) if return_input_gradient else a

    call = __call__

    def get_weight_update(self, loss):
        """loss (B,1,1) -> [W1_update (B,4,10), W2_update (B,10,1)] = loss * E (objects.py:417-427)."""
        g = _actor_weight_update(self._eng, loss, 50, self.batch).t()     # (B, 50)
        return [g[:, 10:].reshape(self.batch, 10, 4).transpose(1, 2), g[:, 0:10].reshape(self.batch, 10, 1)]


class _RLSView:
    """RLS incremental model with the nonlinear task's dimensions (n = 3, m = 1; objects.py:439-549): params (B,4,3),
    Cov (B,4,4) are views of the agent's state planes, ``update`` / ``_reset`` are the step-level calls
    (``rl4_nl_rls_update``).  Created by ``IDHPnonlin`` (shares its engine) or by ``RLS(config)`` with state_dim = 3."""

    def __init__(self, engine, config):
        self._eng = engine
        self.state_dim, self.action_dim = config["state_dim"], config["action_dim"]
        assert (self.state_dim, self.action_dim) == (3, 1)
        self.gamma, self.init_cov = config["rls_gamma"], config["rls_cov"]
        self.batch = engine.n
        self.eps_norm_hist = [0, 0]                       # objects.py:472-474
        self.eps_hist = [np.zeros(self.state_dim)]
        self.covs = []

    def _reset(self):
        """Theta <- 0, Cov <- init_cov * I (objects.py:477-482)."""
        e = self._eng
        e.env_field("THETA", 12).zero_()
        cov = e.env_field("COV", 16)
        cov.zero_()
        c0 = torch.as_tensor(np.asarray(self.init_cov, dtype=np.float64)).to(e.device)
        for d in (0, 5, 10, 15):
            cov[d] = c0

    @property
    def params(self):
        return self._eng.env_field("THETA", 12).t().reshape(self.batch, 4, 3)

    @property
    def Cov(self):
        return self._eng.env_field("COV", 16).t().reshape(self.batch, 4, 4)

    @property
    def F(self):
        return self.params[:, :3, :].transpose(1, 2).clone()

    @property
    def G(self):
        return self.params[:, 3:, :].transpose(1, 2).clone()

    @property
    def epsilon(self):
        return self._eng.env_field("EPS", 3).t().unsqueeze(-1)

    @property
    def eps_norm(self):
        return self._eng.env_field("EPS_NORM")[0]

    def update(self, dx_t, da_t, dx_t1):
        """dx_t (B,3,1), da_t (B,1,1), dx_t1 (B,3,1) -> updates params / Cov in place (objects.py:492-543)."""
        e = self._eng

        def plane(v, w):
            return torch.as_tensor(v, device=e.device).to(torch.float64).reshape(self.batch, w).t().contiguous()

        dx0, da0, dx1 = plane(dx_t, 3), plane(da_t, 1), plane(dx_t1, 3)
        e.set_hp("RLS_GAMMA", self.gamma)
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_nl_rls_update(ctypes.byref(e.params), e.env_field("THETA", 12).data_ptr(),
                                         e.env_field("COV", 16).data_ptr(), dx0.data_ptr(), da0.data_ptr(), dx1.data_ptr(),
                                         e.env_field("EPS", 3).data_ptr(), e.env_field("EPS_NORM").data_ptr(),
                                         e.stride, self.batch, e._stream())
            _lib.check(rc, "rl4_nl_rls_update")
        self.eps_hist.append(self.epsilon.clone())
        self.eps_norm_hist.append(self.eps_norm.clone())
        self.covs.append(self.Cov.clone())


class IDHPnonlin:
    """IDHP attitude tracking on the nonlinear aircraft, one agent per batch entry (objects.py:1006-1564).

    ``IDHPnonlin(env, config, verbose=True, seed=1)`` as in the reference; ``env`` is a batched
    ``rl4afcs_b200.envs.nonlinear.env.Ce500NonLinear``.  ``train()`` runs the 9000-step episode in fused kernel
    launches of ``chunk`` steps; the N(0,1) draw of objects.py:1375 is generated per chunk with torch's Philox
    generator (or supplied: ``noise=`` (steps, B) float32), the initial weights are TruncatedNormal(sigma) from
    ``seed`` or ``weights=`` (W1a (B,40), W2a (B,10), W1c (B,40), W2c (B,30)).

    ``numpy2``: the float32 / float64 promotion rules `_adapt_check` (objects.py:1235-1284) runs under -- True (default) = NEP 50,
    what the verbatim code was OBSERVED to do under numpy >= 2 (every golden fixture); False = numpy 1.x value-based casting,
    derived from the promotion rules and never observed (opt-in, unverified).

    ``log``: "full" = the reference's log dict (objects.py:1083-1176) for the first ``log_agents`` agents, every array
    with a leading agent axis; "compact" = x_full / a / e / reward / a_cmd only; "mc" = the per-step quantities
    MC_test_hparam keeps (functions.py:1040-1052), cheap enough for every agent; None = statistics only.
    """

    def __init__(self, env, config, verbose=True, seed=1, *, weights=None, log="full", log_agents=None,
                 chunk: int = 1000, numpy2: bool = True) -> None:
        from . import nl_engine  # noqa: F401

        assert log in ("full", "compact", "mc", None)
        self._log_mode = log

        self.seed = seed
        self.env = env
        self.config = config
        self.verbose = verbose
        self.gamma, self.tau = config["gamma"], config["tau"]
        self.ms = config["multistep"] > 0 if np.ndim(config["multistep"]) == 0 else None
        self.warmup_time = config["warmup_time"]
        self.lr_decay = config["lr_decay"]
        self._eng = eng = env._engine
        self.batch = eng.n
        self.env._set_weight_matrices(config["kappa"])                       # objects.py:1029
        self.cooldown_timeit = int(config["cooldown_time"] / env.dt)
        per = lambda v, f: f(v) if np.ndim(v) == 0 else np.asarray([f(x) for x in np.ravel(v)])   # noqa: E731
        eng.set_hp("GAMMA", config["gamma"]); eng.set_hp("GAMMA_SQ", per(config["gamma"], lambda g: g ** 2))
        eng.set_hp("TAU", config["tau"]); eng.set_hp("LR_DECAY", config["lr_decay"])
        eng.set_hp("LAMBDA_H", config["lambda_h"]); eng.set_hp("LAMBDA_L", config["lambda_l"])
        eng.set_hp("ETA_A_H", config["actor_config"]["eta_h"]); eng.set_hp("ETA_A_L", config["actor_config"]["eta_l"])
        eng.set_hp("ETA_C_H", config["critic_config"]["eta_h"]); eng.set_hp("ETA_C_L", config["critic_config"]["eta_l"])
        eng.set_hp("RLS_GAMMA", config["rls_config"]["rls_gamma"]); eng.set_hp("RLS_COV0", config["rls_config"]["rls_cov"])
        eng.set_hpi("MULTISTEP", per(config["multistep"], lambda m: 1 if m > 0 else 0))
        eng.set_hpi("NUMPY2", 1 if numpy2 else 0)        # _adapt_check promotion rules: numpy 1.x (reference era) or NEP 50
        eng.set_hpi("WARMUP_STEPS", per(config["warmup_time"], lambda t: int(t / env.dt)))          # objects.py:1222
        eng.set_hpi("COOLDOWN_STEPS", per(config["cooldown_time"], lambda t: int(t / env.dt)))      # objects.py:1025
        elig = config["actor_config"]["elig"]
        eng.set_hpi("ELIG_A", _lib.ELIG[elig] if not isinstance(elig, (list, tuple, np.ndarray))
                    else np.asarray([_lib.ELIG[e] for e in elig], dtype=np.int32))
        self.actor = Actor_big(eng, "W1A", "W2A", 1, _first(elig))
        self.critic = Critic_big(eng, "W1C", "W2C", 3, config["critic_config"]["elig"])
        self.target_critic = Critic_big(eng, "W1T", "W2T", 3, config["critic_config"]["elig"])
        self.model = _RLSView(eng, config["rls_config"])
        self.n, self.m = self.model.state_dim, self.model.action_dim
        self.s_dim = config["in_dims"]
        if weights is None:
            g = torch.Generator(device=eng.device); g.manual_seed(int(seed))
            sig = float(np.ravel(config["sigma"])[0])

            def draw(w):
                out = torch.randn((self.batch, w), generator=g, device=eng.device, dtype=torch.float32)
                bad = out.abs() > 2.0
                while bool(bad.any()):
                    out = torch.where(bad, torch.randn((self.batch, w), generator=g, device=eng.device, dtype=torch.float32), out)
                    bad = out.abs() > 2.0
                return (out * sig).double()
            weights = {"W1a": draw(40), "W2a": draw(10), "W1c": draw(40), "W2c": draw(30)}
        self._init_weights = weights
        self.log_agents = min(self.batch, 16) if log_agents is None else min(int(log_agents), self.batch)
        if log is None:
            self.log_agents = 0
        self.chunk = int(chunk)
        self.hidden_dim_a_lon = self.hidden_dim_c_lon = 10

    def train(self, n_steps=None, noise=None):
        env, eng = self.env, self._eng
        steps = int(env.t_end / env.dt) if n_steps is None else int(n_steps)
        w = self._init_weights
        eng.init(w["W1a"], w["W2a"], w["W1c"], w["W2c"])                     # env.reset + prologue (objects.py:1466-1488)
        env.stepp = 0
        g = torch.Generator(device=eng.device); g.manual_seed(int(self.seed) + 7919)
        logs, k = [], 0
        while k < steps:
            c = min(self.chunk, steps - k)
            nz = (torch.as_tensor(noise[k:k + c]) if noise is not None
                  else torch.randn((c, self.batch), generator=g, device=eng.device, dtype=torch.float32))
            lg = eng.run(c, nz, log_agents=self.log_agents, log_level={"full": 2, "mc": 3}.get(self._log_mode, 1))
            if lg is not None:
                logs.append(lg)
            k += c
        env.stepp = steps
        env.t = env.dt * steps
        self._steps = steps
        if logs:
            lg = torch.cat(logs, dim=0).permute(2, 0, 1)                     # (agents, rows, fields)
            self._store_logs(lg, steps)
        st = eng.stats()
        self.RSE = [st["rse"][:, 0], st["rse"][:, 1]]                        # objects.py:1488,1503-1504
        return self

    def _store_logs(self, lg, steps):
        """The log dict of objects.py:1083-1176 with a leading logged-agent axis: log[key] is (agents, N, width)."""
        env = self.env
        div = self._eng.int_field("DIVERGED_STEP")[: lg.shape[0]].to(torch.int64)
        kk = torch.arange(steps, dtype=torch.int64, device=lg.device).unsqueeze(0).expand(lg.shape[0], steps)
        last = torch.where(div >= 0, div - 1, torch.full_like(div, steps)).unsqueeze(1)
        t = (torch.minimum(kk, last).to(torch.float64) + 1) * env.dt      # objects.py:1495 (N12); frozen after divergence (:1173-1174)
        t = torch.where(last < 0, torch.zeros_like(t), t)                    # diverged at step 0: log['t'][-1] of a zero array
        if self._log_mode == "compact":
            L = _lib.NLL
            self.log = {"t": t.unsqueeze(-1), "x_full": lg[:, :, L["XFULL"]:L["XFULL"] + 12],
                        "a_cmd": lg[:, :, L["SURF"]:L["SURF"] + 1],
                        "x": lg[:, :, [L["XFULL"] + 4, L["XFULL"] + 7, L["XFULL"] + 1]],
                        "e": lg[:, :, L["E_THETA"]:L["E_THETA"] + 1], "a": lg[:, :, L["A"]:L["A"] + 1],
                        "reward": lg[:, :, L["REWARD"]]}
            return
        if self._log_mode == "mc":                                           # functions.py:1040-1052, radians / SI
            M = _lib.NLM
            self.log = {"t": t.unsqueeze(-1), **{k.lower(): lg[:, :, v] for k, v in M.items() if k != "COUNT"}}
            return
        F = _lib.NLF_FIELDS

        def col(name):
            o, w = F[name]
            return lg[:, :, o:o + w]
        nan_like = lambda w: torch.where(torch.isnan(col("ETA_A")), col("ETA_A"), torch.zeros_like(col("ETA_A"))).expand(-1, -1, w)  # noqa: E731
        self.log = {
            "eta_a": col("ETA_A"), "t": t.unsqueeze(-1), "x_full": col("XFULL"), "RSE": col("RSE"), "x": col("X"),
            "a_cmd": col("A_CMD"), "a_eff": col("A_EFF"),
            "s": col("S").expand(-1, -1, self.s_dim),          # objects.py:1133: info['s'][0] broadcast over the row
            "yref": col("YREF").expand(-1, -1, self.s_dim),    # objects.py:1134
            "e": col("E"),
            "a_weights1": col("A_W1"), "a_weights2": col("A_W2"), "c_weights1": col("C_W1"), "c_weights2": col("C_W2"),
            "a_grad": col("A_GRAD"), "c_grad": col("C_GRAD"),
            "a_elig": nan_like(50), "c_elig": nan_like(70),    # allocated, never written by the reference (:1111,1113)
            "rls_params": col("RLS_PARAMS"), "rls_cov": col("RLS_COV"), "rls_eps_hist": col("RLS_EPS"),
            "rls_eps_norm": col("RLS_EPS_NORM"),
            "a": col("A"), "reward": col("REWARD")[:, :, 0],   # not in the reference's dict
        }

    def stats(self) -> dict:
        return self._eng.stats()


def _soft_update(e, fields, source_weights, tau, batch):
    """Polyak update of the weight planes ``fields`` = [(net-plane field, rows), ...] from reference-shaped source arrays
    [W1 (B,in,h), W2 (B,h,out)] through the C-ABI (objects.py:207-215, 353-361)."""
    for (name, rows), src in zip(fields, source_weights):
        sp = torch.as_tensor(src, device=e.device).to(e.tn).reshape(batch, rows).t().contiguous()     # (rows, B) plane
        assert e.stride == batch
        with torch.cuda.device(e.device):
            rc = e.lib.rl4_soft_update(e.policy_id, e.net_field(name, rows).data_ptr(), sp.data_ptr(), float(tau), rows,
                                       e.stride, batch, e._stream())
            _lib.check(rc, "rl4_soft_update")


def _actor_weight_update(e, loss, rows, batch):
    """loss * E (E stored in the trace dtype, cast to the tensor dtype: objects.py:261-271, 417-427) -> (rows, B) planes."""
    lp = torch.as_tensor(loss, device=e.device).to(e.tn).reshape(batch).contiguous()
    out = torch.empty((rows, e.stride), dtype=e.tn, device=e.device)
    with torch.cuda.device(e.device):
        rc = e.lib.rl4_actor_weight_update(e.policy_id, lp.data_ptr(), e.env_field("EA", rows).data_ptr(), out.data_ptr(), rows,
                                           e.stride, batch, e._stream())
        _lib.check(rc, "rl4_actor_weight_update")
    return out[:, :batch]


def _first(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return v[0]
    return v
