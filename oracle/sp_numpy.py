"""TEST INFRASTRUCTURE ONLY -- structure-faithful numpy restatement of the reference's
short-period IDHP loop, one agent per object, shaped like the reference's own classes.

What it follows (wingos80/RL4AFCS):
  * ``ShortPeriodPlant``  envs/linear/env.py:7-264 (restated; the VERBATIM class can be
    plugged in instead -- ``tests/test_oracle_vs_reference.py`` does that in the build container)
  * ``IncrementalModel``  objects.py:439-549 (ditto)
  * ``TinyNet`` / ``CriticNet`` / ``ActorNet``  objects.py:39-281 with TensorFlow float32
    tensors replaced by numpy arrays of dtype ``tn`` and every TF op spelled out
  * ``IDHPspLoop.train``  objects.py:921-1004, with _step_networks :853-882,
    _update_networks :884-908, _adapt_check :783-841

Why it exists next to the C oracle (oracle/sp_oracle.c): it keeps the reference's object
graph, call order and *aliasing* (quirk Q3 arises here naturally because ``x = info['x']``
aliases the plant's array), so it pins the C restatement's control flow; and it is the
"reference-shaped" CPU loop timed by ``bench.py --impl reference``.

Arithmetic: ``@`` on the TensorFlow side is not observable here (no TF), so it is DEFINED as
an in-order FMA chain in ``tn`` -- the same definition the C oracle and the CUDA kernels use.
On the numpy side (plant, RLS) ``@`` is numpy's own; on x86-64 hosts with OpenBLAS it was
measured to equal the in-order FMA chain (re-checked by the tests).
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.fma.restype = ctypes.c_double
_libm.fma.argtypes = [ctypes.c_double] * 3
_libm.fmaf.restype = ctypes.c_float
_libm.fmaf.argtypes = [ctypes.c_float] * 3


def _fma(tn, a, b, c):
    if tn == np.float32:
        return np.float32(_libm.fmaf(float(a), float(b), float(c)))
    return np.float64(_libm.fma(float(a), float(b), float(c)))


def chain_matmul(tn, X, W):
    """(p,q)@(q,r) as in-order FMA chains in dtype ``tn``."""
    X = np.asarray(X, dtype=tn)
    W = np.asarray(W, dtype=tn)
    p, q = X.shape
    r = W.shape[1]
    out = np.zeros((p, r), dtype=tn)
    for i in range(p):
        for j in range(r):
            acc = tn(X[i, 0] * W[0, j])
            for t in range(1, q):
                acc = _fma(tn, X[i, t], W[t, j], acc)
            out[i, j] = acc
    return out


# ------------------------------------------------------------------ plant (env.py:7-264)
class ShortPeriodPlant:
    """Restated Ce500ShortPeriod: same attributes, reset/step 5-tuples and aliasing."""

    def __init__(self, env_config):
        from . import sp_c

        self._k = sp_c.ce500_coeffs()
        self.A = sp_c.ce500_A(self._k)
        self.B = sp_c.ce500_B(self._k)
        self.C = np.array([[1, 0], [0, 1]])
        self.D = np.array([[0], [0]])
        self.x0 = env_config["x0"]
        self.x = self.x0.copy()
        self.dt = env_config["dt"]
        self.t_end = env_config["t_end"]
        self.fault_time = env_config["fault_time"]
        self.fault_scenario = env_config["fault_scenario"]
        self.t = 0
        self.kappa = 28
        self.stepp = 0
        self.state_reference = env_config["reference"]["signal"][0]

    def reset(self, seed=None):
        from . import sp_c

        self.x = self.x0.copy()
        self.t = 0
        self.stepp = 0
        self._k = sp_c.ce500_coeffs()
        self.A = sp_c.ce500_A(self._k)
        self.B = sp_c.ce500_B(self._k)
        self.yref_hist = []
        info = {"yref": self.state_reference[0], "t": 0, "x": self.x, "e": 0,
                "reward_grad": self.kappa * np.array([[-2 * 0, 0]])}
        return self.x, 0, False, False, info

    def step(self, action):
        from . import sp_c

        action = np.deg2rad(action)
        y = self.C @ self.x + self.D * action
        ref = self.state_reference[self.stepp]
        err = (ref - y[0])[0]
        reward = -0.5 * self.kappa * (err * err)     # reference spells e**2 (libm pow, <= 1 ulp off)
        reward_grad = self.kappa * np.array([[-2 * err, 0]])
        xdot = self.A @ self.x + self.B * action
        self.x += xdot * self.dt                     # in place: the source of quirk Q3
        self.t += self.dt
        self.stepp += 1
        if self.stepp == int(self.fault_time / self.dt):
            self.A, self.B = sp_c.ce500_fault(self._k, self.A, self.B, self.fault_scenario)
        self.yref_hist.append(ref)
        info = {"yref": ref, "t": self.t, "x": self.x, "e": err, "reward_grad": reward_grad}
        return self.x, reward, self.t >= self.t_end, False, info


# ------------------------------------------------------------------ RLS (objects.py:439-549)
class IncrementalModel:
    def __init__(self, config):
        self.state_dim = config["state_dim"]
        self.action_dim = config["action_dim"]
        self.gamma = config["rls_gamma"]
        self.init_cov = config["rls_cov"]
        self._reset()
        self.eps_norm = 0.0
        self.epsilon = np.zeros((self.state_dim, 1))

    def _reset(self):
        p = self.state_dim + self.action_dim
        self.params = np.zeros((p, self.state_dim))
        self.Cov = self.init_cov * np.eye(p)

    @property
    def F(self):
        return np.array(self.params[: self.state_dim, :].T)

    @property
    def G(self):
        return np.array(self.params[self.state_dim:, :].T)

    def update(self, dx_t, da_t, dx_t1):
        X = np.concatenate((dx_t, da_t), axis=0)
        eps = dx_t1 - self.params.T @ X
        CX = self.Cov @ X
        K = CX / (self.gamma + X.T @ CX)
        self.params = self.params + K @ eps.T
        self.Cov = (self.Cov - K @ CX.T) / self.gamma
        self.epsilon = eps
        self.eps_norm = np.linalg.norm(eps)


# ------------------------------------------------------------------ networks (objects.py:39-281)
class TinyNet:
    """Bias-free 2-layer MLP, weights [W1 (in,h), W2 (h,out)], records xi / ai (objects.py:111-139)."""

    def __init__(self, tn, W1, W2, out_act, tanh_fn):
        self.tn = tn
        self.W = [np.array(W1, dtype=tn), np.array(W2, dtype=tn)]
        self.acts = ["tanh", out_act]
        self.tanh_fn = tanh_fn
        n_out = self.W[1].shape[1]
        n_par = self.W[0].size + self.W[1].size
        self.E = np.zeros((n_out, n_par))            # float64 storage (objects.py:109, Q15)
        self.xi = [None, None]
        self.ai = [None, None]
        self.gamma_lambda = 0.0
        self.eligibility = None

    def forward(self, s):
        tn = self.tn
        s = np.asarray(s, dtype=tn)
        for i, act in enumerate(self.acts):
            self.xi[i] = s.copy()
            s = chain_matmul(tn, s, self.W[i])
            if act == "tanh":
                s = np.asarray(self.tanh_fn(s), dtype=tn)
                self.ai[i] = tn(1) - s * s
            else:
                self.ai[i] = np.ones_like(s)
        return s

    def _apply_trace(self, grad_rows):
        """grad_rows: list of (row, slice, values) -- the structurally non-zero Jacobian slots."""
        if self.eligibility is None:
            for r, sl, v in grad_rows:
                self.E[r, sl] = v
        elif self.eligibility == "accumulating":
            self.E *= self.gamma_lambda
            for r, sl, v in grad_rows:
                self.E[r, sl] += v
        elif self.eligibility == "replacing":
            g = np.zeros_like(self.E)
            for r, sl, v in grad_rows:
                g[r, sl] = v
            if _chain_norm(g) > _chain_norm(self.E):
                self.E = g
            else:
                self.E *= self.gamma_lambda

    def soft_update(self, source_weights, tau):
        tn = self.tn
        for i, src in enumerate(source_weights):
            self.W[i] = tn(1.0 - tau) * self.W[i] + tn(tau) * src


def _chain_norm(M):
    acc = np.float64(0.0)
    for v in M.ravel():
        acc = _fma(np.float64, v, v, acc)
    return np.sqrt(acc)


class CriticNet(TinyNet):                               # objects.py:142-215
    def __call__(self, s):
        out = self.forward(s)
        W2T = self.W[1].T
        h = self.xi[1].ravel()
        z = self.xi[0]
        self._apply_trace([
            (0, slice(0, 4), h), (0, slice(8, 12), (W2T[0, :] * self.ai[0] * z).ravel()),
            (1, slice(4, 8), h), (1, slice(8, 12), (W2T[1, :] * self.ai[0] * z).ravel()),
        ])
        return out

    def get_weight_update(self, td):
        g = chain_matmul(self.tn, td, self.E.astype(self.tn))
        W2u = g[0, 0:8].reshape(2, 4).T
        W1u = g[0, 8:12].reshape(4, 1).T
        return [W1u, W2u]


class ActorNet(TinyNet):                                # objects.py:217-281
    def __call__(self, s):
        out = self.forward(s)
        tn = self.tn
        W2T = self.W[1].T
        g2 = chain_matmul(tn, self.ai[1], self.xi[1]).ravel()
        g1 = chain_matmul(tn, (chain_matmul(tn, self.ai[1], W2T) * self.ai[0]).T, self.xi[0]).ravel()
        self._apply_trace([(0, slice(0, 4), g2), (0, slice(4, 8), g1)])
        return out

    def get_weight_update(self, loss):
        g = chain_matmul(self.tn, loss, self.E.astype(self.tn))
        return [g[0, 4:8].reshape(1, 4), g[0, 0:4].reshape(4, 1)]

    def input_gradient(self):
        """d a / d z in TF's reverse-mode order (TanhGrad, MatMul grad, TanhGrad, MatMul grad)."""
        tn = self.tn
        g_o = tn(1) * self.ai[1]                                     # (1,1)
        g_h = chain_matmul(tn, g_o, self.W[1].T) * self.ai[0]         # (1,4)
        return chain_matmul(tn, g_h, self.W[0].T)                     # (1,1)


# ------------------------------------------------------------------ agent (objects.py:551-1004)
class IDHPspLoop:
    def __init__(self, env, config, weights, *, tn=np.float32, tanh_fn=np.tanh, rls=None,
                 numpy1_promotion=True):
        self.tn = tn
        self.gamma = config["gamma"]
        self.tau = config["tau"]
        self.ms = config["multistep"] > 0
        self.warmup_time = config["warmup_time"]
        self.error_thresh = config["error_thresh"]
        self.changed = False
        self.cooldown_1 = 0
        self.cooldown_timeit = int(config["cooldown_time"] / env.dt)
        self.env = env
        self.env.kappa = config["kappa"]
        self.numpy1_promotion = numpy1_promotion
        self.actor = ActorNet(tn, weights["W1a"].reshape(1, 4), weights["W2a"].reshape(4, 1), "tanh", tanh_fn)
        self.critic = CriticNet(tn, weights["W1c"].reshape(1, 4), weights["W2c"].reshape(4, 2), "linear", tanh_fn)
        self.target_critic = CriticNet(tn, weights["W1c"].reshape(1, 4), weights["W2c"].reshape(4, 2), "linear", tanh_fn)
        self.actor.eligibility = config["actor_config"]["elig"]
        self.critic.eligibility = config["critic_config"]["elig"]
        self.target_critic.eligibility = config["critic_config"]["elig"]
        self.lambda_h, self.lambda_l = config["lambda_h"], config["lambda_l"]
        for net in (self.actor, self.critic, self.target_critic):
            net.gamma_lambda = self.gamma * self.lambda_h
        self.eta_a_h = config["actor_config"]["eta_h"]
        self.eta_c_h = config["critic_config"]["eta_h"]
        self.eta_a_l = config["actor_config"]["eta_l"]
        self.eta_c_l = config["critic_config"]["eta_l"]
        self.model = IncrementalModel(config["rls_config"]) if rls is None else rls
        self.n, self.m = self.model.state_dim, self.model.action_dim

    # objects.py:783-841
    def _adapt_check(self, step, info):
        tn = self.tn
        thr = np.deg2rad(self.error_thresh)
        c1 = step < int(self.warmup_time / self.env.dt)
        c2 = np.abs(info["e"]) > thr
        c3 = self.model.eps_norm > 5e-5
        if self.cooldown_1 > 0:
            self.cooldown_1 -= 1
        if c1 or c2:
            eta_a, eta_c = tn(self.eta_a_h), tn(self.eta_c_h)
            lg = self.lambda_h * self.gamma
        else:
            eta_a, eta_c = tn(self.eta_a_l), tn(self.eta_c_l)
            lg = self.lambda_l * self.gamma
        if isinstance(self.eta_a, float) and not self.numpy1_promotion:
            differs = eta_a != tn(self.eta_a)                     # NEP 50: python float is weak
        else:
            differs = float(eta_a) != float(self.eta_a)           # numpy 1.x value-based promotion (Q7)
        if differs and self.cooldown_1 == 0:
            self.eta_a, self.eta_c = eta_a, eta_c
            self.lr_a, self.lr_c = tn(eta_a), tn(eta_c)
            self.actor.gamma_lambda = lg
            self.critic.gamma_lambda = lg
            self.cooldown_1 = self.cooldown_timeit
        if (c3 and not self.changed) and not c1:
            self.model._reset()
            self.changed = True

    def train(self, n_steps=None):
        tn = self.tn
        env = self.env
        self.target_critic.soft_update(self.critic.W, tau=1)
        dt = env.dt
        steps = int(env.t_end / dt) if n_steps is None else n_steps
        s, c, done, _, info = env.reset(seed=0)
        nn_in = np.array([[info["e"]]], dtype=tn)
        a = self.actor(nn_in).copy()
        x = info["x"]                                            # alias of the plant's array (Q3)
        a_prev = x_prev = None
        c_grad_prev = dx1dx0_prev = None
        self.eta_a, self.eta_c = self.eta_a_h, self.eta_c_h      # python floats until the first switch
        self.lr_a, self.lr_c = tn(self.eta_a), tn(self.eta_c)    # SGD learning-rate variables
        self.critic(nn_in)                                       # _asserts (objects.py:778)
        self.actor.W[1] = -self.actor.W[1]                       # _invert_controller (objects.py:850)
        gam = tn(self.gamma)
        keys = ("x", "a", "c", "ref", "a_w1", "a_w2", "c_w1", "c_w2", "params", "cov", "eps_norm", "a_e", "c_e")
        log = {k: [] for k in keys}
        for step in range(steps):
            s_next, c, done, _, info = env.step(tn(20) * a)
            x_next = info["x"].copy()
            e = info["e"].copy()
            c_grad = info["reward_grad"].copy()
            # _step_networks (objects.py:853-882): all three nets see [[e]] (Q1)
            z = np.array([[e]], dtype=tn)
            lam = self.critic(z)
            lam_t = self.target_critic(z)
            a_next = self.actor(z).copy()
            F, G = self.model.F, self.model.G
            dadx = self.actor.input_gradient()
            dx1dx0 = F.astype(tn) + chain_matmul(tn, G.astype(tn), dadx)   # (2,2)+(2,1) broadcast (Q2)
            if step > 0:
                cg = c_grad.astype(tn)
                if self.ms:
                    t1 = chain_matmul(tn, (self.gamma * c_grad).astype(tn), dx1dx0_prev)
                    t2 = chain_matmul(tn, chain_matmul(tn, tn(self.gamma ** 2) * lam_t, dx1dx0_prev), dx1dx0)
                    td = lam - c_grad_prev.astype(tn) - t1 - t2
                else:
                    td = lam - cg - chain_matmul(tn, gam * lam_t, dx1dx0)
                upd = self.critic.get_weight_update(td)
                for i in range(2):
                    self.critic.W[i] = self.critic.W[i] - self.lr_c * upd[i]
                self.target_critic.soft_update(self.critic.W, tau=self.tau)
                loss_grad = chain_matmul(tn, cg + gam * lam_t, G.astype(tn))
                upd = self.actor.get_weight_update(loss_grad)
                for i in range(2):
                    self.actor.W[i] = self.actor.W[i] - self.lr_a * upd[i]
                dx0 = x - x_prev
                da0 = a - a_prev
                dx1 = x_next - x
                self.model.update(dx0, da0, dx1)
                self._adapt_check(step, info)
            a_prev, x_prev = a, x
            a, x = a_next, x_next
            c_grad_prev, dx1dx0_prev = c_grad, dx1dx0
            log["x"].append(np.array(x).ravel().copy())
            log["a"].append(float(a[0, 0]))
            log["c"].append(float(c))
            log["ref"].append(float(env.yref_hist[step]))
            log["a_w1"].append(self.actor.W[0].ravel().astype(np.float64))
            log["a_w2"].append(self.actor.W[1].ravel().astype(np.float64))
            log["c_w1"].append(self.critic.W[0].ravel().astype(np.float64))
            log["c_w2"].append(self.critic.W[1].ravel().astype(np.float64))
            log["params"].append(self.model.params.ravel().copy())
            log["cov"].append(self.model.Cov.ravel().copy())
            log["eps_norm"].append(float(getattr(self.model, "eps_norm", 0.0)))
            log["a_e"].append(self.actor.E.ravel().copy())
            log["c_e"].append(self.critic.E.ravel().copy())
            if np.isnan(c):
                break
        return {k: np.array(v) for k, v in log.items()}
