"""The reference-facing Python objects (same names as envs/linear/env.py and objects.py of the
reference, batched) against golden vectors of the VERBATIM reference classes and the oracle."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FAULTS = {"None": None, "invert_elevator": "invert_elevator", "damp_elevator": "damp_elevator", "shift_cg": "shift_cg"}
ELIG = {"None": None, "accumulating": "accumulating", "replacing": "replacing"}


def _ref_signal(oracle):
    base, amp = oracle.default_reference()
    return amp * base


def _env_config(oracle, x0, fault, tracked="alpha"):
    return {"state_dim": 2, "action_dim": 1, "x0": x0, "dt": 0.02, "t_end": 60, "fault_time": 20,
            "fault_scenario": fault, "reference": {"tracked_state": [tracked], "signal": [_ref_signal(oracle)]}}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "sp_env_*.npz"))))
def test_env_step_equals_verbatim_reference(oracle, path):
    from rl4afcs_b200.envs.linear.env import Ce500ShortPeriod

    g = np.load(path)
    B = 5
    env = Ce500ShortPeriod(_env_config(oracle, g["x0"].reshape(2, 1), FAULTS[str(g["fault"])]), batch=B, dtype="mixed")
    env.kappa = float(g["kappa"])
    obs, r, term, trunc, info = env.reset(seed=0)
    assert info["x"].data_ptr() == env.x.data_ptr() and obs.shape == (B, 2, 1)          # alias, like the reference (Q3)
    act = torch.as_tensor(g["action"])
    for k in range(600 if "none" in path else 1500):
        a = act[k].reshape(1, 1, 1).expand(B, 1, 1).float()
        obs, reward, term, trunc, info = env.step(torch.tensor(20.0, dtype=torch.float32) * a)
        x = obs.cpu().numpy()
        assert np.array_equal(x[:, :, 0], np.broadcast_to(g["x"][k], (B, 2))), k
        assert np.array_equal(info["e"].cpu().numpy(), np.full(B, g["e"][k])), k
        assert np.array_equal(info["reward_grad"].cpu().numpy()[:, 0, :], np.broadcast_to(g["reward_grad"][k], (B, 2))), k
        ulp = np.abs(reward.cpu().numpy() - g["reward"][k]) / np.spacing(abs(g["reward"][k]) + 1e-300)
        assert (ulp <= 1.0).all(), k                      # reference: e**2 via libm pow
    assert env.stepp == k + 1 and info["yref"] == _ref_signal(oracle)[k]
    if FAULTS[str(g["fault"])] is not None:
        assert np.array_equal(env.A, g["A"]) and np.array_equal(env.B, g["B"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "sp_rls_*.npz"))))
def test_rls_update_equals_verbatim_reference(path):
    from rl4afcs_b200.objects import RLS

    g = np.load(path)
    B = 3
    m = RLS({"state_dim": 2, "action_dim": 1, "rls_gamma": float(g["gamma"]), "rls_cov": 10 ** 6}, batch=B, dtype="mixed")
    assert m.params.shape == (B, 3, 2) and m.Cov.shape == (B, 3, 3)
    for k in range(len(g["da0"])):
        if k == int(g["reset_at"]):
            m._reset()
        ex = lambda v, w: torch.as_tensor(v).reshape(1, w, 1).expand(B, w, 1)   # noqa: E731
        m.update(ex(g["dx0"][k], 2), ex(g["da0"][k:k + 1], 1), ex(g["dx1"][k], 2))
        assert np.array_equal(m.params.cpu().numpy().reshape(B, 6), np.broadcast_to(g["theta"][k], (B, 6))), k
        assert np.array_equal(m.Cov.cpu().numpy().reshape(B, 9), np.broadcast_to(g["cov"][k], (B, 9))), k
        assert np.array_equal(m.epsilon.cpu().numpy().reshape(B, 2), np.broadcast_to(g["eps"][k], (B, 2))), k
        assert np.array_equal(m.eps_norm.cpu().numpy(), np.full(B, g["eps_norm"][k])), k
    F, G = m.F.cpu().numpy(), m.G.cpu().numpy()
    th = g["theta"][-1].reshape(3, 2)
    assert np.array_equal(F[0], th[:2].T) and np.array_equal(G[0], th[2:].T)


@pytest.mark.parametrize("policy", ["mixed", "fp64"])
@pytest.mark.parametrize("elig", [None, "accumulating", "replacing"])
def test_actor_critic_calls_equal_oracle(oracle, policy, elig):
    """Critic.__call__/Actor.__call__ (+ traces, da/dz), get_weight_update and soft_update vs the
    oracle's per-step log (same inputs)."""
    from rl4afcs_b200.objects import Actor, Critic

    n, steps = 33, 25
    ic = oracle.default_idhp_config()
    ic["actor_config"]["elig"] = ic["critic_config"]["elig"] = elig
    cfg = oracle.make_cfg(ic)
    base, amp = oracle.default_reference()
    rng = np.random.default_rng(5)
    x0 = np.deg2rad(rng.uniform(-2, 2, size=(n, 2)))
    w = oracle.init_weights(n, 6)
    st = oracle.init_states(policy, cfg, x0, w)
    actor = Actor(1, {4: "tanh", 1: "tanh"}, False, 0.1, 1, elig, batch=n, dtype=policy)
    critic = Critic(1, {4: "tanh", 2: "linear"}, False, 0.1, 1, elig, batch=n, dtype=policy)
    target = Critic(1, {4: "tanh", 2: "linear"}, False, 0.1, 1, elig, batch=n, dtype=policy)
    gl = float(cfg["gamma"][0] * cfg["lambda_h"][0])
    actor.gamma_lambda = critic.gamma_lambda = gl
    for k in range(steps):
        pre = st.copy()
        lg = oracle.run(policy, cfg, base, st, k, 1, tanh="t13", n_log=n)[:, 0]
        # load the pre-step weights / traces into the objects
        actor.set_weights([pre["W1a"].reshape(n, 1, 4), pre["W2a"].reshape(n, 4, 1)])
        critic.set_weights([pre["W1c"].reshape(n, 1, 4), pre["W2c"].reshape(n, 4, 2)])
        target.set_weights([pre["W1t"].reshape(n, 1, 4), pre["W2t"].reshape(n, 4, 2)])
        actor._eng.env_field("EA", 8).copy_(torch.as_tensor(pre["Ea"].T.copy()).to(actor._eng.te))
        ce = critic._eng.env_field("EC_H", 12)
        ce[0:4] = torch.as_tensor(pre["Ec"][:, 0:4].T.copy()).to(ce.dtype)
        ce[4:8] = torch.as_tensor(pre["Ec"][:, 8:12].T.copy()).to(ce.dtype)
        ce[8:12] = torch.as_tensor(pre["Ec"][:, 20:24].T.copy()).to(ce.dtype)
        z = torch.as_tensor(lg["e"]).reshape(n, 1, 1)
        lam = critic(z)
        a, dadz = actor(z, return_input_gradient=True)
        assert np.array_equal(lam.double().cpu().numpy().reshape(n, 2), lg["lam"]), k
        assert np.array_equal(a.double().cpu().numpy().ravel(), lg["a"]), k
        assert np.array_equal(dadz.double().cpu().numpy().ravel(), lg["dadz"]), k
        assert np.array_equal(actor.E.double().cpu().numpy().reshape(n, 8), lg["a_e"]), k
        assert np.array_equal(critic.E.double().cpu().numpy().reshape(n, 24), lg["c_e"]), k
        if k > 1:
            W1u, W2u = critic.get_weight_update(torch.as_tensor(lg["td"]).reshape(n, 1, 2))
            got = torch.cat([W1u.reshape(n, 4), W2u.reshape(n, 8)], dim=1).double().cpu().numpy()
            assert np.array_equal(got, lg["c_all_grad"]), k
            W1u, W2u = actor.get_weight_update(torch.as_tensor(lg["loss_grad"]).reshape(n, 1, 1))
            got = torch.cat([W1u.reshape(n, 4), W2u.reshape(n, 4)], dim=1).double().cpu().numpy()
            assert np.array_equal(got, lg["a_all_grad"]), k
            # SGD + Polyak as the reference does them (objects.py:892-895)
            lr = torch.as_tensor(pre["eta_c"]).to(W1u.dtype).cuda().reshape(n, 1, 1)
            cw = critic.trainable_weights
            g = critic.get_weight_update(torch.as_tensor(lg["td"]).reshape(n, 1, 2))
            cw[0].copy_(cw[0] - lr * g[0]); cw[1].copy_(cw[1] - lr * g[1])
            target.soft_update(critic.trainable_weights, tau=float(cfg["tau"][0]))
            assert np.array_equal(cw[0].double().cpu().numpy().reshape(n, 4), st["W1c"]), k
            assert np.array_equal(target.trainable_weights[1].double().cpu().numpy().reshape(n, 8), st["W2t"]), k


@pytest.mark.parametrize("name", ["default_x0zero", "default_x0rand", "shiftcg_acc_1step", "invert_replacing", "trackq_damp"])
def test_idhpsp_train_equals_reference_run(oracle, name):
    """IDHPsp(env, config).train() (BASELINE.json configs[0]: one agent, idhp_sp.py defaults) against golden runs of the
    VERBATIM reference agent (objects.py's IDHPsp on the TensorFlow stand-in, verbatim env + RLS; oracle/make_golden.py)."""
    from rl4afcs_b200.envs.linear.env import Ce500ShortPeriod
    from rl4afcs_b200.objects import IDHPsp

    g = np.load(os.path.join(GOLD, f"sp_loop_{name}.npz"))
    ic = oracle.default_idhp_config()
    ic["multistep"] = int(g["multistep"])
    ic["actor_config"]["elig"] = ELIG[str(g["elig_a"])]
    ic["critic_config"]["elig"] = ELIG[str(g["elig_c"])]
    B = 2
    env = Ce500ShortPeriod(_env_config(oracle, g["x0"].reshape(2, 1), FAULTS[str(g["fault"])], str(g["tracked"])), batch=B, dtype="mixed")
    w = {k: np.broadcast_to(g[f"w_{k}"], (B,) + g[f"w_{k}"].shape).copy() for k in ("W1a", "W2a", "W1c", "W2c")}
    idhp = IDHPsp(env, ic, verbose=0, seed=4, weights=w, log="full", log_agents=B)   # DEFAULT mode = the fixtures' (NEP 50, numpy >= 2)
    steps = int(g["steps"])
    idhp.train(steps)
    for b in range(B):
        assert np.array_equal(idhp.x_hist[b].cpu().numpy(), g["x"])
        assert np.array_equal(idhp.a_hist[b, :, 0].cpu().numpy(), g["a"])
        assert np.array_equal(idhp.ref_hist[b].cpu().numpy(), g["ref"])
        assert np.array_equal(idhp.a_weights_hist1[b].cpu().numpy(), g["a_w1"])
        assert np.array_equal(idhp.a_weights_hist2[b].cpu().numpy(), g["a_w2"])
        assert np.array_equal(idhp.c_weights_hist1[b].cpu().numpy(), g["c_w1"])
        assert np.array_equal(idhp.c_weights_hist2[b].cpu().numpy(), g["c_w2"])
        assert np.array_equal(idhp.a_e_hist[b].cpu().numpy(), g["a_e"]) and np.array_equal(idhp.c_e_hist[b].cpu().numpy(), g["c_e"])
        assert np.array_equal(idhp.a_all_grad_hist[b].cpu().numpy(), g["a_all_grad"])
        assert np.array_equal(idhp.c_all_grad_hist[b].cpu().numpy(), g["c_all_grad"])
        assert np.array_equal(idhp.params_hist[b, 2:].cpu().numpy(), g["params"][2:])
        assert np.array_equal(idhp.cov_hist[b, 2:].cpu().numpy(), g["cov"][2:])
        assert np.array_equal(idhp.eps_norm_hist[b, 2:].cpu().numpy(), g["eps_norm"][2:])
        c = idhp.c_hist[b].cpu().numpy()
        nz = np.abs(g["c"]) > 0
        assert (np.abs(c - g["c"])[nz] / np.spacing(np.abs(g["c"][nz])) <= 1.0).all()
    # object views after the run
    assert idhp.actor.trainable_weights[0].shape == (B, 1, 4) and idhp.critic.trainable_weights[1].shape == (B, 4, 2)
    assert np.array_equal(idhp.model.params[0].cpu().numpy().ravel(), g["params"][-1])
    s = idhp.stats()
    assert s["sum_c"].shape == (B,) and not bool(s["diverged"].any())
    assert env.stepp == steps and len(env.yref_hist) == steps


def test_nonlinear_env_and_agent_api(oracle):
    """Ce500NonLinear.reset/step and IDHPnonlin.train() through the reference-shaped objects
    (idhp_nonlin.py:107-157), checked against the oracle's free run with the same weights and noise."""
    from oracle import nl_c
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear
    from rl4afcs_b200.objects import IDHPnonlin
    from tests import _util_nl

    B, steps = 6, 300
    th = nl_c.theta_reference()
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    env_config = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01,
                  "t_end": 90, "total_steps": 9000, "fault_time": 60, "fault_scenario": "none",
                  "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    idhp_config = {"gamma": 0.6, "multistep": 0, "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": 0.95, "kappa": [1, 2, 1],
                   "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 4.0, "error_thresh": 1, "tau": 0.02, "in_dims": 4,
                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 35.0, "eta_l": 5.0, "elig": "accumulating"},
                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.4, "eta_l": 0.7, "elig": 1233},
                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}
    env = Ce500NonLinear(env_config, batch=B, dtype="mixed", plant="surrogate")
    s, r, term, trunc, info = env.reset()
    assert s.shape == (B, 4) and float(s.abs().max()) == 0.0 and info["x_full"].shape == (B, 12)
    assert abs(float(info["x_full"][0, 3]) - 90.0) < 1e-9 and abs(float(info["x_full"][0, 7]) - 0.0576) < 1e-9
    s, r, _, trunc, info = env.step(np.array([0.1, 0.0, 0.0]))
    assert s.shape == (B, 4) and r.shape == (B, 2, 1, 1) and trunc is False and set(info) >= {"nans", "s", "x_full", "x", "e", "RSE", "reward_grad"}
    assert float(info["action_commanded"][0, 0]) > 0.0           # actuator moved towards +1.5 deg

    w = nl_c.init_weights(B, 8)
    noise = np.random.default_rng(0).standard_normal((steps, B)).astype(np.float32)
    idhp = IDHPnonlin(env, idhp_config, seed=8, verbose=0, weights=w, log_agents=B, chunk=128)
    idhp.train(steps, noise=noise)
    cfg = nl_c.make_cfg()
    st = nl_c.init_states("mixed", cfg, w, B)
    olog = nl_c.run("mixed", cfg, th, noise, st, 0, steps, tanh="t13", n_log=B)
    assert idhp.log["x_full"].shape == (B, steps, 12)
    assert _util_nl.max_rel(idhp.log["x_full"].cpu().numpy()[:, :150, :9], olog["x_full"][:, :150, :9], 1e-2) < 1e-6
    assert idhp.actor.trainable_weights[0].shape == (B, 4, 10) and idhp.critic.trainable_weights[1].shape == (B, 10, 3)
    assert idhp.actor.E.shape == (B, 1, 50) and idhp.model.params.shape == (B, 4, 3) and idhp.model.Cov.shape == (B, 4, 4)
    assert idhp.RSE[0].shape == (B,) and bool((idhp.RSE[0] > 0).all())


def test_mc_run_front_end_and_statistics(oracle):
    """MC_run (functions.py:62-232) on the batch axis: 3 configs x 8 seeds; statistics against the oracle's own
    accumulators and against numpy restatements of utils.get_PSD / get_convergence_time."""
    from rl4afcs_b200 import functions as F

    base, amp = oracle.default_reference()
    env_config = {"state_dim": 2, "action_dim": 1, "x0": np.zeros((2, 1)), "dt": 0.02, "t_end": 60, "fault_time": 20,
                  "fault_scenario": None, "reference": {"tracked_state": ["alpha"], "signal": [amp * base]}}
    configs = {"lambda_hs": [0.6, 0.6, 0.6], "lambda_ls": [0.2, 0.2, 0.2], "kappas": [800, 800, 1140], "cooldown_times": None,
               "sigmas": None, "warmup_times": None, "elig_a": [None, "accumulating", "replacing"],
               "lr_a_hs": [2.0, 2.0, 3.55], "lr_a_ls": [0.02, 0.02, 0.054], "lr_c_hs": [0.2, 0.2, 0.338], "lr_c_ls": None,
               "multistep": [0, 2, 2]}
    seeds = 8
    metrics, idhp = F.MC_run(3, configs, env_config, seeds, dtype="mixed", log_agents=24)
    assert len(metrics) == 3 and all(set(m) >= {"diverged", "unsteady_convergence", "avg_c", "avg_t", "avg_PSD_err"} for m in metrics)
    out = F.MC_run_seed(idhp)
    # in-kernel statistics == the same quantities recomputed from the logged trajectories
    st = idhp.stats()
    ct = F.get_convergence_time(idhp.c_hist, torch.as_tensor(np.repeat([800, 800, 1140], seeds)).cuda()[:, None], 0.02)
    assert torch.equal(ct, st["converged_time"])
    assert torch.equal(out["converged_time"], st["converged_time"])
    assert torch.allclose(idhp.c_hist.sum(dim=-1) / torch.as_tensor(np.repeat([800., 800., 1140.], seeds)).cuda(), st["sum_c"], rtol=1e-12)
    # get_PSD against numpy's formula (utils.py:188-236)
    x = idhp.x_hist[:, :, 0].cpu().numpy()
    spec, omega = F.get_PSD(60, 0.02, idhp.x_hist[:, :, 0])
    ref = np.abs(np.fft.fft(x, axis=-1) * np.conj(np.fft.fft(x, axis=-1)) / 60)[:, :1500]
    assert np.allclose(spec.cpu().numpy(), ref, rtol=1e-9, atol=1e-18) and omega.shape[0] == 1500
    # and the whole batch against the oracle with the same per-agent configs
    n = 3 * seeds
    assert idhp.x_hist.shape == (n, 3000, 2)


def _nl_env_config(th, **over):
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    cfg = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01,
           "t_end": 90, "total_steps": 9000, "fault_time": 60, "fault_scenario": "none",
           "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    cfg.update(over)
    return cfg


def test_idhpnonlin_full_log_dict(oracle):
    """The log dict of IDHPnonlin._reset_logs / _log (objects.py:1083-1176): same keys and widths, leading agent axis."""
    from oracle import nl_c
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear
    from rl4afcs_b200.objects import IDHPnonlin

    B, steps = 5, 260
    th = nl_c.theta_reference()
    idhp_config = {"gamma": 0.6, "multistep": 1, "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": 0.95, "kappa": [1, 2, 1],
                   "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 1.0, "error_thresh": 1, "tau": 0.02, "in_dims": 4,
                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 35.0, "eta_l": 5.0, "elig": "accumulating"},
                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.4, "eta_l": 0.7, "elig": 1233},
                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}
    env = Ce500NonLinear(_nl_env_config(th), batch=B, dtype="mixed", plant="surrogate")
    idhp = IDHPnonlin(env, idhp_config, seed=3, verbose=0, log_agents=B, chunk=100)
    idhp.train(steps)
    widths = {"eta_a": 1, "t": 1, "x_full": 12, "RSE": 2, "x": 3, "a_cmd": 1, "a_eff": 1, "s": 4, "yref": 4, "e": 1,
              "a_weights1": 40, "a_weights2": 10, "c_weights1": 40, "c_weights2": 30, "a_grad": 50, "a_elig": 50,
              "c_grad": 70, "c_elig": 70, "rls_params": 12, "rls_cov": 16, "rls_eps_hist": 3, "rls_eps_norm": 1}
    for k, w in widths.items():                                              # objects.py:1094-1119
        assert tuple(idhp.log[k].shape) == (B, steps, w), (k, idhp.log[k].shape)
    lg = {k: v.cpu().numpy() for k, v in idhp.log.items()}
    assert np.allclose(lg["t"][0, :, 0], 0.01 * (np.arange(steps) + 1))                       # N12
    assert np.array_equal(lg["x"], lg["x_full"][:, :, [4, 7, 1]])
    assert np.array_equal(lg["s"], np.repeat(lg["x_full"][:, :, 4:5], 4, axis=2))              # objects.py:1133 broadcast quirk
    assert np.array_equal(lg["yref"][0, :, 0], th[:steps]) and np.array_equal(lg["e"][:, :, 0], lg["x_full"][:, :, 7] - th[:steps])
    assert np.array_equal(lg["RSE"][:, :, 0], np.sqrt(lg["e"][:, :, 0] ** 2))
    assert np.allclose(lg["RSE"][:, :, 0].sum(axis=1), idhp.RSE[0].cpu().numpy(), rtol=1e-12)
    assert not lg["a_elig"].any() and not lg["c_elig"].any() and not lg["a_grad"][:, 0].any()
    assert lg["a_grad"][:, 5].any() and lg["c_grad"][:, 5].any()
    # the logged gradient is the SGD step actually taken: W(k) = W(k-1) - lr * grad(k), float32
    k = 50
    lr = np.float32(35.0)
    w_prev, w_now = lg["a_weights2"][:, k - 1].astype(np.float32), lg["a_weights2"][:, k].astype(np.float32)
    assert np.array_equal(w_now, w_prev - lr * lg["a_grad"][:, k, 40:].astype(np.float32))
    # learning-rate decay after the warm-up (objects.py:1235-1243): eta_a falls from eta_h towards eta_l
    assert lg["eta_a"][0, 50, 0] == 35.0 and 5.0 < lg["eta_a"][0, -1, 0] < 35.0
    # final log row == final state
    assert np.array_equal(lg["rls_params"][:, -1].reshape(B, 4, 3), idhp.model.params.cpu().numpy())
    assert np.array_equal(lg["a_weights1"][:, -1].reshape(B, 4, 10), idhp.actor.trainable_weights[0].double().cpu().numpy())


def test_mc_test_hparam_front_end(oracle):
    """MC_test_hparam (functions.py:931-1060): 3 algorithms x 4 repetitions in one batch, short episode."""
    from oracle import nl_c
    from rl4afcs_b200 import functions as F
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear

    N, reps, steps, split = 3, 4, 1200, 700
    th = nl_c.theta_reference()
    env = Ce500NonLinear(_nl_env_config(th, t_end=steps * 0.01, total_steps=steps), batch=N * reps, dtype="mixed", plant="surrogate")
    configs = {"etaah": [35.0] * N, "etaal": [5.0] * N, "etach": [1.4] * N, "etacl": [0.7] * N, "lambda_hs": [0.95] * N,
               "lambda_ls": [0.95] * N, "seeds": [0] * N, "ms": [0, 0, 1], "elig": [None, "accumulating", "replacing"]}
    out = F.MC_test_hparam(configs, "unused/", env, N, reps, save=0, show=0, flight_step=split)
    assert [o[0] for o in out] == ["idhp", "idhpat", "midhprt"]
    for algo, cfg, log in out:
        assert set(log) >= {"RSE", "e", "theta", "alpha", "q", "V", "h", "action_cmd", "action_eff", "n_z", "wa_norm",
                            "wc_norm", "Sm", "rls_eps"}                       # functions.py:1008-1021
        assert tuple(log["RSE"].shape) == (reps, 2) and tuple(log["theta"].shape) == (reps, steps) and tuple(log["Sm"].shape) == (reps, 1)
        e = torch.deg2rad(log["e"])
        assert torch.allclose(log["RSE"][:, 0], e[:, :split].abs().sum(dim=1), rtol=1e-9)      # :1036
        assert torch.allclose(log["RSE"][:, 1], e[:, split:].abs().sum(dim=1), rtol=1e-9)      # :1037
        assert torch.allclose(log["n_z"].abs().max(dim=1).values, log["max_nz"], rtol=1e-12)
        assert float(log["wa_norm"].max()) == 1.0 and bool((log["Sm"] > 0).all())
        th_fl = log["theta"][:, split:].cpu().numpy()
        f = np.fft.fft(th_fl, axis=-1)
        psd = np.abs(f * np.conj(f) / ((steps - split) * 0.01))[:, :(steps - split) // 2]
        om = np.arange((steps - split) // 2) / ((steps - split) * 0.01)
        assert np.allclose(log["Sm"][:, 0].cpu().numpy(), (psd * om).sum(axis=1), rtol=1e-8)   # :1033-1034
    # repetition r starts from the same weights in every configuration: the first step's action is identical
    a0 = torch.stack([o[2]["action_cmd"][:, 0] for o in out])
    assert torch.equal(a0[0], a0[1]) and torch.equal(a0[0], a0[2])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "nl_rls_*.npz"))))
def test_rls3_update_equals_verbatim_reference(path):
    """RLS with the nonlinear task's dimensions (3 states + 1 action), step-level: bit for bit the VERBATIM reference
    class (objects.py:439-549) on the golden sequence, reset included."""
    from rl4afcs_b200.objects import RLS

    g = np.load(path)
    B = 3
    m = RLS({"state_dim": 3, "action_dim": 1, "rls_gamma": float(g["gamma"]), "rls_cov": 10 ** 6}, batch=B, dtype="mixed")
    assert m.params.shape == (B, 4, 3) and m.Cov.shape == (B, 4, 4)
    for k in range(len(g["da0"])):
        if k == int(g["reset_at"]):
            m._reset()
        ex = lambda v, w: torch.as_tensor(v).reshape(1, w, 1).expand(B, w, 1)   # noqa: E731
        m.update(ex(g["dx0"][k], 3), ex(g["da0"][k:k + 1], 1), ex(g["dx1"][k], 3))
        assert np.array_equal(m.params.cpu().numpy().reshape(B, 12), np.broadcast_to(g["theta"][k], (B, 12))), k
        assert np.array_equal(m.Cov.cpu().numpy().reshape(B, 16), np.broadcast_to(g["cov"][k], (B, 16))), k
        assert np.array_equal(m.epsilon.cpu().numpy().reshape(B, 3), np.broadcast_to(g["eps"][k], (B, 3))), k
        assert np.array_equal(m.eps_norm.cpu().numpy(), np.full(B, g["eps_norm"][k])), k
    th = g["theta"][-1].reshape(4, 3)
    assert np.array_equal(m.F.cpu().numpy()[0], th[:3].T) and np.array_equal(m.G.cpu().numpy()[0], th[3:].T)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "nl_env_*.npz"))))
def test_nonlinear_env_equals_verbatim_reference_wrapper(path):
    """Ce500NonLinear.reset / step on the GPU against the VERBATIM wrapper class run around the same surrogate plant
    (tests/golden/nl_env_*.npz): every returned quantity over 700 steps, six fault scenarios, both integrators."""
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear

    g = np.load(path)
    B = 3
    th_full = np.zeros(9000); th_full[: len(g["theta_ref"])] = g["theta_ref"]
    env = Ce500NonLinear(_nl_env_config(th_full, fault_scenario=str(g["fault"]), fault_time=float(g["fault_time"])), batch=B,
                         dtype="mixed", integrator=str(g["integrator"]), plant="surrogate")
    env._set_weight_matrices([int(k) for k in g["kappa"]])
    s, r, term, trunc, info = env.reset()
    assert np.array_equal(info["x_full"].cpu().numpy(), np.broadcast_to(g["x_reset"], (B, 12)))
    ulp = lambda a, b, n=2: bool(np.all(np.abs(a - b) <= n * np.spacing(np.maximum(np.abs(a), np.abs(b)))))   # noqa: E731
    for k in range(len(g["actions"])):
        s, r, term, trunc, info = env.step(g["actions"][k])
        same = lambda t, want: np.array_equal(t.cpu().numpy(), np.broadcast_to(want, (B,) + np.shape(want)))   # noqa: E731
        assert same(info["x_full"], g["x_full"][k]) and same(s, g["s"][k]) and same(info["e"], g["e"][k]), k
        assert same(info["action_commanded"], g["a_cmd"][k]) and same(info["action_effective"], g["a_eff"][k]), k
        assert same(info["x"][0][:, :, 0], g["x_lon"][k]) and same(info["x"][1][:, :, 0], g["x_lat"][k]), k
        assert same(info["reward_grad"][0][:, 0], g["rg_lon"][k]), k
        assert np.allclose(info["reward_grad"][1][:, 0].cpu().numpy(), g["rg_lat"][k], rtol=0, atol=0), k
        rr = r.cpu().numpy().reshape(B, 2)
        assert ulp(rr[:, 0], g["reward"][k, 0]) and ulp(rr[:, 1], g["reward"][k, 1], 4), k
        assert ulp(info["RSE"][0].cpu().numpy(), g["RSE"][k, 0]) and ulp(info["RSE"][1].cpu().numpy(), g["RSE"][k, 1], 4), k
        assert info["nans"] == bool(g["nans"][k]) and abs(info["t"] - g["t"][k]) < 1e-12


@pytest.mark.parametrize("name", ["default", "ms_notrace_rk4", "replacing_fault"])
def test_idhpnonlin_train_equals_verbatim_reference_run(name):
    """IDHPnonlin(env, config).train() on the GPU against golden runs of the VERBATIM nonlinear agent (objects.py's
    IDHPnonlin / Actor_big / Critic_big / RLS on the TensorFlow stand-in, verbatim Ce500NonLinear wrapper, surrogate plant;
    oracle/make_golden.py::nl_loop_fixture): the full log dict bit for bit, `_adapt_check` under NEP 50 like the fixture."""
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear
    from rl4afcs_b200.objects import IDHPnonlin

    g = np.load(os.path.join(GOLD, f"nl_loop_{name}.npz"))
    steps, B = int(g["steps"]), 2
    th = np.zeros(9000); th[:steps] = g["theta_ref"]
    elig = {"None": None}.get(str(g["elig"]), str(g["elig"]))
    env = Ce500NonLinear(_nl_env_config(th, fault_scenario=str(g["fault"]), fault_time=float(g["fault_time"]), t_end=steps * 0.01,
                                        total_steps=steps), batch=B, dtype="mixed", integrator=str(g["integrator"]), plant="surrogate")
    idhp_config = {"gamma": 0.6, "multistep": int(g["multistep"]), "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": float(g["lambda_l"]),
                   "kappa": [1, 2, 1], "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": float(g["warmup"]), "error_thresh": 1,
                   "tau": 0.02, "in_dims": 4,
                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 35.0, "eta_l": 5.0, "elig": elig},
                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.4, "eta_l": 0.7, "elig": 1233},
                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}
    w = {k: np.broadcast_to(g[f"w_{k}"], (B,) + g[f"w_{k}"].shape).copy() for k in ("W1a", "W2a", "W1c", "W2c")}
    idhp = IDHPnonlin(env, idhp_config, seed=1, verbose=0, weights=w, log="full", log_agents=B, chunk=250)   # DEFAULT mode = the fixtures'
    idhp.train(steps, noise=np.repeat(g["noise"][:, None], B, axis=1))
    lg = {k: v.cpu().numpy() for k, v in idhp.log.items()}
    for b in range(B):
        for k in ("eta_a", "x_full", "RSE", "x", "a_cmd", "a_eff", "s", "yref", "e", "a_weights2", "c_weights2", "a_grad",
                  "rls_params", "rls_eps_hist", "rls_eps_norm"):
            assert np.array_equal(lg[k][b], g[f"log_{k}"], equal_nan=True), k
        for k in ("a_weights1", "c_weights1", "c_grad", "rls_cov"):
            assert np.array_equal(lg[k][b][9::10], g[f"log10_{k}"], equal_nan=True), k
        assert np.allclose(lg["t"][b], g["log_t"], rtol=0, atol=1e-12)
        assert np.array_equal(idhp.actor.E[b].cpu().numpy().ravel(), g["final_E"])
        assert np.array_equal(idhp.target_critic.trainable_weights[1][b].double().cpu().numpy().ravel(), g["final_W2t"])
        assert np.allclose(float(idhp.RSE[0][b]), g["RSE_total"][0], rtol=1e-14)


def test_mc_run_seed_equals_verbatim_reference(oracle):
    """functions.MC_run_seed for one trained agent against the output dict of the VERBATIM functions.MC_run_seed
    (functions.py:39-60, run on the TensorFlow stand-in; tests/golden/sp_mc_run_seed.npz): trajectories bit for bit,
    per-step norms / episode return to rounding, convergence time exactly."""
    from rl4afcs_b200 import functions as F
    from rl4afcs_b200.envs.linear.env import Ce500ShortPeriod
    from rl4afcs_b200.objects import IDHPsp

    g = np.load(os.path.join(GOLD, "sp_mc_run_seed.npz"))
    steps, B = int(g["steps"]), 2
    env = Ce500ShortPeriod(_env_config(oracle, np.zeros((2, 1)), None), batch=B, dtype="mixed")
    w = {k: np.broadcast_to(g[f"w_{k}"], (B,) + g[f"w_{k}"].shape).copy() for k in ("W1a", "W2a", "W1c", "W2c")}
    idhp = IDHPsp(env, oracle.default_idhp_config(), verbose=0, seed=int(g["seed"]), weights=w, log="full", log_agents=B)
    idhp.train(steps)
    out = {k: v.cpu().numpy() for k, v in F.MC_run_seed(idhp).items()}
    for b in range(B):
        assert np.array_equal(out["x_array"][b], g["out_x_array"]) and np.array_equal(out["a_array"][b], g["out_a_array"])
        assert np.array_equal(out["ref_hist"][b], g["out_ref_hist"])
        nz = np.abs(g["out_c_array"]) > 0
        assert (np.abs(out["c_array"][b] - g["out_c_array"])[nz] <= 2 * np.spacing(np.abs(g["out_c_array"][nz]))).all()
        for k in ("wa_array", "wc_array"):
            assert np.allclose(out[k][b], g[f"out_{k}"], rtol=1e-13, atol=0), k
        for k in ("a_grad", "c_grad"):                                         # float32 norms (objects.py:706-707)
            assert np.allclose(out[k][b], g[f"out_{k}"], rtol=5e-7, atol=0), k
        for k in ("p_array", "cov_array", "eps_array"):                        # logged from step 2 on (objects.py:704)
            assert np.allclose(out[k][b][2:], g[f"out_{k}"][2:], rtol=1e-13, atol=0), k
        assert np.isclose(out["sum_c_array"][b], float(g["out_sum_c_array"]), rtol=1e-12)
        assert out["converged_time"][b] == float(g["out_converged_time"])
    st = idhp.stats()
    assert float(st["converged_time"][0]) == float(g["out_converged_time"])     # the in-kernel statistic agrees too
    assert np.isclose(float(st["sum_c"][0]), float(g["out_sum_c_array"]), rtol=1e-12)


def test_mc_test_hparam_equals_verbatim_reference():
    """functions.MC_test_hparam (2 algorithms x 2 repetitions of the full 90 s nonlinear flight with a c.g.-shift fault,
    one batch of 4 agents) against the `log` dict the VERBATIM functions.MC_test_hparam produced on the TensorFlow /
    plant stand-ins (tests/golden/nl_mc_test_hparam.npz): trajectories bit for bit (every 25th sample stored),
    RSE warm-up / flight, peak n_z, normalised weight norms and the smoothness index to rounding."""
    from rl4afcs_b200 import functions as F
    from rl4afcs_b200.envs.nonlinear.env import Ce500NonLinear

    g = np.load(os.path.join(GOLD, "nl_mc_test_hparam.npz"))
    N, reps = int(g["N"]), int(g["repetitions"])
    B = N * reps
    env = Ce500NonLinear(_nl_env_config(g["theta_ref"], fault_scenario=str(g["fault"]), fault_time=float(g["fault_time"])), batch=B,
                         dtype="mixed", plant="surrogate")
    elig = [None if e == "None" else str(e) for e in g["cfg_elig"]]
    configs = {k: list(g[f"cfg_{k}"]) for k in ("etaah", "etaal", "etach", "etacl", "lambda_hs", "lambda_ls", "seeds", "ms")}
    configs["elig"] = elig
    w = {k: np.stack([g[f"w{r}_{k}"] for _ in range(N) for r in range(reps)]) for k in ("W1a", "W2a", "W1c", "W2c")}
    noise = np.stack([g["noise"][r] for _ in range(N) for r in range(reps)], axis=1)           # (9000, B)
    out = F.MC_test_hparam(configs, "unused/", env, N, reps, noise=noise, weights=w)          # DEFAULT mode = the fixtures'
    assert [o[0] for o in out] == ["idhpat", "midhp"]
    for i, (algo, cfg, log) in enumerate(out):
        lg = {k: v.cpu().numpy() for k, v in log.items()}
        for k in ("e", "theta", "alpha", "q", "V", "h", "action_cmd", "action_eff", "n_z", "rls_eps"):
            assert np.array_equal(lg[k][:, ::25], g[f"log{i}_{k}"], equal_nan=True), (algo, k)
        for k in ("wa_norm", "wc_norm"):
            assert np.allclose(lg[k][:, ::25], g[f"log{i}_{k}"], rtol=1e-13, atol=0, equal_nan=True), (algo, k)
        assert np.allclose(lg["RSE"], g[f"log{i}_RSE"], rtol=1e-11, atol=0), algo
        assert np.allclose(lg["Sm"], g[f"log{i}_Sm"], rtol=1e-9, atol=0), algo
        assert np.allclose(lg["max_nz"], g[f"log{i}_max_abs_nz"], rtol=1e-15, atol=0), algo


def test_mc_run_metrics_equal_verbatim_reference(oracle):
    """functions.MC_run (2 hyper-parameter sets x 10 seeds in one batch) against the `metrics` dict of the VERBATIM
    functions.MC_run (functions.py:62-232; process pool replaced by a serial stand-in, tests/golden/sp_mc_run.npz): counts
    of diverged / late-converging runs, average return over the survivors, average convergence time, PSD error."""
    from rl4afcs_b200 import functions as F

    g = np.load(os.path.join(GOLD, "sp_mc_run.npz"))
    seeds, n_cfg = int(g["seeds"]), 2
    base, amp = oracle.default_reference()
    env_config = {"state_dim": 2, "action_dim": 1, "x0": np.zeros((2, 1)), "dt": 0.02, "t_end": 60, "fault_time": 20,
                  "fault_scenario": None, "reference": {"tracked_state": ["alpha"], "signal": [amp * base]}}
    configs = {}
    for k in ("lambda_hs", "lambda_ls", "kappas", "cooldown_times", "sigmas", "warmup_times", "lr_a_hs", "lr_a_ls", "lr_c_hs", "lr_c_ls",
              "multistep"):
        v = g[f"cfg_{k}"]
        configs[k] = None if v.ndim == 0 else [float(x) for x in v]
    configs["multistep"] = [int(x) for x in configs["multistep"]]
    configs["elig_a"] = [None if e == "None" else str(e) for e in g["cfg_elig_a"]]
    W = np.tile(g["weights"], (n_cfg, 1))                                    # agent = config * seeds + seed
    w = {"W1a": W[:, 0:4], "W2a": W[:, 4:8], "W1c": W[:, 8:12], "W2c": W[:, 12:20]}
    metrics, idhp = F.MC_run(n_cfg, configs, env_config, seeds, dtype="mixed", log_agents=n_cfg * seeds, weights=w)
    st = idhp.stats()
    for c in range(n_cfg):
        m = metrics[c]
        assert m["diverged"] == int(g[f"metrics{c}_diverged"]) and m["unsteady_convergence"] == int(g[f"metrics{c}_unsteady_convergence"])
        assert m["avg_c"] == float(g[f"metrics{c}_avg_c"]) and m["avg_t"] == float(g[f"metrics{c}_avg_t"])
        assert np.isclose(m["avg_PSD_err"], float(g[f"metrics{c}_avg_PSD_err"]), rtol=1e-6)
        sl = slice(c * seeds, (c + 1) * seeds)
        assert np.array_equal(st["converged_time"][sl].cpu().numpy(), g[f"arrays{c}_converged_time"])
        ok = ~st["diverged"][sl].cpu().numpy()
        assert np.allclose(st["sum_c"][sl].cpu().numpy()[ok], g[f"arrays{c}_sum_c"], rtol=1e-12)


@pytest.mark.parametrize("policy", ["fp64", "mixed", "fp32"])
def test_episode_statistics_and_nmae_equal_oracle(oracle, policy):
    """SpEngine.stats() (rl4_sp_agent_stats) against the oracle's episode statistics -- sum(c)/kappa, convergence time,
    divergence, mean|e| and the normalised tracking error nMAE = mean|e| / (max ref - min ref) BASELINE.json names -- bit
    for bit, with per-agent kappa / reference amplitude and a few diverging agents; and the rank summary of
    rl4_stats_reduce (dist.episode_summary_tensor) against a float64 torch reduction of the same planes."""
    from rl4afcs_b200 import _lib, sp_engine
    from rl4afcs_b200 import dist as rdist

    n, steps = 700, 1200
    ic = oracle.default_idhp_config()
    base, _ = oracle.default_reference()
    rng = np.random.default_rng(5)
    x0 = np.deg2rad(rng.uniform(-20, 20, size=(n, 2)))
    w = oracle.init_weights(n, 6)
    cfg = oracle.make_cfg(ic, n=n)
    cfg["kappa"] = rng.uniform(800, 1500, n)
    cfg["ref_amp"] = np.deg2rad(rng.uniform(1, 10, n)) * rng.choice([-1.0, 1.0], n)
    cfg["eta_c_h"] = rng.uniform(0.3, 4.0, n)                                    # sane to absurd: part of the batch blows up
    st = oracle.init_states(policy, cfg, x0, w)
    oracle.run(policy, cfg, base, st, 0, steps, tanh="t13")
    want = oracle.episode_stats(st, cfg, base, steps)
    eng = sp_engine.SpEngine(n, policy=policy)
    sp_engine.apply_idhp_config(eng, ic, dt=0.02)
    eng.set_hp("KAPPA", cfg["kappa"]); eng.set_hp("REF_AMP", cfg["ref_amp"]); eng.set_hp("ETA_C_H", cfg["eta_c_h"])
    eng.set_hpi("FAULT_STEP", -1); eng.set_hpi("FAULT_KIND", 0)
    eng.set_reference(base)
    eng.init(x0, w["W1a"], w["W2a"], w["W1c"], w["W2c"])
    eng.run(steps)
    got = {k: v.cpu().numpy() for k, v in eng.stats().items()}
    assert want["diverged"].any() and not want["diverged"].all()
    for k in ("sum_c", "converged_time", "diverged", "mean_abs_e", "nmae"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    ok = ~want["diverged"]
    assert np.all(got["nmae"][ok] > 0) and np.all(np.isfinite(got["nmae"][ok]))
    # the per-rank summary: deterministic device reduction == float64 reduction of the same per-agent planes
    part = rdist.episode_summary_tensor(eng).cpu().numpy()
    planes = eng.stats_planes().cpu().numpy()
    S = _lib.SPS
    for f in range(S["COUNT"]):
        kept, every = planes[f][ok].sum(), planes[f].sum()
        assert np.isclose(part[2 * f], kept, rtol=1e-12, atol=0), f
        assert np.isclose(part[2 * f + 1], every, rtol=1e-12, atol=0, equal_nan=True), f
    assert part[2 * S["COUNT"]] == ok.sum() and part[2 * S["COUNT"] + 1] == (~ok).sum()
    assert np.array_equal(part, rdist.episode_summary_tensor(eng).cpu().numpy(), equal_nan=True)      # run-to-run identical
    summ = rdist.gather_episode_summary(eng, 1)
    assert summ["agents"] == n and summ["diverged"] == int((~ok).sum())
    assert np.isclose(summ["avg_nmae"], got["nmae"][ok].mean(), rtol=1e-12)


def test_mc_run_sigma_sweep_pairs_seeds_across_configs():
    """functions.MC_run honours `sigmas` per config (functions.py:80,97) and, like the reference's seeds 0..seeds-1 per config,
    starts seed s of every config from the same standard-normal draw, scaled by that config's sigma."""
    from rl4afcs_b200 import functions as F
    from rl4afcs_b200 import sp_engine

    seeds, n_cfg = 6, 3
    t = np.linspace(0, 4, 200)
    env_config = {"state_dim": 2, "action_dim": 1, "x0": np.zeros((2, 1)), "dt": 0.02, "t_end": 4, "fault_time": 20,
                  "fault_scenario": None, "reference": {"tracked_state": ["alpha"], "signal": [np.deg2rad(5) * np.sin(2 * np.pi * t / 10)]}}
    configs = {"sigmas": [0.05, 0.1, 0.2]}
    _, idhp = F.MC_run(n_cfg, configs, env_config, seeds, dtype="mixed", base_seed=3)
    w0 = idhp._init_weights
    std = sp_engine.truncated_normal_weights(seeds, 3, 1.0, "cuda:0")
    for key in ("W1a", "W2a", "W1c", "W2c"):
        W = w0[key].reshape(n_cfg, seeds, -1)
        for c, sig in enumerate(configs["sigmas"]):
            want = (std[key].float() * torch.tensor(sig, dtype=torch.float32)).double()
            assert torch.equal(W[c], want), (key, c)
    assert float(w0["W1a"].abs().max()) <= 0.4 + 1e-7
