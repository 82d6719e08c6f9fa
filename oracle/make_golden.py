"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the reference ITSELF.

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

The reference has no tests and no golden vectors of its own, and its
TensorFlow agent cannot run here; what CAN run unmodified is `Ce500ShortPeriod`
(envs/linear/env.py:7-264) and `RLS` (objects.py:439-549).  The fixtures below are the outputs
of those verbatim classes (numpy 2.3.5, scipy-openblas 0.3.30, x86-64 AVX-512 host):

  sp_env_<fault>.npz     verbatim env driven open-loop by a seeded float32 action sequence
  sp_rls_<tag>.npz       verbatim RLS driven by seeded regressors (incl. a mid-run _reset())
  sp_loop_<case>.npz     full IDHP loop: the reference-shaped numpy loop (oracle/sp_numpy.py,
                         TensorFlow ops restated in float32 numpy) running on the VERBATIM env
                         and the VERBATIM RLS objects, tanh = t13 (and one case with np.tanh)

The GPU box has no /root/reference: tests there read these files only.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, sp_c, sp_numpy  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def env_fixture(fault, seed, steps=1500, x0=(0.02, -0.03), kappa=1140):
    Env = ref_loader.load_reference_linear_env()
    base, amp = sp_c.default_reference()
    ref = amp * base
    env = Env({"state_dim": 2, "action_dim": 1, "x0": np.array(x0, dtype=float).reshape(2, 1), "dt": 0.02,
               "t_end": 60, "fault_time": 20, "fault_scenario": fault,
               "reference": {"tracked_state": ["alpha"], "signal": [ref]}})
    env.kappa = kappa
    env.reset(seed=0)
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, size=steps).astype(np.float32)
    xs, cs, es, gs = [], [], [], []
    for k in range(steps):
        obs, c, done, _, info = env.step(np.float32(20) * a[k].reshape(1, 1))
        xs.append(obs.ravel().copy()); cs.append(float(c)); es.append(float(info["e"]))
        gs.append(info["reward_grad"].ravel().copy())
    return dict(fault=str(fault), x0=np.array(x0), kappa=kappa, action=a, x=np.array(xs), reward=np.array(cs),
                e=np.array(es), reward_grad=np.array(gs), A=env.A.copy(), B=env.B.copy())


def rls_fixture(gamma, seed, steps=400, reset_at=250):
    RLS = ref_loader.load_reference_rls()
    m = RLS({"state_dim": 2, "action_dim": 1, "rls_gamma": gamma, "rls_cov": 10 ** 6})
    rng = np.random.default_rng(seed)
    dx0 = rng.normal(size=(steps, 2, 1)) * 1e-3
    da0 = (rng.normal(size=(steps, 1, 1)) * 1e-2).astype(np.float32)
    Atrue = np.array([[0.985, 0.0195], [-0.0294, 0.9687]]); Btrue = np.array([[-0.0006], [-0.0469]])
    dx1 = Atrue @ dx0 + Btrue @ da0.astype(np.float64) + rng.normal(size=(steps, 2, 1)) * 1e-7
    th, cv, ep, en = [], [], [], []
    for k in range(steps):
        if k == reset_at:
            m._reset()
        m.update(dx0[k], da0[k], dx1[k])
        th.append(m.params.ravel().copy()); cv.append(m.Cov.ravel().copy())
        ep.append(m.epsilon.ravel().copy()); en.append(float(m.eps_norm))
    return dict(gamma=gamma, reset_at=reset_at, dx0=dx0[:, :, 0], da0=da0[:, 0, 0], dx1=dx1[:, :, 0],
                theta=np.array(th), cov=np.array(cv), eps=np.array(ep), eps_norm=np.array(en))


def loop_fixture(name, *, x0, seed, fault=None, elig=(None, None), ms=2, steps=1100, tanh="t13", tracked="alpha"):
    """The VERBATIM agent: objects.py's IDHPsp / Actor / Critic / RLS classes (executed on the TensorFlow stand-in of
    oracle/tf_shim.py) driving the verbatim Ce500ShortPeriod; only the initial weights are injected."""
    Env = ref_loader.load_reference_linear_env()
    O, tf = ref_loader.load_reference_objects()
    base, amp = sp_c.default_reference()
    ic = sp_c.default_idhp_config()
    ic["multistep"] = ms
    ic["actor_config"]["elig"], ic["critic_config"]["elig"] = elig
    t_end = steps * 0.02
    assert int(t_end / 0.02) == steps
    env = Env({"state_dim": 2, "action_dim": 1, "x0": np.array(x0, dtype=float).reshape(2, 1), "dt": 0.02,
               "t_end": t_end, "fault_time": 20, "fault_scenario": fault,
               "reference": {"tracked_state": [tracked], "signal": [amp * base]}})
    w = sp_c.init_weights(1, seed)
    tf.set_tanh((lambda a: sp_c.tanh_t13(np.asarray(a, dtype=np.float32))) if tanh == "t13" else (lambda a: np.tanh(a)))
    idhp = O.IDHPsp(env, ic, verbose=False, seed=seed)
    idhp.actor.set_weights([w["W1a"][0].reshape(1, 4), w["W2a"][0].reshape(4, 1)])
    idhp.critic.set_weights([w["W1c"][0].reshape(1, 4), w["W2c"][0].reshape(4, 2)])
    idhp.train()
    tf.set_tanh(None)
    out = dict(x0=np.array(x0, dtype=float), seed=seed, fault=str(fault), elig_a=str(elig[0]), elig_c=str(elig[1]),
               multistep=ms, steps=steps, tanh=tanh, tracked=tracked, **{f"w_{k}": v[0] for k, v in w.items()})
    f64 = lambda v, w_: np.asarray(v, dtype=np.float64).reshape(steps, w_) if w_ else np.asarray(v, dtype=np.float64).reshape(steps)   # noqa: E731
    out.update(x=f64(idhp.x_hist, 2), a=f64(idhp.a_hist, 0), c=f64(idhp.c_hist, 0), ref=f64(idhp.ref_hist, 0),
               a_w1=f64(idhp.a_weights_hist1, 4), a_w2=f64(idhp.a_weights_hist2, 4), c_w1=f64(idhp.c_weights_hist1, 4),
               c_w2=f64(idhp.c_weights_hist2, 8), params=f64(idhp.params_hist, 6), cov=f64(idhp.cov_hist, 9),
               eps_norm=f64(idhp.eps_norm_hist, 0), a_e=f64(idhp.a_e_hist, 8), c_e=f64(idhp.c_e_hist, 24),
               a_all_grad=f64(idhp.a_all_grad_hist, 8), c_all_grad=f64(idhp.c_all_grad_hist, 12))
    return out


def rls3_fixture(gamma, seed, steps=400, reset_at=250):
    """Verbatim RLS (objects.py:439-549) with the nonlinear task's dimensions: 3 states + 1 action."""
    RLS = ref_loader.load_reference_rls()
    m = RLS({"state_dim": 3, "action_dim": 1, "rls_gamma": gamma, "rls_cov": 10 ** 6})
    rng = np.random.default_rng(seed)
    dx0 = rng.normal(size=(steps, 3, 1)) * 1e-3
    da0 = (rng.normal(size=(steps, 1, 1)) * 1e-2).astype(np.float32)
    Atrue = np.array([[0.99, 0.0, 0.0098], [0.0, 1.0, 0.01], [-0.016, 0.0, 0.984]]); Btrue = np.array([[-0.0003], [0.0], [-0.033]])
    dx1 = Atrue @ dx0 + Btrue @ da0.astype(np.float64) + rng.normal(size=(steps, 3, 1)) * 1e-7
    th, cv, ep, en = [], [], [], []
    for k in range(steps):
        if k == reset_at:
            m._reset()
        m.update(dx0[k], da0[k], dx1[k])
        th.append(m.params.ravel().copy()); cv.append(m.Cov.ravel().copy())
        ep.append(m.epsilon.ravel().copy()); en.append(float(m.eps_norm))
    return dict(gamma=gamma, reset_at=reset_at, dx0=dx0[:, :, 0], da0=da0[:, 0, 0], dx1=dx1[:, :, 0],
                theta=np.array(th), cov=np.array(cv), eps=np.array(ep), eps_norm=np.array(en))


def nl_env_fixture(fault, seed, steps=700, fault_time=3.0, integrator="ode5"):
    """VERBATIM Ce500NonLinear wrapper (envs/nonlinear/env.py:11-319) around the surrogate plant: reset + `steps` steps
    with smooth random three-surface commands; everything the wrapper returns is recorded."""
    from oracle import nl_c
    Env, stub = ref_loader.load_reference_nonlinear_env(integrator)
    th = nl_c.theta_reference()
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    env_config = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01,
                  "t_end": 90, "total_steps": 9000, "fault_time": fault_time, "fault_scenario": fault,
                  "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    env = Env(env_config)
    env._set_weight_matrices([1, 2, 1])                    # objects.py:1029 with idhp_nonlin.py's kappa
    s0, r0, _, _, info0 = env.reset()
    rng = np.random.default_rng(seed)
    t = np.arange(steps) * 0.01
    acts = np.stack([0.5 * np.sin(2 * np.pi * t / 3.1 + rng.uniform(0, 6)) + 0.2 * rng.standard_normal(steps),
                     0.15 * np.sin(2 * np.pi * t / 2.3 + rng.uniform(0, 6)) + 0.05 * rng.standard_normal(steps),
                     0.10 * np.sin(2 * np.pi * t / 4.7 + rng.uniform(0, 6)) + 0.05 * rng.standard_normal(steps)], axis=1)
    acts = np.clip(acts, -1, 1)
    out = {k: [] for k in ("x_full", "s", "reward", "e", "RSE", "a_cmd", "a_eff", "rg_lon", "rg_lat", "x_lon", "x_lat", "nans", "t")}
    for k in range(steps):
        s, r, term, trunc, info = env.step(acts[k].copy())
        assert term is None and trunc is False
        out["x_full"].append(np.array(info["x_full"])); out["s"].append(np.array(s)); out["reward"].append(np.array(r).reshape(2))
        out["e"].append(np.array(info["e"])); out["RSE"].append(np.array(info["RSE"], dtype=np.float64))
        out["a_cmd"].append(np.array(info["action_commanded"])); out["a_eff"].append(np.array(info["action_effective"]))
        out["rg_lon"].append(np.array(info["reward_grad"][0]).ravel()); out["rg_lat"].append(np.array(info["reward_grad"][1]).ravel())
        out["x_lon"].append(info["x"][0].ravel()); out["x_lat"].append(info["x"][1].ravel())
        out["nans"].append(bool(info["nans"])); out["t"].append(info["t"])
    res = {k: np.array(v) for k, v in out.items()}
    res.update(actions=acts, fault=str(fault), fault_time=fault_time, integrator=integrator, x_reset=np.array(info0["x_full"]),
               theta_ref=th[:steps], kappa=np.array([1, 2, 1]))
    return res


def nl_loop_fixture(seed, *, fault="none", fault_time=60, integrator="ode5", elig="accumulating", ms=0, steps=600, warmup=1.5,
                    lambda_l=0.8, plant="surrogate"):
    """The VERBATIM nonlinear agent: objects.py's IDHPnonlin / Actor_big / Critic_big / RLS on the TensorFlow stand-in,
    driving the verbatim Ce500NonLinear wrapper around the surrogate plant.  numpy >= 2 here, so `_adapt_check` follows
    NEP 50 (oracle / kernel flag numpy2 = 1).  A short warm-up puts the learning-rate decay inside the run."""
    from oracle import nl_c
    O, tf = ref_loader.load_reference_objects()
    # plant="binary": the wrapper drives the reference's REAL plant binary, executing natively (oracle/pe_probe/pe_citation.py) --
    # then nothing in the loop is a stand-in except the arithmetic inside TensorFlow ops
    Env, stub = ref_loader.load_reference_nonlinear_env(integrator, plant=plant)
    th = nl_c.theta_reference()
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    env_config = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01,
                  "t_end": steps * 0.01, "total_steps": steps, "fault_time": fault_time, "fault_scenario": fault,
                  "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    assert int(env_config["t_end"] / 0.01) == steps
    idhp_config = {"gamma": 0.6, "multistep": ms, "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": lambda_l, "kappa": [1, 2, 1],
                   "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": warmup, "error_thresh": 1, "tau": 0.02, "in_dims": 4,
                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 35.0, "eta_l": 5.0, "elig": elig},
                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.4, "eta_l": 0.7, "elig": 1233},
                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}     # idhp_nonlin.py:123-146
    noise = np.random.default_rng(seed).standard_normal(steps).astype(np.float32)
    tf.set_tanh(lambda v: sp_c.tanh_t13(np.asarray(v, dtype=np.float32)))
    tf.set_noise(noise[1:])                       # objects.py:1375 draws inside `if step > 0`
    try:
        env = Env(env_config)
        idhp = O.IDHPnonlin(env, idhp_config, verbose=False, seed=seed)
        w = nl_c.init_weights(1, seed)
        idhp.actor.set_weights([w["W1a"][0].reshape(4, 10), w["W2a"][0].reshape(10, 1)])
        idhp.critic.set_weights([w["W1c"][0].reshape(4, 10), w["W2c"][0].reshape(10, 3)])
        idhp.train()
    finally:
        tf.set_tanh(None); tf.set_noise(None)
    L = idhp.log
    out = dict(seed=seed, plant=plant, fault=str(fault), fault_time=fault_time, integrator=integrator, elig=str(elig), multistep=ms, steps=steps,
               warmup=warmup, lambda_l=lambda_l, noise=noise, theta_ref=th[:steps], **{f"w_{k}": v[0] for k, v in w.items()})
    for k in ("eta_a", "t", "x_full", "RSE", "x", "a_cmd", "a_eff", "s", "yref", "e", "a_weights2", "c_weights2", "a_grad",
              "rls_params", "rls_eps_hist", "rls_eps_norm"):
        out[f"log_{k}"] = np.asarray(L[k], dtype=np.float64)
    for k in ("a_weights1", "c_weights1", "c_grad", "rls_cov"):                    # wide: every 10th row + the last
        out[f"log10_{k}"] = np.asarray(L[k], dtype=np.float64)[9::10]
    out["a_elig_nonzero"] = bool(np.any(L["a_elig"])) or bool(np.any(L["c_elig"]))
    out["final_E"] = np.asarray(idhp.actor.E, dtype=np.float64).ravel()
    out["final_W1t"] = idhp.target_critic.get_weights()[0].ravel().astype(np.float64)
    out["final_W2t"] = idhp.target_critic.get_weights()[1].ravel().astype(np.float64)
    out["RSE_total"] = np.asarray(idhp.RSE, dtype=np.float64)
    return out


def mc_run_seed_fixture(seed=3, steps=3000):
    """Output dict of the VERBATIM functions.MC_run_seed (functions.py:39-60: IDHPsp run + per-step norms, episode return,
    convergence time) on the TensorFlow stand-in; the initial weights its IDHPsp drew are recorded as inputs."""
    import contextlib
    import io

    Fn = ref_loader.load_reference_functions()
    O, tf = ref_loader.load_reference_objects()
    Env = ref_loader.load_reference_linear_env()
    base, amp = sp_c.default_reference()
    ic = sp_c.default_idhp_config()
    env = Env({"state_dim": 2, "action_dim": 1, "x0": np.zeros((2, 1)), "dt": 0.02, "t_end": steps * 0.02, "fault_time": 20,
               "fault_scenario": None, "reference": {"tracked_state": ["alpha"], "signal": [amp * base]}})
    tf.set_tanh(lambda a: sp_c.tanh_t13(np.asarray(a, dtype=np.float32)))
    try:
        probe = O.IDHPsp(env, ic, verbose=False, seed=seed)       # same seed -> the weights MC_run_seed's own agent will draw
        w = {"W1a": probe.actor.get_weights()[0].ravel(), "W2a": probe.actor.get_weights()[1].ravel(),
             "W1c": probe.critic.get_weights()[0].ravel(), "W2c": probe.critic.get_weights()[1].ravel()}
        with contextlib.redirect_stdout(io.StringIO()):
            out = Fn.MC_run_seed(seed, env, ic)
    finally:
        tf.set_tanh(None)
    res = {f"out_{k}": np.asarray(v, dtype=np.float64) for k, v in out.items()}
    res.update(seed=seed, steps=steps, **{f"w_{k}": np.asarray(v, dtype=np.float64) for k, v in w.items()})
    return res


def mc_run_fixture(seeds=10):
    """`metrics` of the VERBATIM functions.MC_run (functions.py:62-232) for two hyper-parameter sets x 10 seeds of the full
    short-period episode.  The process pool is replaced by a serial stand-in, pickling / plotting by a capture hook; the
    weights each seed draws are recorded as inputs.  The second set has a large critic rate so that some seeds (3 of 10) diverge."""
    import contextlib
    import io

    Fn = ref_loader.load_reference_functions()
    O, tf = ref_loader.load_reference_objects()
    Env = ref_loader.load_reference_linear_env()
    base, amp = sp_c.default_reference()
    env_config = {"state_dim": 2, "action_dim": 1, "x0": np.zeros((2, 1)), "dt": 0.02, "t_end": 60, "fault_time": 20,
                  "fault_scenario": None, "reference": {"tracked_state": ["alpha"], "signal": [amp * base]}}
    configs = {"lambda_hs": [0.576, 0.6], "lambda_ls": [0.296, 0.2], "kappas": [1140, 800], "cooldown_times": None, "sigmas": None,
               "warmup_times": None, "elig_a": [None, "accumulating"], "lr_a_hs": [3.55, 4.5], "lr_a_ls": [0.054, 0.05],
               "lr_c_hs": [0.338, 0.8], "lr_c_ls": None, "multistep": [2, 0]}

    class _Res:
        def __init__(self, v):
            self.v = v

        def get(self):
            return self.v

    class _SerialPool:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def apply_async(self, fn, args=()):
            return _Res(fn(*args))

        def close(self):
            pass

        def join(self):
            pass
    captured = []
    Fn.Pool = _SerialPool
    Fn.MC_pickling = lambda directory, arrays, idhp_config, metrics, save: captured.append(
        (dict(metrics), {k: np.array(arrays[k]) for k in ("sum_c_array", "converged_time")}))
    Fn.MC_plotting = lambda *a, **k: None
    tf.set_tanh(lambda a: sp_c.tanh_t13(np.asarray(a, dtype=np.float32)))
    try:
        env = Env(env_config)
        ws = []
        for s_ in range(seeds):
            probe = O.IDHPsp(env, sp_c.default_idhp_config(), verbose=False, seed=s_)
            ws.append(np.concatenate([probe.actor.get_weights()[0].ravel(), probe.actor.get_weights()[1].ravel(),
                                      probe.critic.get_weights()[0].ravel(), probe.critic.get_weights()[1].ravel()]))
        with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
            Fn.MC_run(2, configs, env, env_config, seeds, 0.3, "unused/", save=False)
    finally:
        tf.set_tanh(None)
    out = dict(seeds=seeds, weights=np.array(ws, dtype=np.float64))          # (seeds, 4 + 4 + 4 + 8)
    for k, v in configs.items():
        out[f"cfg_{k}"] = np.array(["None" if x is None else str(x) for x in v]) if (v is not None and k == "elig_a") else (
            np.asarray(v, dtype=np.float64) if v is not None else np.array("None"))
    for i, (m, arr) in enumerate(captured):
        for k, v in m.items():
            out[f"metrics{i}_{k}"] = np.asarray(v, dtype=np.float64)
        out[f"arrays{i}_sum_c"] = arr["sum_c_array"]; out[f"arrays{i}_converged_time"] = arr["converged_time"]
    return out


def mc_test_hparam_fixture(repetitions=2, plant="surrogate"):
    """The per-configuration `log` dict the VERBATIM functions.MC_test_hparam (functions.py:931-1060) hands to its plot
    routine -- two algorithms x `repetitions` full 90 s nonlinear episodes on the TensorFlow stand-in / plant stand-in.
    Summary numbers in full, trajectories every 25th sample.  Takes a few minutes."""
    import contextlib
    import io

    from oracle import nl_c
    Fn = ref_loader.load_reference_functions()
    O, tf = ref_loader.load_reference_objects()
    Env, stub = ref_loader.load_reference_nonlinear_env("ode5", plant=plant)      # plant="binary": the reference's real plant, natively
    th = nl_c.theta_reference()
    trim_input = np.array([-0.02855, 0, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.55, 0.55, 0])
    trim_state = np.array([0, 0, 0, 90, 0.0576, 0, 0, 0.0576, 0, 2000, 0, 0])
    env_config = {"state_dim": 4, "action_dim": 3, "trim_input": trim_input, "trim_state": trim_state, "dt": 0.01,
                  "t_end": 90, "total_steps": 9000, "fault_time": 60, "fault_scenario": "shift_cg",
                  "reference": {"tracked_state": ["phi", "theta", "psi"], "signal": [0 * th, th, 0 * th]}}
    N = 2
    configs = {"etaah": [35.0, 25.0], "etaal": [5.0, 5.0], "etach": [1.4, 1.0], "etacl": [0.7, 0.5], "lambda_hs": [0.95, 0.9],
               "lambda_ls": [0.95, 0.8], "seeds": [0, 0], "ms": [0, 1], "elig": ["accumulating", None]}
    noise = np.random.default_rng(77).standard_normal((repetitions, 9000)).astype(np.float32)
    captured = []
    Fn.MC_test_hparam_plot = lambda log, *a, **k: captured.append({kk: np.array(v) for kk, v in log.items()})
    # the weights seed r draws (Actor_big then Critic_big, each from a fresh initializer with that seed)
    weights = []
    for r in range(repetitions):
        env = Env(env_config)
        probe = O.IDHPnonlin(env, {"gamma": 0.6, "multistep": 0, "lr_decay": 0.998, "lambda_h": 0.95, "lambda_l": 0.95, "kappa": [1, 2, 1],
                                   "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 4, "error_thresh": 1, "tau": 0.02, "in_dims": 4,
                                   "actor_config": {"layers": {10: "tanh", 1: "tanh"}, "eta_h": 1.0, "eta_l": 1.0, "elig": None},
                                   "critic_config": {"layers": {10: "tanh", 3: "linear"}, "eta_h": 1.0, "eta_l": 1.0, "elig": 1233},
                                   "rls_config": {"state_dim": 3, "action_dim": 1, "rls_gamma": 1, "rls_cov": 10 ** 6}}, verbose=False, seed=r)
        weights.append({"W1a": probe.actor.get_weights()[0].ravel(), "W2a": probe.actor.get_weights()[1].ravel(),
                        "W1c": probe.critic.get_weights()[0].ravel(), "W2c": probe.critic.get_weights()[1].ravel()})
    tf.set_tanh(lambda v: sp_c.tanh_t13(np.asarray(v, dtype=np.float32)))
    # noise stream: config 0 rep 0, rep 1, ..., config 1 rep 0, ... ; each run consumes 8999 draws (steps 1..8999)
    tf.set_noise(np.concatenate([noise[r, 1:] for _ in range(N) for r in range(repetitions)]))
    try:
        env = Env(env_config)
        with contextlib.redirect_stdout(io.StringIO()):
            Fn.MC_test_hparam(configs, "unused/", env, N, repetitions, save=0, show=0)
    finally:
        tf.set_tanh(None); tf.set_noise(None)
    out = dict(N=N, repetitions=repetitions, plant=plant, noise=noise, theta_ref=th, fault="shift_cg", fault_time=60,
               **{f"cfg_{k}": np.array([str(x) for x in v]) if k == "elig" else np.asarray(v) for k, v in configs.items()})
    for r, w in enumerate(weights):
        for k, v in w.items():
            out[f"w{r}_{k}"] = np.asarray(v, dtype=np.float64)
    for i, lg in enumerate(captured):
        for k, v in lg.items():
            out[f"log{i}_{k}"] = v if v.ndim < 2 or v.shape[1] <= 2 else v[:, ::25]
        out[f"log{i}_max_abs_nz"] = np.max(np.abs(lg["n_z"]), axis=1)
    return out


def utils_fixture():
    """Outputs of the verbatim utils.py functions (samplers with true_random=False, PSD, convergence time, VD_A, KL)."""
    U = ref_loader.load_reference_utils()
    rng = np.random.default_rng(21)
    out = {}
    c = U.pick_continuous_hparams(7, lambda_hs=[0.3, 0.4], lr_a_hs=[2.5, 4.7], lr_c_hs=[0.45, 0.55], kappas=[1000, 1400],
                                  cooldown_times=[1.0, 3.0], true_random=False)
    for k, v in c.items():
        if v is not None and k != "elig_a":
            out[f"cont_{k}"] = np.asarray(v, dtype=np.float64)
    d = U.pick_discrete_hparams(9, lambda_hs=[0.2, 0.3, 0.4], lr_a_hs=[1.0, 2.0, 3.0, 4.0], sigmas=[0.05, 0.1],
                                elig_a=["accumulating", "replacing"], lr_decays=[0.99, 0.998], true_random=False)
    for k, v in d.items():
        if v is not None:
            out[f"disc_{k}"] = np.asarray(v) if k != "elig_a" else np.asarray([str(e) for e in v])
    sig = rng.standard_normal((5, 3000)) * 0.05
    out["psd_in"] = sig
    out["psd_out"], out["psd_omega"] = U.get_PSD(60, 0.02, sig)
    one = rng.standard_normal(3500)
    out["psd1_in"] = one
    out["psd1_out"], out["psd1_omega"] = U.get_PSD(35, 0.01, one)
    c_hist = -0.5 * 1200 * (np.deg2rad(3.0) * np.exp(-np.arange(3000) / 300.0) * np.cos(np.arange(3000) / 40.0)) ** 2
    out["conv_c"] = c_hist
    out["conv_t"] = np.asarray(U.get_convergence_time(c_hist, 1200, 0.02))
    X = list(rng.standard_normal(40) + 0.4); Y = list(rng.standard_normal(40)); Z = list(np.round(rng.standard_normal(25), 1))
    out["vda_X"], out["vda_Y"], out["vda_Z"] = np.asarray(X), np.asarray(Y), np.asarray(Z)
    out["vda_XY"] = np.asarray(U.VD_A(X, Y)[0]); out["vda_YX"] = np.asarray(U.VD_A(Y, X)[0])
    out["vda_ZZ"] = np.asarray(U.VD_A(Z, list(np.asarray(Z)[::-1] + 0.1))[0])
    u = rng.random(30); v = rng.random(30); u[3] = 0.0; v[7] = -1.0
    out["kl_u"], out["kl_v"], out["kl"] = u, v, np.asarray(U.kl_divergence(u, v))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "utils_functions.npz"), **utils_fixture())
    np.savez_compressed(os.path.join(OUT, "nl_rls_g1.npz"), **rls3_fixture(1, 17))
    np.savez_compressed(os.path.join(OUT, "nl_rls_g0998.npz"), **rls3_fixture(0.998, 18))
    for i, (f, integ) in enumerate([("none", "ode5"), ("damp_elevator_and_saturate_elevator", "ode5"), ("shift_cg", "rk4"),
                                    ("slow_all", "ode5"), ("damp_all", "ode5"), ("saturate_aileron", "rk4")]):
        np.savez_compressed(os.path.join(OUT, f"nl_env_{f}.npz"), **nl_env_fixture(f, 200 + i, integrator=integ))
    np.savez_compressed(os.path.join(OUT, "sp_mc_run_seed.npz"), **mc_run_seed_fixture())
    if "--skip-slow" not in sys.argv:
        np.savez_compressed(os.path.join(OUT, "nl_mc_test_hparam.npz"), **mc_test_hparam_fixture())
        np.savez_compressed(os.path.join(OUT, "sp_mc_run.npz"), **mc_run_fixture())
    nl_cases = {"default": dict(seed=41), "ms_notrace_rk4": dict(seed=42, ms=1, elig=None, integrator="rk4"),
                "replacing_fault": dict(seed=43, elig="replacing", fault="damp_elevator_and_saturate_elevator", fault_time=3.0)}
    for name, kw in nl_cases.items():
        np.savez_compressed(os.path.join(OUT, f"nl_loop_{name}.npz"), **nl_loop_fixture(**kw))
    # the same verbatim agent + verbatim wrapper on the reference's REAL plant binary (named nlbin_*: not part of the nl_loop_* glob)
    from oracle.pe_probe import pe_citation
    if pe_citation.available():
        bin_cases = {"default": dict(seed=51), "shiftcg_replacing_ms": dict(seed=52, ms=1, elig="replacing", fault="shift_cg", fault_time=3.0)}
        for name, kw in bin_cases.items():
            np.savez_compressed(os.path.join(OUT, f"nlbin_loop_{name}.npz"), **nl_loop_fixture(plant="binary", **kw))
        if "--skip-slow" not in sys.argv:
            np.savez_compressed(os.path.join(OUT, "nlbin_mc_test_hparam.npz"), **mc_test_hparam_fixture(plant="binary"))
    if "--only-utils" in sys.argv:
        return
    for i, f in enumerate([None, "invert_elevator", "damp_elevator", "shift_cg"]):
        np.savez_compressed(os.path.join(OUT, f"sp_env_{f or 'none'}.npz"), **env_fixture(f, 100 + i))
    np.savez_compressed(os.path.join(OUT, "sp_rls_g1.npz"), **rls_fixture(1, 7))
    np.savez_compressed(os.path.join(OUT, "sp_rls_g0995.npz"), **rls_fixture(0.995, 8))
    cases = {
        "default_x0zero": dict(x0=(0, 0), seed=4),                                  # idhp_sp.py as shipped
        "default_x0rand": dict(x0=(0.02, -0.03), seed=5),                           # exercises Q3
        "shiftcg_acc_1step": dict(x0=(-0.01, 0.03), seed=7, fault="shift_cg", elig=("accumulating", "accumulating"), ms=0),
        "invert_replacing": dict(x0=(0.015, 0.01), seed=8, fault="invert_elevator", elig=("replacing", "replacing")),
        "default_nptanh": dict(x0=(0.02, -0.03), seed=5, tanh="np"),
        "trackq_damp": dict(x0=(0.01, -0.02), seed=9, fault="damp_elevator", tracked="q"),      # envs/linear/env.py:180-184
    }
    for name, kw in cases.items():
        np.savez_compressed(os.path.join(OUT, f"sp_loop_{name}.npz"), **loop_fixture(name, **kw))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
