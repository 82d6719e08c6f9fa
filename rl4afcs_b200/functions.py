"""Batched Monte-Carlo front-end and episode statistics -- the hot-path part of functions.py / utils.py of
wingos80/RL4AFCS (plotting and pickling are out of scope).

  get_PSD, get_convergence_time     utils.py:188-236, 350-369     (torch, batched over agents)
  MC_run_seed                       functions.py:39-60            (the per-run output dict, for a whole batch)
  MC_run                            functions.py:62-232           (configs x seeds flattened onto the agent axis; returns
                                                                   the metrics dict of functions.py:223-227 per config)
  MC_test_hparam                    functions.py:931-1060         (nonlinear task: N configs x repetitions in one batch;
                                                                   returns the per-config log dicts of :1008-1052)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .envs.linear.env import Ce500ShortPeriod
from .objects import IDHPnonlin, IDHPsp
from .utils import get_PSD, get_convergence_time  # noqa: F401  (utils.py:188-236, 350-369)


def MC_run_seed(idhp: IDHPsp) -> dict:
    """The output dict of functions.py:44-60 for every logged agent of a trained batch."""
    out = {"x_array": idhp.x_hist, "a_array": idhp.a_hist.squeeze(-1), "c_array": idhp.c_hist,
           "wa_array": torch.linalg.vector_norm(idhp.a_weights_hist1, dim=-1) + torch.linalg.vector_norm(idhp.a_weights_hist2, dim=-1),
           "wc_array": torch.linalg.vector_norm(idhp.c_weights_hist1, dim=-1) + torch.linalg.vector_norm(idhp.c_weights_hist2, dim=-1),
           "p_array": torch.linalg.vector_norm(idhp.params_hist, dim=-1),
           "cov_array": torch.linalg.vector_norm(idhp.cov_hist, dim=-1), "eps_array": idhp.eps_norm_hist,
           "a_grad": idhp.a_grad_hist, "c_grad": idhp.c_grad_hist, "ref_hist": idhp.ref_hist}
    kappa = idhp.env.kappa
    if np.ndim(kappa):
        kappa = torch.as_tensor(np.asarray(kappa, dtype=np.float64), device=idhp.c_hist.device)[: idhp.c_hist.shape[0]]
        out["sum_c_array"] = idhp.c_hist.sum(dim=-1) / kappa
        out["converged_time"] = get_convergence_time(idhp.c_hist, kappa[:, None], idhp.env.dt)
        return out
    out["sum_c_array"] = idhp.c_hist.sum(dim=-1) / kappa                      # functions.py:53
    out["converged_time"] = get_convergence_time(idhp.c_hist, kappa, idhp.env.dt)
    return out


def MC_run(n_configs, configs, env_config, seeds, *, device="cuda", dtype="mixed", base_seed=0, log_agents=0, weights=None):
    """Monte-Carlo over ``n_configs`` hyper-parameter sets x ``seeds`` seeds in ONE batch (functions.py:62-232 ran
    a process pool of 10).  ``configs`` has the reference's keys (lambda_hs, lambda_ls, kappas, cooldown_times, sigmas,
    warmup_times, elig_a, lr_a_hs, lr_a_ls, lr_c_hs, lr_c_ls, multistep), each None or a list per config (``sigmas`` is
    honoured per config: seed s of every config starts from the same standard-normal draw scaled by that config's sigma).
    Returns (metrics per config, idhp): metrics = dict(avg_PSD_err?, avg_c, avg_t, diverged, unsteady_convergence)
    as functions.py:223-227 (PSD error only when trajectories are logged).  ``weights``: optional dict of (B, w) initial
    weights (agent index = config * seeds + seed).  Like the reference, ``avg_c`` averages the non-diverged runs while
    ``avg_t`` averages ALL runs (its second `np.delete`, functions.py:178, filters on an array that no longer has NaNs)."""
    dflt = dict(lambda_hs=0.34, lambda_ls=0.0, kappas=1200, cooldown_times=2.0, sigmas=0.1, warmup_times=3.0, elig_a=None,
                lr_a_hs=3.0, lr_a_ls=0.05, lr_c_hs=0.5, lr_c_ls=0.0, multistep=0)          # functions.py:75-86

    def col(key):
        v = configs.get(key)
        vals = [dflt[key] if v is None else v[c] for c in range(n_configs)]
        return np.repeat(np.asarray(vals, dtype=object), seeds)

    B = n_configs * seeds
    sig = col("sigmas").astype(np.float64)
    idhp_config = {"multistep": col("multistep").astype(np.int64), "gamma": 0.6, "gamma_rls": 1.0,
                   "lambda_h": col("lambda_hs").astype(np.float64), "lambda_l": col("lambda_ls").astype(np.float64),
                   "kappa": col("kappas").astype(np.float64), "cooldown_time": col("cooldown_times").astype(np.float64),
                   "sigma": sig, "warmup_time": col("warmup_times").astype(np.float64), "error_thresh": 1,
                   "tau": 0.01, "in_dims": 1,
                   "actor_config": {"layers": {4: "tanh", env_config["action_dim"]: "tanh"},
                                    "eta_h": col("lr_a_hs").astype(np.float64), "eta_l": col("lr_a_ls").astype(np.float64),
                                    "elig": list(col("elig_a"))},
                   "critic_config": {"layers": {4: "tanh", env_config["state_dim"]: "linear"},
                                     "eta_h": col("lr_c_hs").astype(np.float64), "eta_l": col("lr_c_ls").astype(np.float64),
                                     "elig": None},
                   "rls_config": {"state_dim": env_config["state_dim"], "action_dim": env_config["action_dim"],
                                  "rls_gamma": 1, "rls_cov": 10 ** 6}}
    env = Ce500ShortPeriod(env_config, batch=B, device=device, dtype=dtype)
    if weights is None:
        # functions.py:80,97,131-139: every config runs seeds 0..seeds-1, so seed s starts every config from the same
        # standard-normal draw, scaled by that config's own sigma
        from . import sp_engine
        weights = sp_engine.truncated_normal_weights(seeds, base_seed, sig, env._engine.device, repeat=n_configs)
    idhp = IDHPsp(env, idhp_config, verbose=0, seed=base_seed, weights=weights, log="full" if log_agents else None,
                  log_agents=log_agents)
    idhp.train()
    st = idhp.stats()
    metrics = []
    for c in range(n_configs):
        sl = slice(c * seeds, (c + 1) * seeds)
        div = st["diverged"][sl]
        ok = ~div
        conv = st["converged_time"][sl]
        m = {"diverged": int(div.sum()), "unsteady_convergence": int((conv > 30).sum()),                # functions.py:161-166
             "avg_c": float(np.around(float(st["sum_c"][sl][ok].mean()) if bool(ok.any()) else np.nan, 4)),
             "avg_t": float(np.around(float(conv.mean()), 4))}                                          # functions.py:178,182
        metrics.append(m)
    if log_agents:
        dt, t_end = env_config["dt"], env_config["t_end"]
        aoa_PSD, _ = get_PSD(t_end, dt, idhp.x_hist[:, :, 0])
        ref_PSD, _ = get_PSD(t_end, dt, idhp.ref_hist)
        psd_err = ((aoa_PSD - ref_PSD) ** 2).sum(dim=-1)                                               # functions.py:169-172
        for c in range(n_configs):
            lo, hi = c * seeds, min((c + 1) * seeds, log_agents)
            if lo < hi:
                e = psd_err[lo:hi]
                e = e[~torch.isnan(e)]
                metrics[c]["avg_PSD_err"] = float(np.around(float(e.mean()), 4)) if e.numel() else float("nan")
    return metrics, idhp


_ALGOS = {(0, None): "idhp", (0, "replacing"): "idhprt", (0, "accumulating"): "idhpat",
          (1, None): "midhp", (1, "replacing"): "midhprt", (1, "accumulating"): "midhpat"}       # functions.py:933,959-972


def MC_test_hparam(configs, directory, env, N, repetitions, save=0, show=0, transparency=0.2, *, noise=None, weights=None,
                   flight_step=5500, numpy2=True):
    """Nonlinear-task Monte-Carlo of functions.py:931-1060: ``N`` hyper-parameter sets (``configs`` = dict of lists with
    the reference's keys etaah, etaal, etach, etacl, lambda_hs, lambda_ls, seeds, ms, elig) x ``repetitions`` seeds, run
    as ONE batch of N * repetitions agents (agent index = config * repetitions + seed).

    ``env``: a batched ``Ce500NonLinear`` with ``batch == N * repetitions``.  As in the reference, repetition r of every
    configuration starts from the same seed-r initial weights and sees the same seed-r noise stream.  ``directory``,
    ``save``, ``show``, ``transparency`` are accepted for call compatibility (plotting / pickling are out of scope).
    ``weights`` (dict of (B, w) arrays) / ``noise`` ((steps, B) float32) replace the internally drawn ones; ``numpy2`` selects the
    NEP 50 promotion rules in `_adapt_check` (see IDHPnonlin).
    Returns a list of (algo, idhp_config, log) with log = the dict of functions.py:1008-1021 as torch tensors
    (angles in degrees, as stored there) plus ``'max_nz'``.
    """
    B = N * repetitions
    assert env.batch == B, f"env.batch must be N * repetitions = {B}"
    rep = lambda key, conv=float: np.repeat(np.asarray([conv(configs[key][i]) for i in range(N)]), repetitions)   # noqa: E731
    ms = rep("ms", int)
    elig = [configs["elig"][i] for i in range(N) for _ in range(repetitions)]
    n, m, mdp_s_dim = 3, 1, 4
    idhp_config = {"gamma": 0.6, "multistep": ms, "lr_decay": 0.998, "lambda_h": rep("lambda_hs"), "lambda_l": rep("lambda_ls"),
                   "kappa": [1, 2, 1], "cooldown_time": 2.0, "sigma": 0.1, "warmup_time": 4, "error_thresh": 1, "tau": 0.02,
                   "in_dims": mdp_s_dim,
                   "actor_config": {"layers": {10: "tanh", m: "tanh"}, "eta_h": rep("etaah"), "eta_l": rep("etaal"), "elig": elig},
                   "critic_config": {"layers": {10: "tanh", n: "linear"}, "eta_h": rep("etach"), "eta_l": rep("etacl"), "elig": 1233},
                   "rls_config": {"state_dim": n, "action_dim": m, "rls_gamma": 1, "rls_cov": 10 ** 6}}      # functions.py:973-997
    dev = env.device
    total_steps = int(env.t_end / env.dt)
    # seed r -> the same initial weights and noise stream for every configuration (functions.py:1013-1023)
    g = torch.Generator(device=dev); g.manual_seed(0)

    def draw(w):
        out = torch.randn((repetitions, w), generator=g, device=dev, dtype=torch.float32)
        bad = out.abs() > 2.0
        while bool(bad.any()):
            out = torch.where(bad, torch.randn((repetitions, w), generator=g, device=dev, dtype=torch.float32), out)
            bad = out.abs() > 2.0
        return (out * 0.1).double().repeat(N, 1)
    if weights is None:
        weights = {"W1a": draw(40), "W2a": draw(10), "W1c": draw(40), "W2c": draw(30)}
    if noise is None:
        noise = torch.randn((total_steps, repetitions), generator=g, device=dev, dtype=torch.float32).repeat(1, N)
    env._engine.set_hpi("FLIGHT_STEP", int(flight_step))
    idhp = IDHPnonlin(env, idhp_config, verbose=0, seed=0, weights=weights, log="mc", log_agents=B, numpy2=numpy2)
    idhp.train(noise=noise)
    st, lg, dt = idhp.stats(), idhp.log, env.dt
    flight = lg["theta"][:, flight_step:]
    t_flight = (total_steps - flight_step) * dt                                                 # 35 s (functions.py:1033)
    theta_PSD, omega = get_PSD(t_flight, dt, torch.rad2deg(flight))
    Sm = (torch.atleast_2d(theta_PSD) * omega.to(dev)).sum(dim=-1)                              # functions.py:1034
    nrm = lambda v: v / v.max(dim=-1, keepdim=True).values                                      # noqa: E731
    full = {"RSE": torch.stack([st["rse"][:, 0] - st["rse_flight"][:, 0], st["rse_flight"][:, 0]], dim=-1),   # :1036-1037
            "e": torch.rad2deg(lg["e"]), "theta": torch.rad2deg(lg["theta"]), "alpha": torch.rad2deg(lg["alpha"]),
            "q": torch.rad2deg(lg["q"]), "V": lg["v"], "h": lg["h"], "action_cmd": torch.rad2deg(lg["a_cmd"]),
            "action_eff": torch.rad2deg(lg["a_eff"]),
            "n_z": torch.div(lg["v"] * lg["q"], torch.full_like(lg["v"], 9.80665)),   # a true division (scalar divisors become reciprocal multiplies)
            "wa_norm": nrm(lg["wa_norm"]), "wc_norm": nrm(lg["wc_norm"]), "Sm": Sm.unsqueeze(-1), "rls_eps": lg["rls_eps"],
            "max_nz": st["nz_peak"]}
    out = []
    for i in range(N):
        sl = slice(i * repetitions, (i + 1) * repetitions)
        cfg_i = {**idhp_config, "multistep": int(ms[sl][0]), "lambda_h": float(idhp_config["lambda_h"][sl][0]),
                 "lambda_l": float(idhp_config["lambda_l"][sl][0]),
                 "actor_config": {**idhp_config["actor_config"], "eta_h": float(idhp_config["actor_config"]["eta_h"][sl][0]),
                                  "eta_l": float(idhp_config["actor_config"]["eta_l"][sl][0]), "elig": configs["elig"][i]},
                 "critic_config": {**idhp_config["critic_config"], "eta_h": float(idhp_config["critic_config"]["eta_h"][sl][0]),
                                   "eta_l": float(idhp_config["critic_config"]["eta_l"][sl][0])}}
        out.append((_ALGOS[(int(ms[sl][0]), configs["elig"][i])], cfg_i, {k: v[sl] for k, v in full.items()}))
    return out
