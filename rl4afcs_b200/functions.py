"""Batched Monte-Carlo front-end and episode statistics -- the hot-path part of functions.py / utils.py of
wingos80/RL4AFCS (plotting and pickling are out of scope).

  get_PSD, get_convergence_time     utils.py:188-236, 350-369     (torch, batched over agents)
  MC_run_seed                       functions.py:39-60            (the per-run output dict, for a whole batch)
  MC_run                            functions.py:62-232           (configs x seeds flattened onto the agent axis; returns
                                                                   the metrics dict of functions.py:223-227 per config)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .envs.linear.env import Ce500ShortPeriod
from .objects import IDHPsp


def get_PSD(t_end, dt, array):
    """Power spectral density |FFT|^2 / t_end of each row, first N/2 bins, and the frequency axis (utils.py:188-236)."""
    fs = 1 / dt
    N = int(t_end * fs)
    upp = int(N / 2)
    omega = torch.arange(0, upp, 1, dtype=torch.float64) / (N * dt)
    x = torch.as_tensor(array)
    if x.ndim == 1:
        x = x[None]
    f = torch.fft.fft(x.to(torch.float64), dim=-1)
    spectra = (f * torch.conj(f)).abs()[..., :upp] / t_end
    return spectra.squeeze(0) if spectra.shape[0] == 1 else spectra, omega


def get_convergence_time(c_hist, kappa, dt):
    """Time of the last sample whose angle-of-attack error exceeds 0.5 deg (utils.py:350-369); batched over the
    leading axis.  (The fused kernel computes the same quantity in-register: ``stats()['converged_time']``.)"""
    c = torch.as_tensor(c_hist, dtype=torch.float64)
    aoa_error = torch.rad2deg(torch.sqrt(-2 * (c / kappa)))
    over = aoa_error > 0.5
    idx = torch.arange(c.shape[-1], device=c.device)
    last = torch.where(over, idx, torch.full_like(idx, -1)).max(dim=-1).values
    return last.to(torch.float64) * dt


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


def MC_run(n_configs, configs, env_config, seeds, *, device="cuda", dtype="mixed", base_seed=0, log_agents=0):
    """Monte-Carlo over ``n_configs`` hyper-parameter sets x ``seeds`` seeds in ONE batch (functions.py:62-232 ran
    a process pool of 10).  ``configs`` has the reference's keys (lambda_hs, lambda_ls, kappas, cooldown_times, sigmas,
    warmup_times, elig_a, lr_a_hs, lr_a_ls, lr_c_hs, lr_c_ls, multistep), each None or a list per config.
    Returns (metrics per config, idhp): metrics = dict(avg_PSD_err?, avg_c, avg_t, diverged, unsteady_convergence)
    as functions.py:223-227 (PSD error only when trajectories are logged)."""
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
                   "sigma": float(sig[0]), "warmup_time": col("warmup_times").astype(np.float64), "error_thresh": 1,
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
    idhp = IDHPsp(env, idhp_config, verbose=0, seed=base_seed, log="full" if log_agents else None, log_agents=log_agents)
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
             "avg_t": float(np.around(float(conv[ok].mean()) if bool(ok.any()) else np.nan, 4))}
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
