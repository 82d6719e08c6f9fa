"""Host-side helpers of utils.py of wingos80/RL4AFCS that the Monte-Carlo front-ends need (plot helpers are out of scope).

  get_PSD, get_convergence_time          utils.py:188-236, 350-369   batched over a leading axis, torch (CPU or CUDA tensors)
  pick_continuous_hparams                utils.py:238-291            hyper-parameter sampler, continuous ranges
  pick_discrete_hparams                  utils.py:293-348            hyper-parameter sampler, discrete sets
  VD_A                                   utils.py:391-434            Vargha-Delaney A effect size
  kl_divergence                          utils.py:436-452

The samplers draw from numpy's legacy global generator in the reference's order (one ``uniform`` / ``choice`` call per
supplied range, in the argument order below), so ``true_random=False`` (seed 0) gives the reference's values.
"""
from __future__ import annotations

import os
import time
from bisect import bisect_left

import numpy as np
import torch

_CONT_ORDER = ("lambda_hs", "lambda_ls", "lr_a_hs", "lr_c_hs", "lr_a_ls", "lr_c_ls", "kappas", "cooldown_times", "sigmas",
               "warmup_times")


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


def _seed(true_random):
    np.random.seed((os.getpid() * int(time.time())) % 123456 if true_random else 0)      # utils.py:250-253


def pick_continuous_hparams(n_configs, lambda_hs=None, lambda_ls=None, lr_a_hs=None, lr_c_hs=None, lr_a_ls=None, lr_c_ls=None,
                            kappas=None, cooldown_times=None, sigmas=None, warmup_times=None, elig_a=None, true_random=True):
    """Each range is ``[low, high]``; returns {name: list of n_configs values or None}.  Values are rounded to 3
    decimals, ``kappas`` truncated to int, ``multistep`` is all zeros (utils.py:266-291)."""
    _seed(true_random)
    given = dict(lambda_hs=lambda_hs, lambda_ls=lambda_ls, lr_a_hs=lr_a_hs, lr_c_hs=lr_c_hs, lr_a_ls=lr_a_ls, lr_c_ls=lr_c_ls,
                 kappas=kappas, cooldown_times=cooldown_times, sigmas=sigmas, warmup_times=warmup_times)
    configs = {"multistep": [0] * n_configs}
    for key in _CONT_ORDER:
        rng = given[key]
        if not rng:
            configs[key] = None
            continue
        draw = np.random.uniform(rng[0], rng[1], n_configs)
        configs[key] = [int(v) for v in draw] if key == "kappas" else [round(v, 3) for v in draw]
    configs["elig_a"] = [round(elig_a, 3) if isinstance(elig_a, float) else elig_a for _ in range(n_configs)] if elig_a else None
    return configs


def pick_discrete_hparams(n_configs, lambda_hs=None, lambda_ls=None, lr_a_hs=None, lr_c_hs=None, lr_a_ls=None, lr_c_ls=None,
                          kappas=None, cooldown_times=None, sigmas=None, warmup_times=None, elig_a=None, lr_decays=None,
                          true_random=True):
    """Each argument is a list of candidate values; returns {name: array of n_configs picks or None} plus
    ``seeds`` = randint(0, 10000) per config (utils.py:312-338)."""
    _seed(true_random)
    given = dict(lambda_hs=lambda_hs, lambda_ls=lambda_ls, lr_a_hs=lr_a_hs, lr_c_hs=lr_c_hs, lr_a_ls=lr_a_ls, lr_c_ls=lr_c_ls,
                 kappas=kappas, cooldown_times=cooldown_times, sigmas=sigmas, warmup_times=warmup_times, elig_a=elig_a,
                 lr_decays=lr_decays)
    configs = {}
    for key in (*_CONT_ORDER, "elig_a", "lr_decays"):
        configs[key] = np.random.choice(given[key], n_configs) if given[key] else None
    configs["seeds"] = np.random.randint(0, 10000, n_configs)
    return configs


def VD_A(X, Y):
    """Vargha-Delaney A of treatment ``X`` against control ``Y`` and its magnitude label (utils.py:391-434).
    A = (2 R1 - m (m + 1)) / (2 n m) with R1 the rank sum of X in the pooled sample (mid-ranks for ties)."""
    X, Y = list(X), list(Y)
    m, n = len(X), len(Y)
    if m != n:                                            # the reference truncates the longer list
        k = min(m, n)
        X, Y = X[:k], Y[:k]
    pooled = np.asarray(Y + X, dtype=np.float64)
    order = np.argsort(pooled, kind="mergesort")
    ranks = np.empty(pooled.size)
    srt = pooled[order]
    i = 0
    while i < srt.size:                                   # mid-ranks (scipy.stats.rankdata 'average')
        j = i
        while j + 1 < srt.size and srt[j + 1] == srt[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    r1 = float(ranks[:m].sum())                           # utils.py:420 sums the FIRST m ranks of (Y + X)
    A = (2 * r1 - m * (m + 1)) / (2 * n * m)
    labels = ["negligible", "small", "medium", "large"]
    return A, labels[bisect_left([0.06, 0.14, 0.21], abs(A - 0.5))]


def kl_divergence(u, v, epsilon=np.finfo(float).eps):
    """sum u log(u / v) with non-positive entries replaced by ``epsilon`` (utils.py:436-452)."""
    u = np.where(np.asarray(u) <= 0, epsilon, u)
    v = np.where(np.asarray(v) <= 0, epsilon, v)
    return np.sum(u * np.log(u / v))
