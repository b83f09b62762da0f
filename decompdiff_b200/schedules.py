"""Noise schedules and posterior coefficient tables of DecompDiff.

Restates /root/reference/models/decompdiff.py:96-131 (position diffusion) and
/root/reference/models/transitions.py:12-28,31-61,97-120 (categorical diffusion) in float64 numpy,
cast to float32 at the end exactly like the reference's `to_torch_const`.
"""
from __future__ import annotations

from typing import Dict

import numpy as np


def cosine_alpha_sqrt_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """sqrt of the per-step alphas of the cosine schedule (transitions.py:12-28)."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    alphas = np.clip(alphas_cumprod[1:] / alphas_cumprod[:-1], a_min=0.001, a_max=1.)
    return np.sqrt(alphas)


def beta_schedule(kind: str, beta_start: float, beta_end: float, timesteps: int) -> np.ndarray:
    """transitions.py:31-61."""
    if kind == 'quad':
        return np.linspace(beta_start ** 0.5, beta_end ** 0.5, timesteps, dtype=np.float64) ** 2
    if kind == 'linear':
        return np.linspace(beta_start, beta_end, timesteps, dtype=np.float64)
    if kind == 'const':
        return beta_end * np.ones(timesteps, dtype=np.float64)
    if kind == 'jsd':
        return 1.0 / np.linspace(timesteps, 1, timesteps, dtype=np.float64)
    if kind == 'sigmoid':
        b = np.linspace(-6, 6, timesteps)
        return 1 / (np.exp(-b) + 1) * (beta_end - beta_start) + beta_start
    raise NotImplementedError(kind)


def position_tables(cfg) -> Dict[str, np.ndarray]:
    T = cfg.num_diffusion_timesteps
    if cfg.beta_schedule == 'cosine':
        alphas = cosine_alpha_sqrt_schedule(T, cfg.pos_beta_s) ** 2
        betas = 1. - alphas
    else:
        betas = beta_schedule(cfg.beta_schedule, cfg.beta_start, cfg.beta_end, T)
        alphas = 1. - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1., ac[:-1])
    post_var = betas * (1. - ac_prev) / (1. - ac)
    out = {
        'betas': betas, 'alphas_cumprod': ac, 'alphas_cumprod_prev': ac_prev,
        'sqrt_alphas_cumprod': np.sqrt(ac), 'sqrt_one_minus_alphas_cumprod': np.sqrt(1. - ac),
        'sqrt_recip_alphas_cumprod': np.sqrt(1. / ac), 'sqrt_recipm1_alphas_cumprod': np.sqrt(1. / ac - 1),
        'posterior_mean_c0_coef': betas * np.sqrt(ac_prev) / (1. - ac),
        'posterior_mean_ct_coef': (1. - ac_prev) * np.sqrt(alphas) / (1. - ac),
        'posterior_var': post_var,
    }
    out = {k: v.astype(np.float32) for k, v in out.items()}
    # decompdiff.py:130 takes the log of the already-fp32 variance, entry 0 := entry 1
    pv = out['posterior_var']
    out['posterior_logvar'] = np.log(np.append(pv[1], pv[1:])).astype(np.float32)
    out['pos_score_coef'] = (betas / np.sqrt(alphas)).astype(np.float32)
    return out


def categorical_tables(timesteps: int, s: float, num_classes: int, prior_probs=None) -> Dict[str, np.ndarray]:
    """DiscreteTransition.__init__ (transitions.py:98-120)."""
    log_alphas = np.log(cosine_alpha_sqrt_schedule(timesteps, s))
    log_cum = np.cumsum(log_alphas)
    l1m = lambda a: np.log(1 - np.exp(a) + 1e-40)
    if prior_probs is None:
        prior = -np.log(num_classes).repeat(num_classes)[None, :]
    else:
        prior = np.log(np.asarray(prior_probs).clip(min=1e-30))
    out = {
        'log_alphas_v': log_alphas, 'log_one_minus_alphas_v': l1m(log_alphas),
        'log_alphas_cumprod_v': log_cum, 'log_one_minus_alphas_cumprod_v': l1m(log_cum),
        'prior_probs': prior,
    }
    return {k: np.asarray(v).astype(np.float32) for k, v in out.items()}
