"""Noise-schedule tables the samplers consume, built on the CPU in fp32 exactly as the reference
builds them at model construction (they are attributes, not buffers, so checkpoints do not carry
them: models/Diffusion_model/NsDiff/NsDiff_net.py:92-137, TMDM/TMDM.py:47-58).

The fp32 rounding of these tables is part of the reference's results -- e.g. betas_tilde is the
difference of two O(T) sums -- so nothing here is evaluated in higher precision.
"""
import math

import torch

# Row order of the packed NsDiff table = argument order of p_sample_loop (nsdiff_utils.py:271).
NSDIFF_ROWS = ("alphas", "one_minus_alphas_bar_sqrt", "alphas_cumprod", "alphas_cumprod_sum", "alphas_cumprod_prev",
               "alphas_cumprod_sum_prev", "betas_tilde", "betas_bar", "betas_tilde_m_1", "betas_bar_m_1")
TMDM_ROWS = ("alphas", "one_minus_alphas_bar_sqrt")


def beta_schedule(schedule, num_timesteps, start, end):
    """All schedule names the reference accepts (nsdiff_utils.py:6-32)."""
    n = num_timesteps
    if schedule == "linear":
        betas = torch.linspace(start, end, n)
    elif schedule == "const":
        betas = end * torch.ones(n)
    elif schedule == "quad":
        betas = torch.linspace(start ** 0.5, end ** 0.5, n) ** 2
    elif schedule == "jsd":
        betas = 1.0 / torch.linspace(n, 1, n)
    elif schedule == "sigmoid":
        betas = torch.sigmoid(torch.linspace(-6, 6, n)) * (end - start) + start
    elif schedule in ("cosine", "cosine_reverse"):
        def f(i):
            return math.cos((i / n + 0.008) / 1.008 * math.pi / 2) ** 2
        betas = torch.tensor([min(1 - f(i + 1) / f(i), 0.999) for i in range(n)])
        if schedule == "cosine_reverse":
            betas = betas.flip(0)
    elif schedule == "cosine_anneal":
        betas = torch.tensor([start + 0.5 * (end - start) * (1 - math.cos(t / (n - 1) * math.pi)) for t in range(n)])
    else:
        raise ValueError("unknown diffusion schedule {!r}".format(schedule))
    return betas.float()


def _windowed_products(alphas, weight_by_alpha):
    """out[t] = sum_{k<=t} (prod_{j=k..t} alpha_j) [* alpha_k]: alpha-tilde / alpha-hat of NsDiff."""
    out = torch.zeros_like(alphas)
    for t in range(alphas.shape[0]):
        rev = alphas[: t + 1].flip(dims=[0])
        terms = torch.cumprod(rev, dim=0)
        if weight_by_alpha:
            terms = terms * rev
        out[t] = terms.sum()
    return out


def nsdiff_tables(diffusion_schedule="linear", diffusion_steps=20, beta_start=1e-4, beta_end=0.02):
    betas = beta_schedule(diffusion_schedule, diffusion_steps, beta_start, beta_end)
    alphas = 1.0 - betas
    acp = alphas.cumprod(dim=0)
    a_tilde = _windowed_products(alphas, False)
    a_hat = _windowed_products(alphas, True)
    b_tilde = a_tilde - a_hat
    b_bar = 1 - acp
    if not bool((b_tilde >= 0).all()) or not bool(((b_bar - b_tilde) >= 0).all()):
        raise AssertionError("schedule violates betas_tilde >= 0 / betas_bar >= betas_tilde")  # NsDiff_net.py:112-114
    om = torch.sqrt(1 - acp)
    if diffusion_schedule == "cosine":
        om = om * 0.9999
    one = torch.ones(1)
    t = {
        "betas": betas, "alphas": alphas, "one_minus_alphas_bar_sqrt": om, "alphas_cumprod": acp,
        "alphas_cumprod_sum": a_tilde, "alphas_hat": a_hat,
        "alphas_cumprod_prev": torch.cat([one, acp[:-1]]), "alphas_cumprod_sum_prev": torch.cat([one, a_tilde[:-1]]),
        "betas_tilde": b_tilde, "betas_bar": b_bar,
        "betas_tilde_m_1": torch.cat([one, b_tilde[:-1]]), "betas_bar_m_1": torch.cat([one, b_bar[:-1]]),
    }
    return t


def tmdm_tables(beta_schedule_name="linear", timesteps=20, beta_start=1e-4, beta_end=0.02):
    betas = beta_schedule(beta_schedule_name, timesteps, beta_start, beta_end)
    alphas = 1.0 - betas
    om = torch.sqrt(1 - alphas.cumprod(dim=0))
    if beta_schedule_name == "cosine":
        om = om * 0.9999
    return {"betas": betas, "alphas": alphas, "one_minus_alphas_bar_sqrt": om}


def stack_rows(tables, rows):
    return torch.stack([tables[r].float() for r in rows]).contiguous()
