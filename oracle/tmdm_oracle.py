"""CPU oracle (test infrastructure, never shipped): TMDM reverse-diffusion sampler.

Functional torch-CPU fp32 restatement of models/Diffusion_model/TMDM/tmdm_diffusion_utils.py:57-119
and tmdm_model.py:23-64 (cat_x = cat_y_pred = True as in TMDM/tmdm.yml:21-22, so the network
input is cat(y_t, y_0_hat) and the DataEmbedding of x computed in TMDM.py:95 is dead).
State-dict keys as in the reference: ``model.diffussion_model.lin{1,2,3}.{lin,embed}`` / ``lin4``.
"""
import torch
import torch.nn.functional as F

from .nsdiff_oracle import DENOISER_PREFIX, linear, make_beta_schedule, tile_rows  # noqa: F401


def tmdm_schedule(beta_schedule="linear", timesteps=20, beta_start=1e-4, beta_end=0.02):
    """TMDM.py:47-58: only ``alphas`` and ``one_minus_alphas_bar_sqrt`` reach the sampler."""
    betas = make_beta_schedule(beta_schedule, timesteps, beta_start, beta_end).float()
    alphas = 1.0 - betas
    om = torch.sqrt(1 - alphas.cumprod(dim=0))
    if beta_schedule == "cosine":
        om = om * 0.9999
    return {"alphas": alphas, "one_minus_alphas_bar_sqrt": om}


def denoiser_forward(sd, y_t, y_0_hat, t):
    """tmdm_model.py:39-64: three (Linear * embed[t]) -> softplus blocks (no normalise), then lin4."""
    h = torch.cat((y_t, y_0_hat), dim=-1)
    for name in ("lin1", "lin2", "lin3"):
        out = linear(h, sd[DENOISER_PREFIX + name + ".lin.weight"], sd[DENOISER_PREFIX + name + ".lin.bias"])
        h = F.softplus(sd[DENOISER_PREFIX + name + ".embed.weight"][t].view(1, 1, -1) * out)
    return linear(h, sd[DENOISER_PREFIX + "lin4.weight"], sd[DENOISER_PREFIX + "lin4.bias"])


def _sqrt_alpha_bar(sched, t):
    return (1 - sched["one_minus_alphas_bar_sqrt"][t].square()).sqrt()


def p_sample(sd, sched, y, y_0_hat, y_T_mean, t, z):
    """tmdm_diffusion_utils.py:57-91 (z is drawn before the network call, :69)."""
    alpha_t = sched["alphas"][t]
    s1m = sched["one_minus_alphas_bar_sqrt"][t]
    s1m_prev = sched["one_minus_alphas_bar_sqrt"][t - 1]
    sab = _sqrt_alpha_bar(sched, t)
    sab_prev = _sqrt_alpha_bar(sched, t - 1)
    gamma_0 = (1 - alpha_t) * sab_prev / (s1m.square())
    gamma_1 = (s1m_prev.square()) * (alpha_t.sqrt()) / (s1m.square())
    gamma_2 = 1 + (sab - 1) * (alpha_t.sqrt() + sab_prev) / (s1m.square())
    eps_theta = denoiser_forward(sd, y, y_0_hat, t)
    y_0 = 1 / sab * (y - (1 - sab) * y_T_mean - eps_theta * s1m)
    mean = gamma_0 * y_0 + gamma_1 * y + gamma_2 * y_T_mean
    beta_t_hat = (s1m_prev.square()) / (s1m.square()) * (1 - alpha_t)
    return mean + beta_t_hat.sqrt() * z


def p_sample_t_1to0(sd, sched, y, y_0_hat, y_T_mean):
    """tmdm_diffusion_utils.py:94-104."""
    s1m = sched["one_minus_alphas_bar_sqrt"][0]
    sab = _sqrt_alpha_bar(sched, 0)
    eps_theta = denoiser_forward(sd, y, y_0_hat, 0)
    return 1 / sab * (y - (1 - sab) * y_T_mean - eps_theta * s1m)


def p_sample_loop(sd, sched, y_0_hat, y_T_mean, n_steps, draw):
    """tmdm_diffusion_utils.py:107-119: unit-variance prior around y_T_mean, n_steps draws."""
    cur = draw(y_T_mean) + y_T_mean
    seq = [cur]
    for t in reversed(range(1, n_steps)):
        cur = p_sample(sd, sched, cur, y_0_hat, y_T_mean, t, draw(cur))
        seq.append(cur)
    seq.append(p_sample_t_1to0(sd, sched, seq[-1], y_0_hat, y_T_mean))
    return seq


def evaluation_step(sd, net_param, y_0_hat, sched=None, draw=None):
    """tmdm_adapter.py:116-155 with the condition mean y_0_hat [B, label_len+pred_len, F] supplied."""
    T = net_param["diffusion_steps"]
    K = net_param["n_z_samples"]
    S = min(int(net_param["parallel_sample"]), K)
    if K % S != 0:
        raise ValueError("n_z_samples must be divisible by parallel_sample")
    pred_len = net_param["pred_len"]
    if sched is None:
        sched = tmdm_schedule(net_param.get("beta_schedule", "linear"), T,
                              net_param.get("beta_start", 1e-4), net_param.get("beta_end", 0.02))
    if draw is None:
        draw = torch.randn_like
    b, rows, nf = y_0_hat.shape
    preds = []
    for _ in range(K // S):
        tile = tile_rows(y_0_hat, S)
        seq = p_sample_loop(sd, sched, tile, tile, T, draw)
        preds.append(seq[T].reshape(b, S, rows, nf)[:, :, -pred_len:, :])
    return torch.cat(preds, dim=1).permute(0, 2, 3, 1)
