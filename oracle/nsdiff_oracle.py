"""CPU oracle (test infrastructure, never shipped): NsDiff reverse-diffusion sampler.

Functional torch-CPU fp32 restatement of the reference algorithm.  Every function cites
the reference lines it follows (paths relative to the reference root).  The op ORDER of
every fp32 expression is kept as in the reference so that, fed the same noise, the oracle
reproduces the reference to the last bit on the same torch build (checked by
tests/test_oracle_golden.py against fixtures made by oracle/make_golden.py).

Weights are passed as a plain ``dict[str, Tensor]`` using the reference's state-dict key
names (e.g. ``model.diffussion_model.lin1.lin.weight``; the double "s" is the reference's).
"""
import math

import torch
import torch.nn.functional as F

EPS = 10e-8  # models/Diffusion_model/NsDiff/nsdiff_utils.py:5 and NsDiff_model.py:37 (== 1e-7)

DENOISER_PREFIX = "model.diffussion_model."

SCHEDULE_KEYS = (
    "alphas",
    "one_minus_alphas_bar_sqrt",
    "alphas_cumprod",
    "alphas_cumprod_sum",
    "alphas_cumprod_prev",
    "alphas_cumprod_sum_prev",
    "betas_tilde",
    "betas_bar",
    "betas_tilde_m_1",
    "betas_bar_m_1",
)


def make_beta_schedule(schedule="linear", num_timesteps=1000, start=1e-5, end=1e-2):
    """nsdiff_utils.py:6-32 (same table in TMDM/tmdm_diffusion_utils.py:6-32)."""
    if schedule == "linear":
        return torch.linspace(start, end, num_timesteps)
    if schedule == "const":
        return end * torch.ones(num_timesteps)
    if schedule == "quad":
        return torch.linspace(start ** 0.5, end ** 0.5, num_timesteps) ** 2
    if schedule == "jsd":
        return 1.0 / torch.linspace(num_timesteps, 1, num_timesteps)
    if schedule == "sigmoid":
        betas = torch.linspace(-6, 6, num_timesteps)
        return torch.sigmoid(betas) * (end - start) + start
    if schedule in ("cosine", "cosine_reverse"):
        max_beta, s = 0.999, 0.008
        vals = []
        for i in range(num_timesteps):
            hi = math.cos(((i + 1) / num_timesteps + s) / (1 + s) * math.pi / 2) ** 2
            lo = math.cos((i / num_timesteps + s) / (1 + s) * math.pi / 2) ** 2
            vals.append(min(1 - hi / lo, max_beta))
        betas = torch.tensor(vals)
        return betas.flip(0) if schedule == "cosine_reverse" else betas
    if schedule == "cosine_anneal":
        return torch.tensor(
            [start + 0.5 * (end - start) * (1 - math.cos(t / (num_timesteps - 1) * math.pi))
             for t in range(num_timesteps)]
        )
    raise ValueError("unknown schedule {!r}".format(schedule))


def _tilde_alpha(alpha):
    """NsDiff_net.py:34-43: alpha_tilde[t] = sum_k prod_{j=t-k..t} alpha_j (fp32, O(T^2))."""
    alpha = alpha.float()
    out = torch.zeros_like(alpha)
    for t in range(alpha.shape[0]):
        out[t] = torch.cumprod(alpha[: t + 1].flip(dims=[0]), dim=0).sum()
    return out


def _hat_alpha(alpha):
    """NsDiff_net.py:45-54: like _tilde_alpha with every term multiplied by its own alpha again."""
    alpha = alpha.float()
    out = torch.zeros_like(alpha)
    for t in range(alpha.shape[0]):
        rev = alpha[: t + 1].flip(dims=[0])
        out[t] = (torch.cumprod(rev, dim=0) * rev).sum()
    return out


def nsdiff_schedule(diffusion_schedule="linear", diffusion_steps=20, beta_start=1e-4, beta_end=0.02):
    """The ten per-step fp32 tables p_sample_loop consumes (NsDiff_net.py:92-137).

    Note ``betas_tilde = alphas_tilde - alphas_hat`` cancels two numbers of size ~T in
    fp32 (NsDiff_net.py:109); that rounding is part of the reference's results, so the
    tables are built exactly as there and never recomputed in higher precision.
    """
    betas = make_beta_schedule(diffusion_schedule, diffusion_steps, beta_start, beta_end).float()
    alphas = 1.0 - betas
    alphas_cumprod = alphas.cumprod(dim=0)
    betas_bar = 1 - alphas_cumprod
    alphas_cumprod_sum = _tilde_alpha(alphas)
    alphas_hat = _hat_alpha(alphas)
    betas_tilde = alphas_cumprod_sum - alphas_hat
    one = torch.ones(1)
    one_minus_alphas_bar_sqrt = torch.sqrt(1 - alphas_cumprod)
    if diffusion_schedule == "cosine":
        one_minus_alphas_bar_sqrt = one_minus_alphas_bar_sqrt * 0.9999
    return {
        "betas": betas,
        "alphas": alphas,
        "one_minus_alphas_bar_sqrt": one_minus_alphas_bar_sqrt,
        "alphas_cumprod": alphas_cumprod,
        "alphas_cumprod_sum": alphas_cumprod_sum,
        "alphas_hat": alphas_hat,
        "alphas_cumprod_prev": torch.cat([one, alphas_cumprod[:-1]]),
        "alphas_cumprod_sum_prev": torch.cat([one, alphas_cumprod_sum[:-1]]),
        "betas_tilde": betas_tilde,
        "betas_bar": betas_bar,
        "betas_tilde_m_1": torch.cat([one, betas_tilde[:-1]]),
        "betas_bar_m_1": torch.cat([one, betas_bar[:-1]]),
    }


def linear(x, w, b):
    """nn.Linear as the reference's modules execute it.  ATen picks its matmul folding from
    ``requires_grad`` of the weight (still True for an nn.Parameter under no_grad), and the two
    foldings round differently, so the oracle presents the weights the same way."""
    with torch.no_grad():
        return F.linear(x, w.detach().requires_grad_(True), b.detach().requires_grad_(True))


def _cond_linear(sd, name, x, t):
    """denoise.py:14-20: embed[t] * (x @ W^T + b); one t for the whole batch."""
    out = linear(x, sd[DENOISER_PREFIX + name + ".lin.weight"], sd[DENOISER_PREFIX + name + ".lin.bias"])
    gamma = sd[DENOISER_PREFIX + name + ".embed.weight"][t]
    return gamma.view(1, 1, -1) * out


def denoiser_forward(sd, y_t, y_0_hat, gx, t):
    """denoise.py:35-51 via NsDiff_net.py:163-172.  Returns (eps_theta, sigma_theta), both [R,O,F].

    Three ConditionalLinear(->128) + softplus + L2-normalise blocks, then two heads reading the
    same 128-vector: eps = lin4(h); sigma = softplus(sigma_lin(softplus(h))).
    """
    h = torch.cat((y_t, y_0_hat, gx), dim=-1)
    for name in ("lin1", "lin2", "lin3"):
        h = F.normalize(F.softplus(_cond_linear(sd, name, h, t)), dim=-1)
    eps = linear(h, sd[DENOISER_PREFIX + "lin4.weight"], sd[DENOISER_PREFIX + "lin4.bias"])
    sig = F.softplus(linear(F.softplus(h), sd[DENOISER_PREFIX + "sigma_lin.weight"],
                            sd[DENOISER_PREFIX + "sigma_lin.bias"]))
    return eps, sig


def _sigma_y0_and_noise(sched, t, gx, sigma_theta):
    """nsdiff_utils.py:139-147 (and :225-230): quadratic estimate of Sigma_Y0, forward noise."""
    a = sched["alphas"][t]
    bt = sched["betas_tilde"][t]
    bb = sched["betas_bar"][t]
    btm = sched["betas_tilde_m_1"][t]
    bbm = sched["betas_bar_m_1"][t]
    lambda_0 = a * (1 - a) * btm
    lambda_1 = ((1 - a) ** 2 * btm + a * (1 - a) * (bbm - btm)) * gx - sigma_theta * (a * btm + a * (1 - a))
    lambda_2 = gx ** 2 * (1 - a) ** 2 * (bbm - btm) - sigma_theta * gx * (a * bbm - a * btm + (1 - a) ** 2)
    sigma_y0_hat = (-lambda_1 + ((lambda_1) ** 2 - 4 * lambda_0 * lambda_2).sqrt()) / (2 * lambda_0)
    noise = (bb - bt) * gx + bt * sigma_y0_hat
    return sigma_y0_hat, noise


def _y0_reparam(sched, t, y, y_T_mean, eps_theta, noise):
    """nsdiff_utils.py:135,151-152: sqrt(alpha_bar) is re-derived from one_minus_alphas_bar_sqrt."""
    s1m = sched["one_minus_alphas_bar_sqrt"][t]
    sqrt_alpha_bar_t = (1 - s1m.square()).sqrt()
    return 1 / sqrt_alpha_bar_t * (y - (1 - sqrt_alpha_bar_t) * y_T_mean - eps_theta * torch.sqrt(noise))


def _gammas(sched, t, gx, y_sigma):
    """nsdiff_utils.py:40-56 (cal_sigma12) + :80-92 (calc_gammas)."""
    at = sched["alphas"][t]
    btm = sched["betas_tilde_m_1"][t]
    bbm = sched["betas_bar_m_1"][t]
    Sigma_1 = (1 - at) ** 2 * gx + at * (1 - at) * y_sigma
    Sigma_2 = (bbm - btm) * gx + btm * y_sigma
    sqrt_alpha_t = at.sqrt()
    sqrt_alpha_bar_t_m_1 = sched["alphas_cumprod_prev"][t].sqrt()
    den = at * Sigma_2 + Sigma_1
    gamma_0 = sqrt_alpha_bar_t_m_1 * Sigma_1 / den
    gamma_1 = sqrt_alpha_t * Sigma_2 / den
    gamma_2 = ((sqrt_alpha_t * (at - 1)) * Sigma_2 + (1 - sqrt_alpha_bar_t_m_1) * Sigma_1) / den
    return gamma_0, gamma_1, gamma_2


def p_sample(sd, sched, y, y_0_hat, gx, y_T_mean, t, z):
    """nsdiff_utils.py:111-158: one reverse step y_t -> y_{t-1}; ``z`` is the N(0,1) draw (:130)."""
    eps_theta, sigma_theta = denoiser_forward(sd, y, y_0_hat, gx, t)
    sigma_y0_hat, noise = _sigma_y0_and_noise(sched, t, gx, sigma_theta)
    y_0 = _y0_reparam(sched, t, y, y_T_mean, eps_theta, noise)
    g0, g1, g2 = _gammas(sched, t, gx, sigma_y0_hat)
    return g0 * y_0 + g1 * y + g2 * y_T_mean + torch.sqrt(sigma_theta) * z


def p_sample_t_1to0(sd, sched, y, y_0_hat, gx, y_T_mean):
    """nsdiff_utils.py:209-239: the last step (t index 0) returns y_0 reparam, no noise."""
    eps_theta, sigma_theta = denoiser_forward(sd, y, y_0_hat, gx, 0)
    _, noise = _sigma_y0_and_noise(sched, 0, gx, sigma_theta)
    return _y0_reparam(sched, 0, y, y_T_mean, eps_theta, noise)


def p_sample_loop(sd, sched, y_0_hat, gx, y_T_mean, n_steps, draw):
    """nsdiff_utils.py:271-284.  ``draw(like)`` supplies each N(0,1) tensor in reference order:
    one for y_T (:273) then one per t = T-1 .. 1 (:130) -- exactly n_steps draws.
    Returns the list of n_steps+1 tensors (y_T, ..., y_0)."""
    cur = torch.sqrt(gx) * draw(y_T_mean) + y_T_mean
    seq = [cur]
    for t in reversed(range(1, n_steps)):
        # the reference draws z after the network call; the network draws nothing in eval
        cur = p_sample(sd, sched, cur, y_0_hat, gx, y_T_mean, t, draw(cur))
        seq.append(cur)
    assert len(seq) == n_steps
    seq.append(p_sample_t_1to0(sd, sched, seq[-1], y_0_hat, gx, y_T_mean))
    return seq


def tile_rows(x, repeat_n):
    """NsDiff_model.py:229-236: [B,O,F] -> [B*S,O,F] with row = b*S + s."""
    return x.repeat(repeat_n, 1, 1, 1).transpose(0, 1).flatten(0, 1)


def evaluation_step(sd, net_param, batch, sched=None, draw=None, y_0_hat=None, gx=None,
                    variant_adds_eps=True):
    """NsDiff_model_variants.evaluation_step (NsDiff_model.py:404-495) and
    NsDiff_model.evaluation_step (:180-268) with f(x) / g(x) supplied by the caller.

    ``y_0_hat`` None -> zeros (variants without f(x), :446); ``gx`` None -> g(x) from
    sigma_oracle if the state dict has it, else ones (:458).  ``variant_adds_eps`` adds EPS
    to gx as the variants class does (:450) and the base class does not (:223).
    Returns outs [B,O,F,K] as a permuted view of contiguous [B,K,O,F], like the reference.
    """
    from . import sigma_oracle

    windows = net_param["windows"]
    pred_len = net_param["pred_len"]
    nf = net_param["dataset_nf"]
    T = net_param["diffusion_steps"]
    K = net_param["n_z_samples"]
    S = int(net_param["parallel_sample"])
    if sched is None:
        sched = nsdiff_schedule(net_param.get("diffusion_schedule", "linear"), T,
                                net_param.get("beta_start", 1e-4), net_param.get("beta_end", 0.02))
    if draw is None:
        draw = torch.randn_like
    batch_x = batch[:, :windows, :]
    b = batch_x.shape[0]
    if y_0_hat is None:
        y_0_hat = torch.zeros(b, pred_len, nf)
    if gx is None:
        if "cond_pred_model_g.mlp.0.weight" in sd:
            gx = sigma_oracle.sigma_estimation(sd, batch_x, net_param["rolling_length"], pred_len)
            if variant_adds_eps:
                gx = gx + EPS
        else:
            gx = torch.ones(b, pred_len, nf)
    preds = []
    for _ in range(K // S):
        y0_tile = tile_rows(y_0_hat, S)
        gx_tile = tile_rows(gx, S)
        seq = p_sample_loop(sd, sched, y0_tile, gx_tile, y0_tile, T, draw)
        preds.append(seq[T].reshape(b, S, pred_len, nf)[:, :, -pred_len:, :])
    preds = torch.concat(preds, dim=1)
    return preds.permute(0, 2, 3, 1)
