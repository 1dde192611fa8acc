"""CPU restatement of the DiffusionTS conditional sampler.  TEST INFRASTRUCTURE ONLY.

Functional torch-CPU fp32 over a plain state dict with the reference's key names, in the reference's own
op order, so that (single-threaded) it reproduces the reference bit for bit:

  Transformer / blocks ........ models/Diffusion_model/DiffusionTS/diffusionts_transformer.py:123-438
  AdaLayerNorm, Conv_MLP, PE .. models/Diffusion_model/DiffusionTS/diffusionts_model_utils.py:66-85,151-161,187-202
  schedule buffers ............ models/Diffusion_model/DiffusionTS/DiffusionTS.py:73-125
  model_predictions ........... DiffusionTS.py:152-160
  fast_sample_infill .......... DiffusionTS.py:277-310
  langevin_fn ................. DiffusionTS.py:359-407
  evaluation_step ............. models/Diffusion_model/DiffusionTS/DiffusionTS_model.py:72-109

PINNED: tests/golden/dts_*.npz are produced by the unmodified reference (oracle/make_golden_dts.py) and
tests/test_oracle_golden.py holds this file to them.

Numerical note (measured on the reference itself): the Langevin refinement is a freshly-initialised Adagrad
step, i.e. ``x -= lr * g / (|g| + 1e-10)`` -- a sign step.  A sign flip of a near-zero gradient moves an element
by 2*lr, so a whole sampling loop is chaotic: the reference run with 1 and with 8 CPU threads on the same noise
differs by O(1) per value.  Parity for this family is therefore pinned per step (same input, same noise), and for
whole loops only on one thread / statistically.
"""
import math

import torch
import torch.nn.functional as F

P = "model.model."


def synth_state_dict(shapes, seed):
    """Deterministic weights for fixtures: numpy RandomState (frozen stream) keyed by parameter name, scaled like a
    default init.  ``shapes``: {key: shape} of the float parameters.  Schedules/buffers are not touched."""
    import zlib
    import numpy as np
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 32))
        u = rs.uniform(-1.0, 1.0, size=shp).astype(np.float32)
        if k.endswith(".pe"):
            w = 0.02 * u
        elif ".ln2." in k or ".norm." in k:
            w = (1.0 + 0.1 * u) if k.endswith("weight") else 0.1 * u
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if len(shp) == 1:
                fan_in = shp[0]
            w = u / math.sqrt(max(fan_in, 1))
            if k.endswith("combine_s.weight"):
                w = w * 0.01     # the Fourier head resynthesises un-normalised rfft amplitudes (x seq_len): keep x0 unsaturated
            if ".trend.trend.3." in k:
                w = w * 0.1      # six polynomial trend heads add up
        sd[k] = torch.from_numpy(np.ascontiguousarray(w))
    return sd


def schedule_buffers(timesteps, beta_schedule="cosine"):
    """DiffusionTS.py:18-35,73-125 (float64 tables cast to float32)."""
    if beta_schedule == "linear":
        scale = 1000 / timesteps
        betas = torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
    elif beta_schedule == "cosine":
        s = 0.008
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
        ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    else:
        raise ValueError(f"unknown beta schedule {beta_schedule}")
    alphas = 1. - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.)
    pv = betas * (1. - ac_prev) / (1. - ac)
    t = {
        "betas": betas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac), "sqrt_one_minus_alphas_cumprod": torch.sqrt(1. - ac),
        "log_one_minus_alphas_cumprod": torch.log(1. - ac), "sqrt_recip_alphas_cumprod": torch.sqrt(1. / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1. / ac - 1), "posterior_variance": pv,
        "posterior_log_variance_clipped": torch.log(pv.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1. - ac),
        "posterior_mean_coef2": (1. - ac_prev) * torch.sqrt(alphas) / (1. - ac),
        "loss_weight": torch.sqrt(alphas) * torch.sqrt(1. - ac) / betas / 100,
    }
    return {k: v.to(torch.float32) for k, v in t.items()}


def _sinusoidal(t, dim):
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def _ada_ln(sd, pre, x, t):
    d = x.shape[-1]
    emb = F.linear(F.silu(_sinusoidal(t, d)), sd[pre + "linear.weight"], sd[pre + "linear.bias"]).unsqueeze(1)
    scale, shift = torch.chunk(emb, 2, dim=2)
    return F.layer_norm(x, (d,)) * (1 + scale) + shift


def _attention(sd, pre, x, ctx, n_head):
    B, T, C = x.shape
    TE = ctx.shape[1]
    hs = C // n_head
    k = F.linear(ctx, sd[pre + "key.weight"], sd[pre + "key.bias"]).view(B, TE, n_head, hs).transpose(1, 2)
    q = F.linear(x, sd[pre + "query.weight"], sd[pre + "query.bias"]).view(B, T, n_head, hs).transpose(1, 2)
    v = F.linear(ctx, sd[pre + "value.weight"], sd[pre + "value.bias"]).view(B, TE, n_head, hs).transpose(1, 2)
    att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(k.size(-1)))
    att = F.softmax(att, dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(y, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def _mlp(sd, pre, x):
    h = F.gelu(F.linear(x, sd[pre + "0.weight"], sd[pre + "0.bias"]))
    return F.linear(h, sd[pre + "2.weight"], sd[pre + "2.bias"])


def _conv_mlp(sd, pre, x):
    return F.conv1d(x.transpose(1, 2), sd[pre + "sequential.1.weight"], sd[pre + "sequential.1.bias"],
                    padding=1).transpose(1, 2)


def _trend_block(sd, pre, x, out_dim):
    """diffusionts_transformer.py:12-35: conv over the embedding axis with the positions as channels."""
    h = F.gelu(F.conv1d(x, sd[pre + "trend.trend.0.weight"], sd[pre + "trend.trend.0.bias"], padding=1))
    h = F.conv1d(h.transpose(1, 2), sd[pre + "trend.trend.3.weight"], sd[pre + "trend.trend.3.bias"], padding=1).transpose(1, 2)
    lin = torch.arange(1, out_dim + 1, 1) / (out_dim + 1)
    poly = torch.stack([lin ** float(p + 1) for p in range(3)], dim=0)
    return torch.matmul(h.transpose(1, 2), poly).transpose(1, 2)


def _fourier_layer(x, low_freq=1, factor=1):
    """diffusionts_transformer.py:52-103: keep the top-k rfft bins per (row, channel), resynthesise."""
    b, t, d = x.shape
    xf = torch.fft.rfft(x, dim=1)
    if t % 2 == 0:
        xf = xf[:, low_freq:-1]
        f = torch.fft.rfftfreq(t)[low_freq:-1]
    else:
        xf = xf[:, low_freq:]
        f = torch.fft.rfftfreq(t)[low_freq:]
    top_k = int(factor * math.log(xf.shape[1]))
    _, idx = torch.topk(xf.abs(), top_k, dim=1, largest=True, sorted=True)
    ma, mb = torch.meshgrid(torch.arange(b), torch.arange(d), indexing="ij")
    tup = (ma.unsqueeze(1), idx, mb.unsqueeze(1))
    xf = xf[tup]
    f = f[None, :, None].expand(b, -1, d)[tup].unsqueeze(2)
    xf = torch.cat([xf, xf.conj()], dim=1)
    f = torch.cat([f, -f], dim=1)
    tt = torch.arange(t, dtype=torch.float)[None, None, :, None]
    amp = xf.abs().unsqueeze(2)
    phase = xf.angle().unsqueeze(2)
    return (amp * torch.cos(2 * math.pi * f * tt + phase)).sum(dim=1)


def transformer_forward(sd, cfg, x, t):
    """Transformer.forward (diffusionts_transformer.py:420-437) -> (trend, season_error)."""
    nh, seq = cfg["n_heads"], x.shape[1]
    emb = _conv_mlp(sd, P + "emb.", x)
    h = emb + sd[P + "pos_enc.pe"]
    for i in range(cfg["n_layer_enc"]):
        pre = P + "encoder.blocks.%d." % i
        a = _ada_ln(sd, pre + "ln1.", h, t)
        h = h + _attention(sd, pre + "attn.", a, a, nh)
        h = h + _mlp(sd, pre + "mlp.", F.layer_norm(h, (h.shape[-1],), sd[pre + "ln2.weight"], sd[pre + "ln2.bias"]))
    enc = h
    h = emb + sd[P + "pos_dec.pe"]
    b, c, d = h.shape
    nf = cfg["dataset_nf"]
    season = torch.zeros((b, c, d))
    trend = torch.zeros((b, c, nf))
    means = []
    for i in range(cfg["n_layer_dec"]):
        pre = P + "decoder.blocks.%d." % i
        a = _ada_ln(sd, pre + "ln1.", h, t)
        h = h + _attention(sd, pre + "attn1.", a, a, nh)
        h = h + _attention(sd, pre + "attn2.", _ada_ln(sd, pre + "ln1_1.", h, t), enc, nh)
        x1, x2 = F.conv1d(h, sd[pre + "proj.weight"], sd[pre + "proj.bias"]).chunk(2, dim=1)
        tr, se = _trend_block(sd, pre, x1, seq), _fourier_layer(x2)
        h = h + _mlp(sd, pre + "mlp.", F.layer_norm(h, (d,), sd[pre + "ln2.weight"], sd[pre + "ln2.bias"]))
        m = torch.mean(h, dim=1, keepdim=True)
        h = h - m
        means.append(F.linear(m, sd[pre + "linear.weight"], sd[pre + "linear.bias"]))
        season += se
        trend += tr
    mean = torch.cat(means, dim=1)
    res = _conv_mlp(sd, P + "inverse.", h)
    res_m = torch.mean(res, dim=1, keepdim=True)
    ws = sd[P + "combine_s.weight"]
    pad = (ws.shape[-1] - 1) // 2
    s_in = season.transpose(1, 2)
    if pad:
        s_in = F.pad(s_in, (pad, pad), mode="circular")
    season_error = F.conv1d(s_in, ws).transpose(1, 2) + res - res_m
    trend = F.conv1d(mean, sd[P + "combine_m.weight"]) + res_m + trend
    return trend, season_error


def output(sd, cfg, x, t):
    trend, season = transformer_forward(sd, cfg, x, t)
    return trend + season


def _ext(tab, t, x):
    return tab.gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))


def model_predictions(sd, cfg, tabs, x, t, clip_x_start=True):
    x_start = output(sd, cfg, x, t)
    if clip_x_start:
        x_start = torch.clamp(x_start, min=-1., max=1.)
    pred_noise = (_ext(tabs["sqrt_recip_alphas_cumprod"], t, x) * x - x_start) / \
        _ext(tabs["sqrt_recipm1_alphas_cumprod"], t, x)
    return pred_noise, x_start


def langevin_k(t0, num_timesteps, learning_rate):
    """Number of refinement iterations and the learning rate at step t0 (DiffusionTS.py:372-381)."""
    if t0 < num_timesteps * 0.05:
        return 0, learning_rate
    if t0 > num_timesteps * 0.9:
        return 3, learning_rate
    if t0 > num_timesteps * 0.75:
        return 2, learning_rate * 0.5
    return 1, learning_rate * 0.25


def langevin_grad(sd, cfg, x, t, mean, sigma, target, partial_mask, coef):
    """Gradient of the refinement loss at x (DiffusionTS.py:387-399)."""
    p = x.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        x_start = output(sd, cfg, p, t)
        if float(sigma) == 0:
            logp = coef * ((mean - p) ** 2 / 1.).mean(dim=0).sum()
            infill = ((x_start[partial_mask] - target[partial_mask]) ** 2).mean(dim=0).sum()
        else:
            logp = coef * ((mean - p) ** 2 / sigma).mean(dim=0).sum()
            infill = (((x_start[partial_mask] - target[partial_mask]) ** 2) / sigma.mean()).mean(dim=0).sum()
        (logp + infill).backward()
    return p.grad


def langevin_fn(sd, cfg, num_timesteps, coef, partial_mask, target, learning_rate, sample, mean, sigma, t, draw,
                coef_=0.):
    K, lr = langevin_k(t[0].item(), num_timesteps, learning_rate)
    p = sample
    for _ in range(K):
        g = langevin_grad(sd, cfg, p, t, mean, sigma, target, partial_mask, coef)
        # torch.optim.Adagrad from a zero state, lr_decay 0, eps 1e-10: sum = g*g; p -= lr * g / (sqrt(sum) + eps)
        p = p.detach().addcdiv(g, (g * g).sqrt().add_(1e-10), value=-lr)
        eps = draw(p.shape)
        p = (p + coef_ * sigma.mean().item() * eps).detach()
    sample[~partial_mask] = p[~partial_mask]
    return sample


def sampling_times(num_timesteps, sampling_timesteps):
    times = torch.linspace(-1, num_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def infill_step(sd, cfg, tabs, img, time, time_next, target, partial_mask, coef, learning_rate, draw, eta=0.):
    """One iteration of the loop body of fast_sample_infill (DiffusionTS.py:287-306), time_next >= 0."""
    batch = img.shape[0]
    T = tabs["betas"].shape[0]
    tc = torch.full((batch,), time, dtype=torch.long)
    pred_noise, x_start = model_predictions(sd, cfg, tabs, img, tc, clip_x_start=True)
    alpha = tabs["alphas_cumprod"][time]
    alpha_next = tabs["alphas_cumprod"][time_next]
    sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
    c = (1 - alpha_next - sigma ** 2).sqrt()
    pred_mean = x_start * alpha_next.sqrt() + c * pred_noise
    noise = draw(img.shape)
    img = pred_mean + sigma * noise
    img = langevin_fn(sd, cfg, T, coef, partial_mask, target, learning_rate, img, pred_mean, sigma, tc, draw)
    target_t = _ext(tabs["sqrt_alphas_cumprod"], tc, target) * target + \
        _ext(tabs["sqrt_one_minus_alphas_cumprod"], tc, target) * draw(target.shape)
    img[partial_mask] = target_t[partial_mask]
    return img


def fast_sample_infill(sd, cfg, tabs, shape, target, partial_mask, sampling_timesteps, coef, learning_rate, draw,
                       eta=0.):
    with torch.no_grad():
        T = tabs["betas"].shape[0]
        img = draw(tuple(shape))
        for time, time_next in sampling_times(T, sampling_timesteps):
            if time_next < 0:
                tc = torch.full((shape[0],), time, dtype=torch.long)
                _, img = model_predictions(sd, cfg, tabs, img, tc, clip_x_start=True)
                continue
            img = infill_step(sd, cfg, tabs, img, time, time_next, target, partial_mask, coef, learning_rate, draw,
                              eta)
        img[partial_mask] = target[partial_mask]
        return img


def evaluation_step(sd, cfg, tabs, batch, draw):
    """DiffusionTS_model.evaluation_step (:72-109).  Note the reference's row bookkeeping: ``x.repeat(S,1,1)`` orders
    the chunk rows (sample, node) but the result is reshaped as (node, sample) -- reproduced here as written."""
    L, O, nf = cfg["windows"], cfg["pred_len"], cfg["dataset_nf"]
    K = cfg["n_z_samples"]
    S = min(cfg["parallel_sample"], K)
    if K % S != 0:
        raise ValueError("n_z_samples must be divisible by parallel_sample")
    batch_x = batch[:, :L, :]
    batch_y = batch[:, L:L + O, :] if batch.shape[1] - L >= O else None
    x = torch.cat([batch_x, torch.zeros(batch_x.shape[0], O, nf)], dim=1)
    gt_mask = torch.cat([torch.ones(L, nf, dtype=torch.bool), torch.zeros(O, nf, dtype=torch.bool)], dim=0)
    mask = gt_mask.expand(x.shape[0], -1, -1)
    samples = []
    for _ in range(K // S):
        rx, rm = x.repeat(S, 1, 1), mask.repeat(S, 1, 1)
        s = fast_sample_infill(sd, cfg, tabs, rx.shape, rx * rm, rm, cfg["diffusion_steps"],
                               cfg.get("infill_coef", 1e-1), cfg.get("infill_learning_rate", 5e-2), draw,
                               cfg.get("eta", 0.0))
        samples.append(s[:, -O:, :].reshape(x.shape[0], S, O, nf))
    preds = torch.cat(samples, dim=1)
    return preds.reshape(x.shape[0], K, O, nf).permute(0, 2, 3, 1), batch_y
