"""CPU oracle (test infrastructure, never shipped): f(x), the Non-stationary-Transformer condition encoder.

PARITY UNPINNED.  models/Diffusion_model/NsDiff/mu_backbone.py:53-183 (and TMDM/tmdm_ns_transformer.py:40-174)
assemble this model from blocks of torch-timeseries==0.1.10, a dependency that is not vendored in the
reference tree and cannot be installed here; no shipped checkpoint carries ``cond_pred_model.*`` weights and
the reference has no test for it.  The parts the reference spells out itself are restated from those files;
the library blocks are restated from the published Non-stationary Transformer design (Liu et al. 2022):
  DataEmbedding   = circular Conv1d(k=3, no bias) token embedding + sinusoidal positions (x_mark=None)
  DSAttention     = softmax(scale * (Q K^T * tau + delta)) V, causal mask when mask_flag
  EncoderLayer    = x = LN(x + attn(x)); x = LN(x + conv2(act(conv1(x))))          (1x1 convs)
  DecoderLayer    = self-attn(tau, no delta) -> cross-attn(tau, delta) -> FFN, three LNs
Weights: dict with the product's parameter names (``cond_pred_model.`` prefix stripped by the caller).
"""
import math

import torch
import torch.nn.functional as F


def _projector(sd, p, x, stats):
    """mu_backbone.py:30-41."""
    b = x.shape[0]
    w = sd[p + "series_conv.weight"]
    h = F.conv1d(F.pad(x, (1, 1), mode="circular"), w)               # [B,1,E]
    h = torch.cat([h, stats], dim=1).view(b, -1)
    i = 0
    while (p + "backbone.%d.weight" % i) in sd:
        bias = sd.get(p + "backbone.%d.bias" % i)
        h = F.linear(h, sd[p + "backbone.%d.weight" % i], bias)
        if (p + "backbone.%d.weight" % (i + 2)) in sd:
            h = F.relu(h)
        i += 2
    return h


def _embedding(sd, p, x):
    w = sd[p + "value_embedding.tokenConv.weight"]
    tok = F.conv1d(F.pad(x.permute(0, 2, 1), (1, 1), mode="circular"), w).transpose(1, 2)
    return tok + sd[p + "position_embedding.pe"][:, : x.shape[1]]


def _attention(sd, p, q_in, kv_in, n_heads, tau, delta, causal):
    B, Lq, _ = q_in.shape
    S = kv_in.shape[1]
    q = F.linear(q_in, sd[p + "query_projection.weight"], sd[p + "query_projection.bias"]).view(B, Lq, n_heads, -1)
    k = F.linear(kv_in, sd[p + "key_projection.weight"], sd[p + "key_projection.bias"]).view(B, S, n_heads, -1)
    v = F.linear(kv_in, sd[p + "value_projection.weight"], sd[p + "value_projection.bias"]).view(B, S, n_heads, -1)
    scale = 1.0 / math.sqrt(q.shape[-1])
    scores = torch.einsum("blhe,bshe->bhls", q, k)
    scores = scores * tau.view(B, 1, 1, 1)
    if delta is not None:
        scores = scores + delta.view(B, 1, 1, S)
    if causal:
        mask = torch.ones(Lq, S, dtype=torch.bool).triu(1)
        scores = scores.masked_fill(mask, float("-inf"))
    a = torch.softmax(scale * scores, dim=-1)
    out = torch.einsum("bhls,bshd->blhd", a, v).reshape(B, Lq, -1)
    return F.linear(out, sd[p + "out_projection.weight"], sd[p + "out_projection.bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"])


def _ffn(sd, p, x, act):
    y = F.conv1d(x.transpose(-1, 1), sd[p + "conv1.weight"], sd[p + "conv1.bias"])
    y = F.conv1d(act(y), sd[p + "conv2.weight"], sd[p + "conv2.bias"])
    return y.transpose(-1, 1)


def ns_transformer(sd, cfg, x_enc, vae=False):
    """-> dec_out [B, label_len+pred_len, F] de-normalised (callers slice the last pred_len)."""
    act = F.relu if cfg["activation"] == "relu" else F.gelu
    H = cfg["n_heads"]
    label_len, pred_len = cfg["label_len"], cfg["pred_len"]
    x_raw = x_enc
    mean_enc = x_enc.mean(1, keepdim=True)
    x = x_enc - mean_enc
    std_enc = torch.sqrt(torch.var(x, dim=1, keepdim=True, unbiased=False) + 1e-5)
    x = x / std_enc
    x_dec = torch.cat([x[:, -label_len:, :], torch.zeros(x.shape[0], pred_len, x.shape[2])], dim=1)
    tau = _projector(sd, "tau_learner.", x_raw, std_enc).exp()
    delta = _projector(sd, "delta_learner.", x_raw, mean_enc)
    h = _embedding(sd, "enc_embedding.", x)
    for l in range(cfg["e_layers"]):
        p = "encoder.attn_layers.%d." % l
        h = _ln(sd, p + "norm1.", h + _attention(sd, p + "attention.", h, h, H, tau, delta, False))
        h = _ln(sd, p + "norm2.", h + _ffn(sd, p, h, act))
    h = _ln(sd, "encoder.norm.", h)
    if vae:
        z = F.linear(F.relu(F.linear(h, sd["z_mean.0.weight"], sd["z_mean.0.bias"])), sd["z_mean.2.weight"], sd["z_mean.2.bias"])
        h = F.linear(F.relu(F.linear(z, sd["z_out.0.weight"], sd["z_out.0.bias"])), sd["z_out.2.weight"], sd["z_out.2.bias"])
    d = _embedding(sd, "dec_embedding.", x_dec)
    for l in range(cfg["d_layers"]):
        p = "decoder.layers.%d." % l
        d = _ln(sd, p + "norm1.", d + _attention(sd, p + "self_attention.", d, d, H, tau, None, True))
        d = _ln(sd, p + "norm2.", d + _attention(sd, p + "cross_attention.", d, h, H, tau, delta, False))
        d = _ln(sd, p + "norm3.", d + _ffn(sd, p, d, act))
    d = F.linear(_ln(sd, "decoder.norm.", d), sd["decoder.projection.weight"], sd["decoder.projection.bias"])
    return d * std_enc + mean_enc
