"""CPU oracle (test infrastructure, never shipped): g(x) = SigmaEstimation, the "gx" path.

Follows models/Diffusion_model/NsDiff/g_backbone.py:19-72 and sigma.py:34-71.
Weights come from the reference state dict under ``cond_pred_model_g.mlp.{0,2,3,5,6}``.
"""
import torch
import torch.nn.functional as F

from .nsdiff_oracle import linear

G_PREFIX = "cond_pred_model_g."


def trailing_variance(x_enc, window_size):
    """sigma.py:34-71 (discard_rep=False): replicate-pad window_size-1 steps at the front,
    biased variance over every trailing window -> [B,T,N]."""
    if x_enc.dim() != 3:
        raise ValueError("x_enc must be a 3D tensor with shape (B, T, N)")
    T = x_enc.shape[1]
    if window_size < 1 or window_size > T:
        raise ValueError("window_size must be between 1 and T (got window_size={}, T={})".format(window_size, T))
    padded = F.pad(x_enc, (0, 0, window_size - 1, 0), mode="replicate")
    win = padded.unfold(1, window_size, 1).permute(0, 1, 3, 2)
    return win.var(dim=2, unbiased=False)


def sigma_estimation(sd, x_enc, rolling_length, pred_len, prefix=G_PREFIX):
    """g_backbone.py:49-72: variance series (last T-R steps, +1e-7) -> per-(b,f) MLP
    (T-R)->512->ReLU->LayerNorm([F,512])->512->ReLU->LayerNorm([F,512])->O -> softplus.
    LayerNorm normalises jointly over (F,512) of one batch row.  Returns [B,O,F]."""
    B, T, N = x_enc.shape
    sigma = trailing_variance(x_enc, rolling_length)
    sigma = sigma[:, -(T - rolling_length):, :] + 10e-8
    h = sigma.permute(0, 2, 1)
    hid = sd[prefix + "mlp.0.weight"].shape[0]
    h = F.relu(linear(h, sd[prefix + "mlp.0.weight"], sd[prefix + "mlp.0.bias"]))
    h = F.layer_norm(h, [N, hid], sd[prefix + "mlp.2.weight"], sd[prefix + "mlp.2.bias"])
    h = F.relu(linear(h, sd[prefix + "mlp.3.weight"], sd[prefix + "mlp.3.bias"]))
    h = F.layer_norm(h, [N, hid], sd[prefix + "mlp.5.weight"], sd[prefix + "mlp.5.bias"])
    h = linear(h, sd[prefix + "mlp.6.weight"], sd[prefix + "mlp.6.bias"])
    return F.softplus(h).permute(0, 2, 1)[:, -pred_len:, :]
