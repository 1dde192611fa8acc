"""CPU restatement of the DiffSTG graph-conv sampler.  TEST INFRASTRUCTURE ONLY.

Functional torch-CPU fp32 over a plain state dict with the reference's key names, same ops in the same order:

  UGnet / ResidualBlock / TcnBlock / SpatialBlock .. models/Diffusion_model/DiffSTG/ugnet.py:15-295
  GaussianDiffusion / InferenceSchedule ............. models/Diffusion_model/DiffSTG/diffusion_schedulers.py:39-125
  gaussian_posterior, evaluation_step ............... models/Diffusion_model/DiffSTG/graph_diffusion_model.py:46-73, 204-282
  ResGatedGraphConv ................................. un-vendored torch_geometric==2.5.3 (models/layer/gnn_conv.py:18-19)

PINNED (tests/golden/stg_*.npz, oracle/make_golden_stg.py) against the unmodified reference run with the
``oracle/_stubs/torch_geometric/nn/res_gated.py`` stand-in for the one third-party layer; that layer itself is
"parity unpinned" (its published definition is restated, nothing in the reference pins it, no DiffSTG checkpoint ships).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def block_plan(cfg):
    """[(key prefix, kind, c_in, c_out, T_in)] in execution order (ugnet.py:190-239)."""
    d_h, mults, n_blocks = cfg["d_h"], cfg["channel_multipliers"], cfg["n_blocks"]
    T_in = 2 * (cfg["T_p"] + cfg["T_h"])
    n_res = len(mults)
    down, up = [], []
    out_c = in_c = d_h
    idx = 0
    for i in range(n_res):
        out_c = in_c * mults[i]
        for _ in range(n_blocks):
            down.append(("model.down.%d.res." % idx, "res", in_c, out_c, T_in))
            idx += 1
            in_c = out_c
        if i < n_res - 1:
            down.append(("model.down.%d." % idx, "downsample", in_c, in_c, T_in))
            idx += 1
            T_in = math.floor((T_in - 1) / 2 + 1)
    middle = [("model.middle.res1.", "res", out_c, out_c, T_in), ("model.middle.res2.", "res", out_c, out_c, T_in)]
    in_c = out_c
    idx = 0
    for i in reversed(range(n_res)):
        out_c = in_c
        for _ in range(n_blocks):
            up.append(("model.up.%d.res." % idx, "res", in_c + out_c, out_c, T_in))
            idx += 1
        out_c = in_c // mults[i]
        up.append(("model.up.%d.res." % idx, "res", in_c + out_c, out_c, T_in))
        idx += 1
        in_c = out_c
        if i > 0:
            up.append(("model.up.%d." % idx, "upsample", in_c, in_c, T_in))
            idx += 1
            T_in = T_in * 2
    return down, middle, up


def time_embedding(t, dim):
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    e = t.float()[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, (0, 1, 0, 0))
    return e


def res_gated_graph_conv(sd, pre, x, edge_index):
    k = F.linear(x, sd[pre + "lin_key.weight"], sd[pre + "lin_key.bias"])
    q = F.linear(x, sd[pre + "lin_query.weight"], sd[pre + "lin_query.bias"])
    v = F.linear(x, sd[pre + "lin_value.weight"], sd[pre + "lin_value.bias"])
    src, dst = edge_index[0], edge_index[1]
    out = torch.zeros_like(k).index_add_(0, dst, torch.sigmoid(k[dst] + q[src]) * v[src])
    if pre + "lin_skip.weight" in sd:
        out = out + F.linear(x, sd[pre + "lin_skip.weight"])
    if pre + "bias" in sd:
        out = out + sd[pre + "bias"]
    return out


def _tcn(sd, pre, x, c_in, c_out):
    pad = 2
    out = F.conv2d(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"], padding=(1, pad))[:, :, :, :-pad]
    skip = x if c_in == c_out else F.conv2d(x, sd[pre + "shortcut.weight"], sd[pre + "shortcut.bias"])
    return out + skip


def residual_block(sd, pre, x, t_emb, edge_index, c_in, c_out, Td_h):
    h = _tcn(sd, pre + "tcn1.", x, c_in, c_out)
    h += F.conv2d(t_emb[:, :, None, None], sd[pre + "t_conv.weight"], sd[pre + "t_conv.bias"])
    h = _tcn(sd, pre + "tcn2.", h, c_out, c_out)
    h = F.layer_norm(h.transpose(1, 3), (1, c_out), sd[pre + "norm.weight"], sd[pre + "norm.bias"]).transpose(1, 3)
    h = F.conv2d(h, sd[pre + "downsampling.weight"], sd[pre + "downsampling.bias"], padding=(0, Td_h // 2))
    h = h.transpose(1, 3).squeeze(2)
    s = torch.relu(res_gated_graph_conv(sd, pre + "spatial.gnn.", h.reshape(h.shape[0], -1), edge_index))
    h = s.reshape(s.shape[0], Td_h, -1).unsqueeze(2).transpose(1, 3)
    h = F.conv_transpose2d(h, sd[pre + "upsampling.weight"], sd[pre + "upsampling.bias"], padding=(0, Td_h // 2))
    sc = x if c_in == c_out else F.conv2d(x, sd[pre + "shortcut.weight"], sd[pre + "shortcut.bias"])
    return h + sc


def ugnet_forward(sd, cfg, x, t, x_masked, edge_index):
    """UGnet.forward (ugnet.py:252-295): x, x_masked [N, T, F]; t [n_t] float -> eps prediction [N, T, F]."""
    down, middle, up = block_plan(cfg)
    Td_h = cfg["Td_h"]
    x = x.unsqueeze(2).transpose(1, 3)
    xm = x_masked.unsqueeze(2).transpose(1, 3)
    x = torch.cat((x, xm), dim=-1)
    x = F.conv2d(x, sd["model.x_proj.weight"], sd["model.x_proj.bias"])
    te = time_embedding(t, cfg["d_h"])
    hs = [x]

    def run(blk, x):
        pre, kind, c_in, c_out, _ = blk
        if kind == "res":
            return residual_block(sd, pre, x, te, edge_index, c_in, c_out, Td_h)
        if kind == "downsample":
            return F.conv2d(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"], stride=(1, 2), padding=(0, 1))
        return F.conv_transpose2d(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"], stride=(1, 2), padding=(0, 1))

    for blk in down:
        x = run(blk, x)
        hs.append(x)
    for blk in middle:
        x = run(blk, x)
    for blk in up:
        if blk[1] == "upsample":
            x = run(blk, x)
        else:
            x = torch.cat((x, hs.pop()), dim=1)
            x = run(blk, x)
    e = F.conv2d(x, sd["model.out.0.weight"], sd["model.out.0.bias"])
    e = F.linear(e, sd["model.out.1.weight"], sd["model.out.1.bias"])
    return e.squeeze(2).transpose(1, 2)


class GaussianDiffusion:
    """diffusion_schedulers.py:39-67 (float64 numpy tables)."""

    def __init__(self, T, schedule):
        self.T = T
        if schedule == "linear":
            self.beta = np.linspace(1e-4, 2e-2, T)
        elif schedule == "quad":
            self.beta = np.linspace(1e-4 ** 0.5, 2e-2 ** 5, T) ** 2
        elif schedule == "cosine":
            cos = lambda t: np.cos(math.pi * 0.5 * (t / T + 0.008) / (1 + 0.008)) ** 2
            ab = cos(np.arange(0, T + 1, 1)) / cos(0)
            self.beta = np.clip(1 - (ab[1:] / ab[:-1]), None, 0.999)
        self.alpha = np.concatenate((np.array([1.0]), 1 - self.beta))
        self.alphabar = np.cumprod(self.alpha)


def inference_schedule(kind, T, inference_T, i):
    """InferenceSchedule.__call__ (diffusion_schedulers.py:101-125) -> (t1, t2) as Python ints."""
    assert 0 <= i < inference_T
    if kind == "linear":
        t1 = T - int((float(i) / inference_T) * T)
        t2 = T - int((float(i + 1) / inference_T) * T)
    elif kind == "cosine":
        t1 = T - int(np.sin((float(i) / inference_T) * np.pi / 2) * T)
        t2 = T - int(np.sin((float(i + 1) / inference_T) * np.pi / 2) * T)
    else:
        raise ValueError("Unknown inference schedule: {}".format(kind))
    return int(np.clip(t1, 1, T)), int(np.clip(t2, 0, T - 1))


def posterior_coefficients(diff, t, target_t, trick="ddim"):
    """gaussian_posterior's scalars (graph_diffusion_model.py:46-73) -> (a, b, c, uses_noise):
    x_target = a * (xt - b * pred) + c * (z if uses_noise else pred)."""
    atbar, atbar_target = diff.alphabar[t], diff.alphabar[target_t]
    if trick == "ddpm" or t <= 1:
        at = diff.alpha[t]
        atbar_prev = diff.alphabar[t - 1]
        beta_tilde = diff.beta[t - 1] * (1 - atbar_prev) / (1 - atbar)
        return float(1 / np.sqrt(at)), float((1 - at) / np.sqrt(1 - atbar)), float(np.sqrt(beta_tilde)), True
    if trick == "ddim":
        return float(np.sqrt(atbar_target / atbar)), float(np.sqrt(1 - atbar)), float(np.sqrt(1 - atbar_target)), False
    raise ValueError("Unknown inference trick {}".format(trick))


def duplicate_edge_index(parallel, edge_index, num_nodes):
    ei = edge_index.reshape((2, 1, -1))
    ind = torch.arange(0, parallel).view(1, -1, 1) * num_nodes
    return (ei + ind).reshape((2, -1))


def evaluation_step(sd, cfg, x, edge_index, num_nodes, draw):
    """DiffSTG.evaluation_step (graph_diffusion_model.py:204-282) for a single graph: x [Node, T_h(+T_p), F] scaled."""
    T_h, T_p = cfg["T_h"], cfg["T_p"]
    T = T_h + T_p
    P_, Sq = cfg["parallel_sampling"], cfg["sequential_sampling"]
    diff = GaussianDiffusion(cfg["diffusion_steps"], cfg["diffusion_schedule"])
    history = x[:, :T_h, :]
    truth = None
    if x.shape[1] - T_h >= T_p:
        future = x[:, T_h:, :]
        assert future.size(1) == T_p, "pred_len is not equal to the length of the prediction"
        truth = torch.cat([history, future], dim=1)
    x_masked = torch.cat((history, torch.zeros(history.shape[0], T_p, history.shape[2])), dim=1)
    edge_index = edge_index.reshape(2, -1)
    if P_ > 1:
        edge_index = duplicate_edge_index(P_, edge_index, num_nodes)
        x_masked = x_masked.repeat(P_, 1, 1)
    outs = []
    steps = cfg["inference_diffusion_steps"]
    with torch.no_grad():
        for _ in range(Sq):
            xt = draw(x_masked.shape)
            for i in range(steps):
                t1, t2 = inference_schedule(cfg["inference_schedule"], diff.T, steps, i)
                pred = ugnet_forward(sd, cfg, xt.float(), torch.tensor([t1]).int().float(), x_masked, edge_index)
                a, b, c, noisy = posterior_coefficients(diff, t1, t2, cfg.get("inference_trick") or "ddim")
                z = draw(xt.shape) if noisy else None
                xt = a * (xt - b * pred)
                xt = xt + c * (z if noisy else pred)
            outs.append(xt.float())
    pl = torch.cat(outs, dim=0)
    return pl.reshape(Sq * P_, -1, T, 1).permute(1, 2, 3, 0), truth
