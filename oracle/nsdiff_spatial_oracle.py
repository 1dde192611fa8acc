"""CPU restatement of NsDiff_spatial (third class).  TEST INFRASTRUCTURE ONLY.

Functional torch-CPU fp32 over a plain state dict with the reference's key names, same ops in the same order:

  NsDiff_model_spatial.evaluation_step / duplicate_edge_index .. models/Diffusion_model/NsDiff/NsDiff_model.py:695-801
  NsDiff_net_spatial (schedule tables + UGnet denoiser) ......... models/Diffusion_model/NsDiff/NsDiff_net.py:175-264
  UGnet with eps / sigma heads ................................. models/Diffusion_model/NsDiff/ugnet.py:194-293
  Model_spatial (f(x) with a graph block between enc and dec) ... models/Diffusion_model/NsDiff/mu_backbone.py:186-345
  p_sample / p_sample_t_1to0 / p_sample_loop .................... models/Diffusion_model/NsDiff/nsdiff_utils.py:111-284

PINNED (tests/golden/nsx_*.npz, oracle/make_golden_nsx.py): the UGnet denoiser and the whole evaluation_step (tile order,
edge duplication, T draws per chunk, posterior algebra) against the unmodified reference classes, with the
``oracle/_stubs/torch_geometric/nn/res_gated.py`` stand-in for ResGatedGraphConv ("parity unpinned" for that layer).
``model_spatial_forward`` is PARITY UNPINNED like ``fx_oracle`` (its transformer blocks come from the un-vendored
torch-timeseries==0.1.10); the parts the reference spells out itself (normalisation, the (1, T+1) down / up convolutions,
the graph blocks' placement, de-normalisation) are restated from mu_backbone.py:301-345.

Reference quirk kept on purpose: rows of a chunk are ordered ``b*S + s`` (NsDiff_model.py:749-757) while the duplicated
edge list addresses rows as ``s*V + b`` (:792-801), so the graph conv couples rows the way the second layout says.
"""
import torch
import torch.nn.functional as F

from . import diffstg_oracle as stg
from . import fx_oracle as fx
from . import nsdiff_oracle as nso
from . import sigma_oracle

UG = "model.diffussion_model."
FX = "cond_pred_model."


def block_plan(cfg, prefix=UG):
    """[(key prefix, kind, c_in, c_out, T_in)] in execution order (NsDiff/ugnet.py:205-245; T = pred_len, T_in//2)."""
    d_h, mults, n_blocks = cfg["d_h"], cfg["channel_multipliers"], cfg["n_blocks"]
    T_in = cfg["pred_len"]
    n_res = len(mults)
    down, up = [], []
    out_c = in_c = d_h
    idx = 0
    for i in range(n_res):
        out_c = in_c * mults[i]
        for _ in range(n_blocks):
            down.append((prefix + "down.%d.res." % idx, "res", in_c, out_c, T_in))
            idx += 1
            in_c = out_c
        if i < n_res - 1:
            down.append((prefix + "down.%d." % idx, "downsample", in_c, in_c, T_in))
            idx += 1
            T_in = T_in // 2
    middle = [(prefix + "middle.res1.", "res", out_c, out_c, T_in), (prefix + "middle.res2.", "res", out_c, out_c, T_in)]
    in_c = out_c
    idx = 0
    for i in reversed(range(n_res)):
        out_c = in_c
        for _ in range(n_blocks):
            up.append((prefix + "up.%d.res." % idx, "res", in_c + out_c, out_c, T_in))
            idx += 1
        out_c = in_c // mults[i]
        up.append((prefix + "up.%d.res." % idx, "res", in_c + out_c, out_c, T_in))
        idx += 1
        in_c = out_c
        if i > 0:
            up.append((prefix + "up.%d." % idx, "upsample", in_c, in_c, T_in))
            idx += 1
            T_in = T_in * 2
    assert T_in == cfg["pred_len"], "T_in should be equal to T"
    return down, middle, up


def ugnet_forward(sd, cfg, y_t, y_0_hat, gx, t, edge_index, prefix=UG):
    """UGnet.forward (NsDiff/ugnet.py:257-293): rows [N, T_p, F]; t: int step -> (eps_pred, sigma), both [N, T_p, F]."""
    down, middle, up = block_plan(cfg, prefix)
    Td_h = cfg["Td_h"]
    x = torch.cat((y_t, y_0_hat, gx), dim=-1).unsqueeze(2).transpose(1, 3)
    x = F.conv2d(x, sd[prefix + "x_proj.weight"], sd[prefix + "x_proj.bias"])
    te = stg.time_embedding(torch.tensor([t]), cfg["d_h"])
    hs = [x]

    def run(blk, x):
        pre, kind, c_in, c_out, _ = blk
        if kind == "res":
            return stg.residual_block(sd, pre, x, te, edge_index, c_in, c_out, Td_h)
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
            x = run(blk, torch.cat((x, hs.pop()), dim=1))
    e = F.conv2d(x, sd[prefix + "out.0.weight"], sd[prefix + "out.0.bias"])
    e = F.linear(e, sd[prefix + "out.1.weight"], sd[prefix + "out.1.bias"])
    e = e.squeeze(2).transpose(1, 2)
    eps = nso.linear(e, sd[prefix + "lin4.weight"], sd[prefix + "lin4.bias"])
    sig = F.softplus(nso.linear(F.softplus(e), sd[prefix + "sigma_lin.weight"], sd[prefix + "sigma_lin.bias"]))
    return eps, sig


def model_spatial_forward(sd, cfg, x_enc, edge_index, prefix=FX):
    """Model_spatial.forward (mu_backbone.py:301-345) -> dec_out [B, label_len+pred_len, F] de-normalised."""
    w = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    act = F.relu if cfg["activation"] == "relu" else F.gelu
    H, fT = cfg["n_heads"], cfg["fT_h"]
    label_len, pred_len = cfg["windows"] // 2, cfg["pred_len"]
    x_raw = x_enc
    mean_enc = x_enc.mean(1, keepdim=True)
    x = x_enc - mean_enc
    std_enc = torch.sqrt(torch.var(x, dim=1, keepdim=True, unbiased=False) + 1e-5)
    x = x / std_enc
    x_dec = torch.cat([x[:, -label_len:, :], torch.zeros(x.shape[0], pred_len, x.shape[2])], dim=1)
    tau = fx._projector(w, "tau_learner.", x_raw, std_enc).exp()
    delta = fx._projector(w, "delta_learner.", x_raw, mean_enc)
    h = fx._embedding(w, "enc_embedding.", x)
    for l in range(cfg["e_layers"]):
        p = "encoder.attn_layers.%d." % l
        h = fx._ln(w, p + "norm1.", h + fx._attention(w, p + "attention.", h, h, H, tau, delta, False))
        h = fx._ln(w, p + "norm2.", h + fx._ffn(w, p, h, act))
    h = fx._ln(w, "encoder.norm.", h)
    # (1, T+1) convolution down to fT_h steps, graph blocks on [rows, fT_h*d_model], transposed convolution back
    h = h.unsqueeze(2).transpose(1, 3)
    h = F.conv2d(h, w["downsampling.weight"], w["downsampling.bias"], padding=(0, fT // 2)).transpose(1, 3).squeeze(2)
    s = h.reshape(h.shape[0], -1)
    for l in range(cfg["spatial_layers"]):
        s = torch.relu(stg.res_gated_graph_conv(w, "spatial_encoder.%d.gnn." % l, s, edge_index))
    h = s.reshape(h.shape[0], fT, -1).unsqueeze(2).transpose(1, 3)
    h = F.conv_transpose2d(h, w["upsampling.weight"], w["upsampling.bias"], padding=(0, fT // 2)).transpose(1, 3).squeeze(2)
    d = fx._embedding(w, "dec_embedding.", x_dec)
    for l in range(cfg["d_layers"]):
        p = "decoder.layers.%d." % l
        d = fx._ln(w, p + "norm1.", d + fx._attention(w, p + "self_attention.", d, d, H, tau, None, True))
        d = fx._ln(w, p + "norm2.", d + fx._attention(w, p + "cross_attention.", d, h, H, tau, delta, False))
        d = fx._ln(w, p + "norm3.", d + fx._ffn(w, p, d, act))
    d = F.linear(fx._ln(w, "decoder.norm.", d), w["decoder.projection.weight"], w["decoder.projection.bias"])
    return d * std_enc + mean_enc


def p_sample(sd, cfg, sched, y, y_0_hat, gx, y_T_mean, t, z, edge_index):
    """nsdiff_utils.py:111-158 with the graph denoiser."""
    eps_theta, sigma_theta = ugnet_forward(sd, cfg, y, y_0_hat, gx, t, edge_index)
    sigma_y0_hat, noise = nso._sigma_y0_and_noise(sched, t, gx, sigma_theta)
    y_0 = nso._y0_reparam(sched, t, y, y_T_mean, eps_theta, noise)
    g0, g1, g2 = nso._gammas(sched, t, gx, sigma_y0_hat)
    return g0 * y_0 + g1 * y + g2 * y_T_mean + torch.sqrt(sigma_theta) * z


def p_sample_loop(sd, cfg, sched, y_0_hat, gx, y_T_mean, n_steps, draw, edge_index):
    """nsdiff_utils.py:271-284: one draw for y_T, one per t = T-1..1, none at t = 0 -> final y_0 estimate."""
    cur = torch.sqrt(gx) * draw(y_T_mean) + y_T_mean
    for t in reversed(range(1, n_steps)):
        cur = p_sample(sd, cfg, sched, cur, y_0_hat, gx, y_T_mean, t, draw(cur), edge_index)
    eps_theta, sigma_theta = ugnet_forward(sd, cfg, cur, y_0_hat, gx, 0, edge_index)
    _, noise = nso._sigma_y0_and_noise(sched, 0, gx, sigma_theta)
    return nso._y0_reparam(sched, 0, cur, y_T_mean, eps_theta, noise)


def evaluation_step(sd, cfg, x, edge_index, num_nodes, draw=None, y_0_hat=None, sched=None):
    """NsDiff_model_spatial.evaluation_step (NsDiff_model.py:695-790): x [Node, L(+O), F] scaled ->
    (outs [Node, O, F, K], batch_y or None).  ``y_0_hat`` overrides f(x) (used where f(x) cannot be pinned)."""
    L, O, nf = cfg["windows"], cfg["pred_len"], cfg["dataset_nf"]
    T, K, S = cfg["diffusion_steps"], cfg["n_z_samples"], int(cfg["parallel_sample"])
    if sched is None:
        sched = nso.nsdiff_schedule(cfg.get("diffusion_schedule", "linear"), T, cfg.get("beta_start", 1e-4),
                                    cfg.get("beta_end", 0.02))
    if draw is None:
        draw = torch.randn_like
    edge_index = edge_index.reshape(2, -1)
    batch_x = x[:, :L, :]
    batch_y = None
    if x.shape[1] - L >= O:
        batch_y = x[:, L:, :]
        assert batch_y.size(1) == O, "pred_len is not equal to the length of the prediction"
    b = batch_x.shape[0]
    with torch.no_grad():
        if y_0_hat is None:
            y_0_hat = model_spatial_forward(sd, cfg, batch_x, edge_index)[:, -O:, :]
        gx = sigma_oracle.sigma_estimation(sd, batch_x, cfg["rolling_length"], O)
        par_edges = stg.duplicate_edge_index(S, edge_index, num_nodes) if S > 1 else edge_index
        preds = []
        for _ in range(K // S):
            y0_tile = nso.tile_rows(y_0_hat, S)
            gx_tile = nso.tile_rows(gx, S)
            y = p_sample_loop(sd, cfg, sched, y0_tile, gx_tile, y0_tile, T, draw, par_edges)
            preds.append(y.reshape(b, S, O, nf))
    return torch.concat(preds, dim=1).permute(0, 2, 3, 1), batch_y
