"""Generate tests/golden/nsx_*.npz by running the UNMODIFIED reference NsDiff_spatial classes (build container only).

    python oracle/make_golden_nsx.py

What runs as reference code: ``NsDiff_net_spatial`` (schedule tables + UGnet denoiser with eps / sigma heads),
``G.SigmaEstimation``, ``p_sample_loop`` and ``NsDiff_model_spatial.evaluation_step`` (tile order, edge duplication,
draw order).  What cannot: ``ns_Transformer.Model_spatial`` is assembled from torch-timeseries==0.1.10 blocks (absent, the
stubs hold no arithmetic), so for the evaluation_step fixture that one attribute is replaced by a module that evaluates
``nsdiff_spatial_oracle.model_spatial_forward`` -- f(x) of this class stays "parity unpinned" exactly like fx_oracle.
``gnn.ResGatedGraphConv`` resolves to ``oracle/_stubs/torch_geometric/nn/res_gated.py`` (published definition).
Weights are ``diffusionts_oracle.synth_state_dict``; fixtures store the seed and the key/shape list.
"""
import json
import os
import sys
from types import SimpleNamespace

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness, diffusionts_oracle as dto, nsdiff_spatial_oracle as nsx  # noqa: E402
from oracle.make_golden_dts import Tape, save  # noqa: E402
from oracle.make_golden_stg import ring_graph  # noqa: E402

BASE = dict(task_model="NsDiff_spatial", scaler_type="StandardScaler", load_pretrain=False, diffusion_schedule="linear",
            beta_start=1e-4, beta_end=0.02, gnn_name="ResGatedGraphConv", gnn_param={"bias": True, "root_weight": True},
            f_gnn_name="ResGatedGraphConv", f_gnn_param={"bias": True, "root_weight": True}, activation="gelu",
            factor=3, dropout=0.0, output_attention=False, p_hidden_layers=2)
SMALL = dict(BASE, dataset_nf=2, windows=16, pred_len=8, rolling_length=8, diffusion_steps=6, n_z_samples=6,
             parallel_sample=3, d_h=4, channel_multipliers=[2, 2], n_blocks=1, Td_h=4, d_model=64, n_heads=4, d_ff=64,
             e_layers=1, d_layers=1, fT_h=4, spatial_layers=1, p_hidden_dims=[16, 16])
YAMLISH = dict(BASE, dataset_nf=1, windows=100, pred_len=100, rolling_length=50, diffusion_steps=20, n_z_samples=4,
               parallel_sample=2, d_h=4, channel_multipliers=[2, 2], n_blocks=2, Td_h=10, d_model=64, n_heads=4, d_ff=128,
               e_layers=2, d_layers=1, fT_h=10, spatial_layers=2, p_hidden_dims=[64, 64])


def fx_shapes(cfg):
    """Key -> shape of Model_spatial's parameters under the reference's names (mu_backbone.py:191-262)."""
    from updgm_b200.nsdiff_spatial import NsTransformerSpatial
    m = NsTransformerSpatial(SimpleNamespace(**dict(cfg, seq_len=cfg["windows"], label_len=cfg["windows"] // 2)))
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def build_reference(cfg, seed, fx_sd):
    import models.Diffusion_model.NsDiff.NsDiff_model as ref

    class OracleFx(nn.Module):
        def __init__(self, configs):
            super().__init__()

        def forward(self, x, dec_inp, edge_index):
            d = nsx.model_spatial_forward(fx_sd, cfg, x, edge_index)
            return d[:, -cfg["pred_len"]:, :], d

    keep = ref.ns_Transformer.Model_spatial
    ref.ns_Transformer.Model_spatial = OracleFx
    try:
        m = ref.NsDiff_model_spatial(dict(cfg, device="cpu"), "NsDiff_model").eval()
    finally:
        ref.ns_Transformer.Model_spatial = keep
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()
              if k.startswith("model.") or k.startswith("cond_pred_model_g.")}
    sd = dto.synth_state_dict(shapes, seed)
    for k in shapes:                      # TcnBlock registers its conv twice (conv / net.0): one tensor
        if ".net.0." in k:
            sd[k] = sd[k.replace(".net.0.", ".conv.")]
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"scaler_mean", "scaler_std"}, res
    return m, shapes, sd


def main():
    ref_harness.activate()
    from torch_geometric.data import Data
    torch.set_num_threads(1)
    for name, cfg, seed, V, extra in (("nsx_small_evalstep.npz", SMALL, 41, 5, [(0, 2)]),
                                      ("nsx_yaml_evalstep.npz", YAMLISH, 43, 6, [(0, 3), (1, 4)])):
        fshapes = fx_shapes(cfg)
        fx_sd = {nsx.FX + k: v for k, v in dto.synth_state_dict(fshapes, seed + 1).items()}
        # positional table is a fixed buffer, not a weight
        from updgm_b200.fx_encoder import PositionalEmbedding
        pe = PositionalEmbedding(cfg["d_model"]).pe
        for k in fx_sd:
            if k.endswith("position_embedding.pe"):
                fx_sd[k] = pe.clone()
        m, shapes, sd = build_reference(cfg, seed, fx_sd)
        ei = ring_graph(V, extra)
        torch.manual_seed(11)
        x = 1.0 + torch.randn(V, cfg["windows"], cfg["dataset_nf"]).cumsum(1) * 0.1
        arrays = dict(cfg=json.dumps(cfg), seed=seed, keys=json.dumps({k: list(v) for k, v in shapes.items()}),
                      fx_keys=json.dumps({k: list(v) for k, v in fshapes.items()}), x=x, edge_index=ei)
        # schedule tables as the reference builds them
        for k in ("alphas", "betas_tilde", "betas_bar", "betas_tilde_m_1", "betas_bar_m_1", "one_minus_alphas_bar_sqrt",
                  "alphas_cumprod_prev"):
            arrays["sched:" + k] = getattr(m.model, k)
        # the denoiser alone at three steps, on S replicas in the duplicated-edge layout
        S, O, nf = cfg["parallel_sample"], cfg["pred_len"], cfg["dataset_nf"]
        par = m.duplicate_edge_index(S, ei, V, "cpu")
        for t in (0, 1, cfg["diffusion_steps"] - 1):
            y = torch.randn(V * S, O, nf)
            y0 = torch.randn(V * S, O, nf) * 0.5
            gx = torch.rand(V * S, O, nf) + 0.2
            with torch.no_grad():
                eps, sig = m.model(y, y0, gx, torch.tensor([t]), par)
            arrays.update({"den%d:y" % t: y, "den%d:y0" % t: y0, "den%d:gx" % t: gx, "den%d:eps" % t: eps,
                           "den%d:sig" % t: sig})
        with Tape() as tape, torch.no_grad():
            outs, batch_y = m.evaluation_step(Data(x=x.clone(), edge_index=ei.clone(), num_nodes=V))
            y0_hat = nsx.model_spatial_forward(fx_sd, cfg, x, ei)[:, -O:, :]
            gx = m.cond_pred_model_g(x)
        assert batch_y is None and torch.isfinite(outs).all(), "non-finite reference output: pick another seed"
        arrays.update(outs=outs.contiguous(), n_draws=len(tape.draws), y0_hat=y0_hat, gx=gx)
        arrays.update({"z%03d" % i: z for i, z in enumerate(tape.draws)})
        save(name, **arrays)


if __name__ == "__main__":
    main()
