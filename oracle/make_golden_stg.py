"""Generate tests/golden/stg_*.npz by running the UNMODIFIED reference DiffSTG (build container only).

    python oracle/make_golden_stg.py

torch_geometric is absent: the reference's ``gnn.ResGatedGraphConv`` resolves to the stand-in
``oracle/_stubs/torch_geometric/nn/res_gated.py`` (published definition of the layer; "parity unpinned" for that layer),
everything else is the reference's own torch code.  Weights are ``diffusionts_oracle.synth_state_dict`` (no DiffSTG
checkpoint ships); fixtures store the seed and the key/shape list.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness, diffusionts_oracle as dto  # noqa: E402
from oracle.make_golden_dts import Tape, save  # noqa: E402

# ews_results/model_compare/DiffSTG/biomass/model_trained.yaml (BASELINE config 5), smaller sampling counts
YAML = dict(F=1, T_h=100, T_p=100, Td_h=10, channel_multipliers=[2, 2], d_h=4, diffusion_schedule="linear",
            diffusion_steps=200, gnn_name="ResGatedGraphConv", gnn_param={"bias": True, "root_weight": True},
            inference_diffusion_steps=20, inference_schedule="cosine", inference_trick="ddim",
            loss_weight_schedule="constant", mask_ratio=0.0, n_blocks=2, parallel_sampling=2, sequential_sampling=2,
            scaler_type="StandardScaler", task_model="DiffSTG")
SMALL = dict(YAML, T_h=12, T_p=8, Td_h=4, d_h=4, channel_multipliers=[2, 2], n_blocks=1, diffusion_steps=50,
             inference_diffusion_steps=6, inference_schedule="linear", parallel_sampling=3, sequential_sampling=2, F=1)


def ring_graph(n, extra):
    """Small directed edge list (both directions of an undirected graph), like from_networkx produces."""
    und = [(i, (i + 1) % n) for i in range(n)] + extra
    src = [a for a, b in und] + [b for a, b in und]
    dst = [b for a, b in und] + [a for a, b in und]
    order = sorted(range(len(src)), key=lambda e: (src[e], dst[e]))
    return torch.tensor([[src[e] for e in order], [dst[e] for e in order]], dtype=torch.long)


def build_reference(cfg, seed):
    from models.Diffusion_model.DiffSTG.graph_diffusion_model import DiffSTG
    m = DiffSTG(dict(cfg, device="cpu")).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith("model.")}
    sd = dto.synth_state_dict(shapes, seed)
    for k in shapes:                      # TcnBlock registers its conv twice (conv / net.0): one tensor
        if ".net.0." in k:
            sd[k] = sd[k.replace(".net.0.", ".conv.")]
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"scaler_mean", "scaler_std"}
    return m, shapes


def main():
    ref_harness.activate()
    from torch_geometric.data import Data
    torch.set_num_threads(1)
    for name, cfg, seed, V, extra in (("stg_small_evalstep.npz", SMALL, 31, 5, [(0, 2)]),
                                      ("stg_yaml_evalstep.npz", YAML, 37, 6, [(0, 3), (1, 4)])):
        m, shapes = build_reference(cfg, seed)
        ei = ring_graph(V, extra)
        torch.manual_seed(9)
        x = torch.randn(V, cfg["T_h"], 1).cumsum(1) * 0.1
        arrays = dict(cfg=json.dumps(cfg), seed=seed, keys=json.dumps({k: list(v) for k, v in shapes.items()}),
                      x=x, edge_index=ei)
        # one denoiser call at two steps
        T = cfg["T_h"] + cfg["T_p"]
        xm = torch.cat([x, torch.zeros(V, cfg["T_p"], 1)], dim=1)
        for t in (1, cfg["diffusion_steps"]):
            xt = torch.randn(V, T, 1)
            with torch.no_grad():
                eps = m.model(xt, torch.tensor([t]).int().float(), (xm, ei, None))
            arrays.update({"eps%d:xt" % t: xt, "eps%d:out" % t: eps})
        with Tape() as tape:
            outs, truth = m.evaluation_step(Data(x=x.clone(), edge_index=ei.clone(), num_nodes=V))
        assert truth is None
        arrays.update(outs=outs.contiguous(), n_draws=len(tape.draws))
        arrays.update({"z%03d" % i: z for i, z in enumerate(tape.draws)})
        save(name, **arrays)


if __name__ == "__main__":
    main()
