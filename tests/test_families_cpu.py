"""CPU checks for the DiffusionTS / DiffSTG families: the oracles against reference-made fixtures (bit-exact, one
thread), and the host-side logic of the product objects (key sets, integer schedules, graph bookkeeping).  No compute
call of the product runs here -- it has no CPU path."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import diffusionts_oracle as dto, diffstg_oracle as so

torch.set_num_threads(1)


def _load(name):
    g = np.load(os.path.join(GOLDEN, name))
    return g, json.loads(str(g["cfg"])), json.loads(str(g["keys"])), int(g["seed"])


def _stg_weights(shapes, seed):
    sd = dto.synth_state_dict(shapes, seed)
    for k in shapes:
        if ".net.0." in k:
            sd[k] = sd[k.replace(".net.0.", ".conv.")]
    return sd


# ------------------------------------------------------------------------------------------ DiffusionTS oracle
def test_dts_oracle_schedule_and_whole_evaluation_step_bit_exact():
    g, cfg, shapes, seed = _load("dts_small_evalstep.npz")
    tabs = dto.schedule_buffers(cfg["timesteps"], cfg["beta_schedule"])
    for k, v in tabs.items():
        assert torch.equal(v, torch.from_numpy(g["tab:" + k])), k
    sd = dto.synth_state_dict(shapes, seed)
    draws = iter([torch.from_numpy(g["z%03d" % i]) for i in range(int(g["n_draws"]))])

    def draw(shape):
        z = next(draws)
        assert tuple(z.shape) == tuple(shape)
        return z.clone()

    outs, by = dto.evaluation_step(sd, cfg, tabs, torch.from_numpy(g["batch"]), draw)
    assert by is None and torch.equal(outs.contiguous(), torch.from_numpy(g["outs"]))
    assert next(draws, None) is None              # consumed exactly the reference's number of draws


def test_dts_oracle_forward_gradient_and_loop_body_bit_exact():
    g, cfg, shapes, seed = _load("dts_yaml_steps.npz")
    sd = dto.synth_state_dict(shapes, seed)
    tabs = dto.schedule_buffers(cfg["timesteps"], cfg["beta_schedule"])
    for t in (0, 99):
        x = torch.from_numpy(g["fwd%d:x" % t])
        with torch.no_grad():
            tr, se = dto.transformer_forward(sd, cfg, x, torch.full((x.shape[0],), t, dtype=torch.long))
        assert torch.equal(tr, torch.from_numpy(g["fwd%d:trend" % t])) and torch.equal(se, torch.from_numpy(g["fwd%d:season" % t]))
    target = torch.from_numpy(g["target"])
    R, L = target.shape[0], cfg["windows"]
    mask = torch.cat([torch.ones(L, 1, dtype=torch.bool), torch.zeros(cfg["pred_len"], 1, dtype=torch.bool)]).expand(R, -1, -1)
    for time in (99, 3):
        key = "step%d:" % time
        draws = iter([torch.from_numpy(g[key + "z%d" % i]) for i in range(int(g[key + "n_draws"]))])
        with torch.no_grad():
            out = dto.infill_step(sd, cfg, tabs, torch.from_numpy(g[key + "img_in"]).clone(), time, time - 1, target, mask,
                                  0.1, 0.05, lambda s: next(draws).clone())
        assert torch.equal(out, torch.from_numpy(g[key + "img_out"])), time
    tc = torch.full((R,), 99, dtype=torch.long)
    pn, xs = dto.model_predictions(sd, cfg, tabs, torch.from_numpy(g["step99:img_in"]), tc)
    assert torch.equal(xs, torch.from_numpy(g["step99:x_start"])) and torch.equal(pn, torch.from_numpy(g["step99:pred_noise"]))
    an = tabs["alphas_cumprod"][98]
    gr = dto.langevin_grad(sd, cfg, torch.from_numpy(g["step99:ddim"]), tc, xs * an.sqrt() + (1 - an).sqrt() * pn,
                           torch.tensor(0.), target, mask, 0.1)
    assert torch.equal(gr, torch.from_numpy(g["step99:grad"]))


# ------------------------------------------------------------------------------------------ DiffSTG oracle
@pytest.mark.parametrize("name", ["stg_small_evalstep.npz", "stg_yaml_evalstep.npz"])
def test_stg_oracle_bit_exact(name):
    g, cfg, shapes, seed = _load(name)
    sd = _stg_weights(shapes, seed)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V = x.shape[0]
    xm = torch.cat([x, torch.zeros(V, cfg["T_p"], 1)], 1)
    for t in (1, cfg["diffusion_steps"]):
        with torch.no_grad():
            e = so.ugnet_forward(sd, cfg, torch.from_numpy(g["eps%d:xt" % t]), torch.tensor([t]).int().float(), xm, ei)
        assert torch.equal(e, torch.from_numpy(g["eps%d:out" % t]))
    draws = iter([torch.from_numpy(g["z%03d" % i]) for i in range(int(g["n_draws"]))])
    outs, truth = so.evaluation_step(sd, cfg, x, ei, V, lambda s: next(draws).clone())
    assert truth is None and torch.equal(outs.contiguous(), torch.from_numpy(g["outs"]))
    assert next(draws, None) is None


# ------------------------------------------------------------------------------------------ product host logic
def test_dts_product_keys_schedules_and_draw_count():
    from updgm_b200.diffusionts import DiffusionTS_model
    for name in ("dts_small_evalstep.npz", "dts_yaml_steps.npz"):
        g, cfg, shapes, seed = _load(name)
        m = DiffusionTS_model(dict(cfg, device="cpu"))
        own = m.state_dict()
        assert {k: list(v.shape) for k, v in own.items() if k.startswith("model.model.")} == shapes
        assert tuple(own["gt_mask"].shape) == (cfg["windows"] + cfg["pred_len"], cfg["dataset_nf"])
        tabs = dto.schedule_buffers(cfg["timesteps"], cfg["beta_schedule"])
        for k, v in tabs.items():
            assert torch.equal(own["model." + k], v), k
        assert m.time_pairs() == dto.sampling_times(cfg["timesteps"], cfg["diffusion_steps"])
        for t in range(cfg["timesteps"]):
            assert m.langevin_schedule(t, 0.05) == dto.langevin_k(t, cfg["timesteps"], 0.05)
        res = m.load_state_dict(dto.synth_state_dict(shapes, seed), strict=False)
        assert not res.unexpected_keys
    g, cfg, shapes, seed = _load("dts_small_evalstep.npz")
    m = DiffusionTS_model(dict(cfg, device="cpu"))
    assert m.draws_per_chunk() * (cfg["n_z_samples"] // cfg["parallel_sample"]) == int(g["n_draws"])
    g, cfg, shapes, seed = _load("dts_yaml_steps.npz")
    assert DiffusionTS_model(dict(cfg, device="cpu")).draws_per_chunk() == 327      # 1 + 99*2 + 128
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.evaluation_step(torch.zeros(2, cfg["windows"], cfg["dataset_nf"]))


def test_stg_product_keys_plan_and_graph_bookkeeping(tmp_path):
    from updgm_b200.diffstg import DiffSTG, graph_csr
    from updgm_b200 import uncertainty as U
    g, cfg, shapes, seed = _load("stg_yaml_evalstep.npz")
    m = DiffSTG(dict(cfg, device="cpu"))
    own = m.state_dict()
    assert {k: list(v.shape) for k, v in own.items() if k.startswith("model.")} == shapes
    assert set(own) - set(shapes) == {"scaler_mean", "scaler_std"}
    sd = _stg_weights(shapes, seed)
    sd.update(scaler_mean=torch.zeros(1), scaler_std=torch.ones(1))
    m.load_state_dict(sd, strict=True)
    diff = so.GaussianDiffusion(cfg["diffusion_steps"], cfg["diffusion_schedule"])
    plan = m.step_plan()
    assert len(plan) == cfg["inference_diffusion_steps"]
    for i, (t1, t2, a, b, c, noisy) in enumerate(plan):
        assert (t1, t2) == so.inference_schedule(cfg["inference_schedule"], diff.T, cfg["inference_diffusion_steps"], i)
        assert (a, b, c, noisy) == so.posterior_coefficients(diff, t1, t2, "ddim")
    assert plan[-1][5] and m.draws_per_round() * cfg["sequential_sampling"] == int(g["n_draws"])
    # CSR over targets keeps edge order; replicas share it
    ei = torch.tensor([[2, 0, 1, 0, 2], [1, 1, 0, 2, 2]])
    rowptr, col = graph_csr(ei, 3)
    assert rowptr.tolist() == [0, 1, 3, 5] and col.tolist() == [1, 2, 0, 0, 2]
    with pytest.raises(IndexError):
        graph_csr(torch.tensor([[0], [3]]), 3)
    # graphml -> edge_index in the order list(G.to_directed().edges) gives (what from_networkx emits)
    import networkx as nx
    G = nx.Graph()
    G.add_nodes_from(["n0", "n1", "n2", "n3"])
    G.add_edges_from([("n0", "n2"), ("n1", "n2"), ("n0", "n3")])
    path = tmp_path / "g.graphml"
    nx.write_graphml(G, path)
    gd = U.load_diffstg_graph(str(path))
    assert gd.num_nodes == 4 and gd.edge_index.tolist() == [[0, 0, 1, 2, 2, 3], [2, 3, 2, 0, 1, 0]]
    with pytest.raises(ValueError, match="graph_file is required"):
        U.load_diffstg_graph(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample_windows(torch.zeros(1, 6, cfg["T_h"], 1), torch.from_numpy(g["edge_index"]), 6)


def test_factory_builds_the_new_families():
    from updgm_b200 import loader
    g, cfg, shapes, seed = _load("dts_small_evalstep.npz")
    assert type(loader.diffusion_models("DiffusionTS", dict(cfg, device="cpu"))).__name__ == "DiffusionTS_model"
    g, cfg, shapes, seed = _load("stg_small_evalstep.npz")
    assert type(loader.diffusion_models("DiffSTG", dict(cfg, device="cpu"))).__name__ == "DiffSTG"


# ------------------------------------------------------------------------------------------ NsDiff_spatial
def _nsx_weights(g, shapes, seed):
    from oracle import nsdiff_spatial_oracle as nsx
    from updgm_b200.fx_encoder import PositionalEmbedding
    cfg = json.loads(str(g["cfg"]))
    sd = _stg_weights(shapes, seed)
    fshapes = json.loads(str(g["fx_keys"]))
    fx_sd = {nsx.FX + k: v for k, v in dto.synth_state_dict(fshapes, seed + 1).items()}
    pe = PositionalEmbedding(cfg["d_model"]).pe
    for k in fx_sd:
        if k.endswith("position_embedding.pe"):
            fx_sd[k] = pe.clone()
    sd.update(fx_sd)
    return sd


@pytest.mark.parametrize("name", ["nsx_small_evalstep.npz", "nsx_yaml_evalstep.npz"])
def test_nsx_oracle_matches_reference_run(name):
    from oracle import nsdiff_spatial_oracle as nsx, nsdiff_oracle as nso
    g, cfg, shapes, seed = _load(name)
    sd = _nsx_weights(g, shapes, seed)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V, S = x.shape[0], cfg["parallel_sample"]
    sched = nso.nsdiff_schedule(cfg["diffusion_schedule"], cfg["diffusion_steps"], cfg["beta_start"], cfg["beta_end"])
    for k in ("alphas", "betas_tilde", "betas_bar", "betas_tilde_m_1", "betas_bar_m_1", "one_minus_alphas_bar_sqrt",
              "alphas_cumprod_prev"):
        assert torch.equal(sched[k], torch.from_numpy(g["sched:" + k])), k
    par = so.duplicate_edge_index(S, ei, V)
    for t in (0, 1, cfg["diffusion_steps"] - 1):
        with torch.no_grad():
            eps, sig = nsx.ugnet_forward(sd, cfg, *[torch.from_numpy(g["den%d:%s" % (t, n)]) for n in ("y", "y0", "gx")],
                                         t, par)
        assert torch.equal(eps, torch.from_numpy(g["den%d:eps" % t])) and torch.equal(sig, torch.from_numpy(g["den%d:sig" % t]))
    draws = iter([torch.from_numpy(g["z%03d" % i]) for i in range(int(g["n_draws"]))])
    outs, by = nsx.evaluation_step(sd, cfg, x, ei, V, lambda like: next(draws).clone())
    assert by is None and torch.equal(outs.contiguous(), torch.from_numpy(g["outs"]))
    assert next(draws, None) is None and int(g["n_draws"]) == cfg["diffusion_steps"] * (cfg["n_z_samples"] // S)


def test_nsx_product_keys_and_loader(tmp_path):
    from updgm_b200.nsdiff_spatial import NsDiff_model_spatial
    from updgm_b200 import loader
    g, cfg, shapes, seed = _load("nsx_yaml_evalstep.npz")
    m = NsDiff_model_spatial(dict(cfg, device="cpu"), "NsDiff_model")
    own = m.state_dict()
    assert {k: list(v.shape) for k, v in own.items()
            if k.startswith("model.") or k.startswith("cond_pred_model_g.")} == shapes
    fshapes = json.loads(str(g["fx_keys"]))
    assert {k[len("cond_pred_model."):]: list(v.shape) for k, v in own.items() if k.startswith("cond_pred_model.")} == fshapes
    sd = _nsx_weights(g, shapes, seed)
    sd.update(scaler_mean=torch.zeros(1), scaler_std=torch.ones(1))
    m.load_state_dict(sd, strict=True)
    for k in ("alphas", "betas_tilde", "betas_bar_m_1", "one_minus_alphas_bar_sqrt"):
        assert torch.equal(getattr(m.model, k), torch.from_numpy(g["sched:" + k])), k
    ei = torch.from_numpy(g["edge_index"])
    assert torch.equal(m.duplicate_edge_index(3, ei, 6, "cpu"), so.duplicate_edge_index(3, ei, 6))
    assert loader.NOT_YET == {}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        from updgm_b200.diffstg import GraphData
        m.evaluation_step(GraphData(x=torch.from_numpy(g["x"]), edge_index=ei, num_nodes=6))
    with pytest.raises(ValueError, match="divisible"):
        NsDiff_model_spatial(dict(cfg, device="cpu", pred_len=101), "NsDiff_model")


def test_nsx_block_plan_and_shapes_match_oracle_and_reference_keys():
    """U-Net structure of the spatial denoiser: product plan == oracle plan (same blocks, channels, lengths), for the fixture
    architectures and a three-resolution one; odd lengths that the reference's UGnet would mis-size are rejected."""
    from oracle import nsdiff_spatial_oracle as nsx
    from updgm_b200.diffstg import block_plan
    from updgm_b200.nsdiff_spatial import ugnet_shapes
    for name in ("nsx_small_evalstep.npz", "nsx_yaml_evalstep.npz"):
        g, cfg, shapes, seed = _load(name)
        own = [sum(block_plan(cfg, T_total=cfg["pred_len"]), [])]
        ref = [[(p[len(nsx.UG):],) + tuple(rest) for (p, *rest) in sum(nsx.block_plan(cfg), [])]]
        assert own == ref
        sh = ugnet_shapes(cfg)
        assert {nsx.UG + k: list(v) for k, v in sh.items()} == {k: v for k, v in shapes.items() if k.startswith(nsx.UG)}
    cfg3 = dict(cfg, channel_multipliers=[2, 2, 2], pred_len=96, n_blocks=1)
    plan = sum(block_plan(cfg3, T_total=96), [])
    assert [b[4] for b in plan if b[1] == "res"][:3] == [96, 48, 24] and plan[-1][4] == 96
    assert sum(nsx.block_plan(cfg3), []) == [(nsx.UG + p,) + tuple(rest) for (p, *rest) in plan]
