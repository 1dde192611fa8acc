"""GPU tests of the drop-in surface for DiffusionTS / DiffSTG: a reference-format checkpoint directory
(`model_trained` = torch.save({'net_param','state_dict'}) + `model_trained.yaml`), [Node,Time,F] input, graphml
topology, `.pt` prediction cache, through `uncertainty_ews` -- the call the reference's figure scripts make."""
import json
import os

import numpy as np
import pytest
import torch
import yaml

from conftest import GOLDEN
from oracle import diffusionts_oracle as dto, mpv_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _U():
    from updgm_b200 import uncertainty
    return uncertainty


def _load(name):
    g = np.load(os.path.join(GOLDEN, name))
    return g, json.loads(str(g["cfg"])), json.loads(str(g["keys"])), int(g["seed"])


def _write_model_dir(path, net, sd, dataset, select):
    os.makedirs(path, exist_ok=True)
    torch.save({"net_param": dict(net, device="cuda"), "state_dict": sd}, os.path.join(path, "model_trained"))
    with open(os.path.join(path, "model_trained.yaml"), "w") as f:
        yaml.safe_dump({"net": net, "dataset": dataset, "train": {"train_model_select": select}}, f)


def test_diffusionts_uncertainty_ews_end_to_end(tmp_path):
    U = _U()
    g, cfg, shapes, seed = _load("dts_small_evalstep.npz")
    net = dict(cfg, task_model="DiffusionTS", n_z_samples=6, parallel_sample=3, diffusion_steps=4)
    from updgm_b200.diffusionts import DiffusionTS_model
    proto = DiffusionTS_model(dict(net, device="cpu"))
    sd = proto.state_dict()
    sd.update(dto.synth_state_dict(shapes, seed))
    sd["scaler_mean"], sd["scaler_std"] = torch.tensor([0.1, -0.2]), torch.tensor([1.5, 0.7])
    mdir = str(tmp_path / "dts")
    _write_model_dir(mdir, net, sd, {"windows": 24, "pred_len": 24, "sampling_t": 0.1}, "DiffusionTS_model")
    gen = torch.Generator().manual_seed(2)
    series = torch.tanh(torch.randn(3, 80, 2, generator=gen).cumsum(1) * 0.1)
    res = U.uncertainty_ews(model_save_file=mdir, torch_time_series=series, time_data=np.arange(80) * 0.1,
                            dynamic_type="SIS", sample_window_step=8, device=DEV, cache_path=str(tmp_path / "c"))
    W = (80 - 24) // 8 + 1
    assert res["task_model"] == "DiffusionTS" and len(res["ews"]) == W
    el = res["pred_future_list"][2]
    assert tuple(el.shape) == (3, 24, 2, 6) and torch.isfinite(el).all()
    pm, mpv = mpv_oracle.network_mpv(el.numpy(), sd["scaler_mean"].numpy(), sd["scaler_std"].numpy())
    assert float(res["ews"][2]) == pytest.approx(mpv, rel=1e-4) and float(res["pred_mean"][2]) == pytest.approx(pm, rel=1e-4)
    again = U.uncertainty_ews(model_save_file=mdir, torch_time_series=series, time_data=np.arange(80) * 0.1,
                              dynamic_type="SIS", device=DEV, cache_path=str(tmp_path / "c"))
    assert again["sample_window_step"] == 8 and torch.equal(again["pred_future_list"][2], el)      # cache round trip


def test_diffstg_uncertainty_ews_end_to_end(tmp_path):
    import networkx as nx
    U = _U()
    g, cfg, shapes, seed = _load("stg_small_evalstep.npz")
    net = dict(cfg, task_model="DiffSTG")
    from updgm_b200.diffstg import DiffSTG
    sd = DiffSTG(dict(net, device="cpu")).state_dict()
    w = dto.synth_state_dict(shapes, seed)
    for k in shapes:
        if ".net.0." in k:
            w[k] = w[k.replace(".net.0.", ".conv.")]
    sd.update(w)
    sd["scaler_mean"], sd["scaler_std"] = torch.tensor([0.3]), torch.tensor([2.0])
    mdir = str(tmp_path / "stg")
    _write_model_dir(mdir, net, sd, {"windows": 12, "pred_len": 8, "sampling_t": 0.1, "interval_step": 6}, "DiffSTG")
    G = nx.cycle_graph(5)
    G.add_edge(0, 2)
    nx.write_graphml(G, tmp_path / "g.graphml")
    gen = torch.Generator().manual_seed(4)
    series = torch.randn(5, 60, 1, generator=gen).cumsum(1) * 0.2
    res = U.uncertainty_ews(model_save_file=mdir, torch_time_series=series, time_data=np.arange(60) * 0.1,
                            dynamic_type="biomass", graph_file=str(tmp_path / "g.graphml"), device=DEV,
                            cache_path=str(tmp_path / "c"),
                            infer_params={"parallel_sampling": 3, "sequential_sampling": 2})
    W = (60 - 12) // 6 + 1                       # DiffSTG default step = dataset.interval_step
    assert res["sample_window_step"] == 6 and len(res["ews"]) == W and res["graph_file"].endswith("g.graphml")
    el = res["pred_future_list"][1]
    assert tuple(el.shape) == (5, 8, 1, 6) and torch.isfinite(el).all()
    pm, mpv = mpv_oracle.network_mpv(el.numpy(), sd["scaler_mean"].numpy(), sd["scaler_std"].numpy())
    assert float(res["ews"][1]) == pytest.approx(mpv, rel=1e-4) and float(res["pred_mean"][1]) == pytest.approx(pm, rel=1e-4)
    with pytest.raises(ValueError, match="graph_file is required"):
        U.uncertainty_ews(model_save_file=mdir, torch_time_series=series, time_data=np.arange(60) * 0.1,
                          dynamic_type="biomass", device=DEV, cache_path=str(tmp_path / "c2"))
