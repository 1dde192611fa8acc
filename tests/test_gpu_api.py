"""GPU tests of the drop-in surface: the reference's checkpoint, YAML, inputs and cache files through
the reference-named API, checked against reference-made fixtures and the CPU oracle."""
import os
import shutil

import numpy as np
import pytest
import torch
import yaml

from conftest import GOLDEN, load_golden, load_wo_fx_checkpoint
from oracle import fx_oracle, mpv_oracle, nsdiff_oracle, sigma_oracle, tmdm_oracle

pytestmark = pytest.mark.gpu

CKPT_DIR = os.path.join(GOLDEN, "ews_results", "NsDiff_machine", "wo_fx")


def _U():
    from updgm_b200 import uncertainty
    return uncertainty


@pytest.fixture()
def model8():
    m, p = _U().load_model_from_dir(CKPT_DIR, device=torch.device("cuda:0"),
                                    infer_params={"n_z_samples": 8, "parallel_sample": 4})
    return m, p


def _close(out, ref, tol=2e-5):
    out, ref = out.double().cpu(), ref.double().cpu()
    rms = ref.pow(2).mean().sqrt()
    assert float((out - ref).abs().max() / rms) <= tol


def test_load_model_from_dir_shipped_checkpoint(model8):
    m, p = model8
    _, sd = load_wo_fx_checkpoint()
    assert p["task_model"] == "NsDiff_model_variants" and p["n_z_samples"] == 8 and p["parallel_sample"] == 4
    assert not m.training and m.scaler == "StandardScaler"
    own = m.state_dict()
    assert set(own) == set(sd)
    for k, v in sd.items():
        assert torch.equal(own[k].cpu(), v), k
    assert m.scaler_mean.device.type == "cuda"
    np.testing.assert_allclose(m.scaler_mean.cpu().numpy(), [47.3036, 1.4155], rtol=1e-4)


def test_evaluation_step_matches_reference_with_injected_noise(model8):
    m, _ = model8
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    scaled = m.scaler_transform(g["window_raw"].cuda()).clone().unsqueeze(0)
    assert torch.equal(scaled.cpu(), g["window_scaled"])
    outs, batch_y = m.evaluation_step(scaled, noise=g["noise"])
    assert batch_y is None and outs.device.type == "cpu" and outs.dtype == torch.float32
    assert tuple(outs.shape) == (1, 200, 2, 8)
    assert outs.stride() == (8 * 200 * 2, 2, 1, 200 * 2), "permuted view of contiguous [B,K,O,F] like the reference"
    _close(outs, g["outs"])
    gx = m.cond_pred_model_g(scaled)
    np.testing.assert_allclose(gx.cpu().numpy(), g["gx"].numpy(), rtol=2e-5, atol=1e-7)
    # a window that also carries the horizon returns batch_y like the reference (:406-410)
    long = torch.cat([scaled, scaled], dim=1)
    _, by = m.evaluation_step(long, noise=g["noise"])
    assert by is not None and tuple(by.shape) == (1, 200, 2)


def test_slbp_sensitivity_ews_end_to_end(tmp_path, model8):
    """slbp_sensitivity_ews on a synthetic SLBP series: cache file, MPV and prediction error against the
    oracle applied to the very cache the GPU wrote (reduction parity), windows/time points bit-exact."""
    U = _U()
    root = tmp_path / "NsDiff_x"
    (root / "models").mkdir(parents=True)
    shutil.copyfile(os.path.join(CKPT_DIR, "model_trained"), root / "models" / "dataset_w200p200st100")
    shutil.copyfile(os.path.join(CKPT_DIR, "model_trained.yaml"), root / "models" / "dataset_w200p200st100.yaml")
    g = torch.Generator().manual_seed(3)
    n_raw = 460 * 1000
    ar = torch.zeros(460, 2)
    e = torch.randn(460, 2, generator=g) * 0.1
    for i in range(1, 460):
        ar[i] = 0.99 * ar[i - 1] + e[i]
    _, sd = load_wo_fx_checkpoint()
    sampled = ar * sd["scaler_std"] + sd["scaler_mean"]
    raw = sampled.repeat_interleave(1000, dim=0)                       # sampling_t=100 keeps every 1000th row
    tdata = np.arange(n_raw) * 0.1
    res = U.slbp_sensitivity_ews(str(root), "dataset_w200p200st100", raw, tdata, data_trend="increase", pred_dim=1,
                                 sample_window_step=10, infer_params={"n_z_samples": 16, "parallel_sample": 8},
                                 device=torch.device("cuda:0"))
    W = (460 - 200) // 10 + 1
    assert len(res["pred_future_list"]) == W == len(res["mpv"]) and len(res["prediction_error"]) == 7
    assert res["cache_path"].endswith("datas/dataset_w200p200st100_pred_future_increase_10.pt")
    assert list(res["time_points"][:2]) == [tdata[199 * 1000], tdata[209 * 1000]]
    cached = U._load_tensor_list(res["cache_path"])
    assert len(cached) == W and tuple(cached[0].shape) == (200, 2, 16)
    for w in (0, 5, W - 1):
        assert torch.equal(cached[w], res["pred_future_list"][w])
        assert float(res["mpv"][w]) == pytest.approx(mpv_oracle.slbp_mpv(cached[w].numpy(), 1), rel=1e-5)
    tgt = (sampled[200:400] - sd["scaler_mean"]) / sd["scaler_std"]
    assert float(res["prediction_error"][0]) == pytest.approx(
        mpv_oracle.slbp_prediction_error(cached[0].numpy(), tgt.numpy(), 1), rel=1e-5)
    assert res["mpv"][0].dtype == np.float32 and res["mpv"][0].shape == ()
    # second call reads the cache (no model needed) and reproduces the numbers through the re-upload path
    res2 = U.slbp_sampling_analysis(str(root), "dataset_w200p200st100", raw, tdata, pred_dim=1, sample_window_step=10)
    assert res2["available"] and len(res2["mpv"]) == W
    assert res2["mpv"][3] == pytest.approx(float(res["mpv"][3]), rel=1e-6)
    assert res2["intrinsic_dimension"][0] == mpv_oracle.intrinsic_dimension(cached[0].numpy())
    gxr = U.slbp_gx_analysis(str(root), "dataset_w200p200st100", raw, tdata, pred_dim=0, sample_window_step=10,
                             device=torch.device("cuda:0"))
    assert gxr["cache_path"].endswith("_pred_future_increase_10_gx.pt") and len(gxr["gx_mpv"]) == W
    x0 = ((sampled[:200] - sd["scaler_mean"]) / sd["scaler_std"]).unsqueeze(0)
    ref_gx = sigma_oracle.sigma_estimation(sd, x0, 100, 200)
    assert gxr["gx_mpv"][0] == pytest.approx(float(ref_gx[0, :, 0].mean()), rel=2e-5)
    mp = U.slbp_mpv_analysis(str(root), "dataset_w200p200st100", raw, tdata, cache_path=res["cache_path"], pred_dim=1,
                             sample_window_step=5)
    assert mp["sample_window_step"] == 10 and mp["uncertainty_source"] == "sampling"      # step recovered from the cache


def test_uncertainty_ews_network_sweep_with_gx(tmp_path):
    """uncertainty_ews over a 3-node synthetic network series with the F=2 shipped model is not a physical
    configuration, so build a model dir from the shipped checkpoint and feed [Node,T,F] directly."""
    U = _U()
    mdir = tmp_path / "model"
    shutil.copytree(CKPT_DIR, mdir)
    cfg = yaml.safe_load(open(mdir / "model_trained.yaml"))
    cfg["dataset"]["sampling_t"] = 0.1
    yaml.safe_dump(cfg, open(mdir / "model_trained.yaml", "w"))
    g = torch.Generator().manual_seed(9)
    _, sd = load_wo_fx_checkpoint()
    series = (torch.randn(3, 260, 2, generator=g) * 0.1).cumsum(dim=1) * sd["scaler_std"] + sd["scaler_mean"]
    tdata = np.arange(260) * 0.1
    res = U.uncertainty_ews(model_save_file=str(mdir), torch_time_series=series, time_data=tdata, dynamic_type="SLBP",
                            sample_window_step=20, infer_params={"n_z_samples": 12, "parallel_sample": 5},
                            uncertainty_method="both", device=torch.device("cuda:0"),
                            cache_path=str(tmp_path / "cache"))
    W = (260 - 200) // 20 + 1
    assert len(res["ews"]) == W and len(res["pred_mean"]) == W and len(res["nsdiff_g"]["ews"]) == W
    el = res["pred_future_list"][1]
    assert tuple(el.shape) == (3, 200, 2, 10), "K floors to (12 // 5) * 5 like the reference"
    mean, std = sd["scaler_mean"].numpy(), sd["scaler_std"].numpy()
    pm, mpv = mpv_oracle.network_mpv(el.numpy(), mean, std)          # fresh compute => raw units (model present)
    assert float(res["ews"][1]) == pytest.approx(mpv, rel=1e-5) and float(res["pred_mean"][1]) == pytest.approx(pm, rel=1e-5)
    assert res["cache_path"].endswith("cache/data.pt") and (U.flush_cache_writes() or os.path.exists(res["cache_path"]))
    assert res["nsdiff_g"]["cache_path"].endswith("cache/data_gx.pt")
    assert list(res["time_points"]) == [tdata[199 + 20 * i] for i in range(W)]
    # cache-only read: no model -> statistics in normalised units, step inferred from the cache
    res2 = U.uncertainty_ews(model_save_file=str(mdir), torch_time_series=series, time_data=tdata, dynamic_type="SLBP",
                             cache_path=str(tmp_path / "cache"), uncertainty_method="sampling", save_nsdiff_g=False)
    # 4 windows over a span of 60 fit steps 16..20; the reference's tie-break picks the one closest to the default 10
    assert res2["sample_window_step"] == 16 and res2["loaded_net_param"] is None
    _, mpv_n = mpv_oracle.network_mpv(el.numpy())
    assert float(res2["ews"][1]) == pytest.approx(mpv_n, rel=1e-5)
    gxo = U.uncertainty_ews(model_save_file=str(mdir), torch_time_series=series, time_data=tdata, dynamic_type="SLBP",
                            cache_path=str(tmp_path / "cache"), uncertainty_method="gx")
    assert gxo["uncertainty_source"] == "gx" and gxo["pred_future_list"] is None and len(gxo["ews"]) == W


def test_philox_sweep_independent_of_batch_split(model8, monkeypatch):
    U = _U()
    m, _ = model8
    g = torch.Generator().manual_seed(5)
    stacked = (torch.randn(6, 1, 200, 2, generator=g) * 0.1).cumsum(dim=2)
    torch.manual_seed(77)
    m._windows_drawn = 0
    a = U.sample_sweep(m, stacked).clone()
    m._windows_drawn = 0
    monkeypatch.setattr(U, "SWEEP_BATCH_BYTES", 2 * 8 * 200 * 2 * 4)       # two windows per launch
    b = U.sample_sweep(m, stacked)
    assert torch.equal(a, b)
    assert torch.equal(b.upd_stats["scaled"]["mpv"], a.upd_stats["scaled"]["mpv"]) if hasattr(a, "upd_stats") else True


def test_full_nsdiff_model_with_fx_random_weights():
    """BASELINE config-2 architecture (NsDiff task model, f(x) + g(x) + denoiser, F=1), seeded random weights:
    GPU evaluation_step vs the oracle with the same injected noise (f(x) restated in oracle/fx_oracle.py)."""
    from updgm_b200.nsdiff import NsDiff_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "NsDiff", "biomass", "model_trained.yaml")))
    net = dict(cfg["net"], device=torch.device("cuda:0"), n_z_samples=6, parallel_sample=3)
    torch.manual_seed(123)
    m = NsDiff_model(net, "NsDiff_model").eval()
    m.scaler_std.fill_(1.0)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(4, 100, 1, generator=g) * 0.1).cumsum(dim=1) + 5.0
    noise = torch.randn(2, 20, 4 * 3, 100, 1, generator=g)
    outs, _ = m.evaluation_step(x.cuda(), noise=noise)
    fx_sd = {k[len("cond_pred_model."):]: v for k, v in sd.items() if k.startswith("cond_pred_model.")}
    fx_cfg = dict(net, seq_len=100, label_len=50)
    with torch.no_grad():
        y0 = fx_oracle.ns_transformer(fx_sd, fx_cfg, x)[:, -100:, :]
        gx = sigma_oracle.sigma_estimation(sd, x, 50, 100)
        it = iter(noise.reshape(-1, 12, 100, 1))
        ref = nsdiff_oracle.evaluation_step(sd, net, x, y_0_hat=y0, gx=gx, draw=lambda like: next(it))
    with torch.no_grad():
        y0_gpu, gx_gpu = m.condition(x.cuda())
    np.testing.assert_allclose(y0_gpu.cpu().numpy(), y0.numpy(), rtol=2e-4, atol=2e-4)
    _close(outs, ref, tol=2e-4)          # f(x) runs as fp32 library GEMMs whose summation order differs from the CPU's


def test_tmdm_model_random_weights():
    from updgm_b200.tmdm import TMDM_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "TMDM", "neuronal", "model_trained.yaml")))
    net = dict(cfg["net"], device=torch.device("cuda:0"), n_z_samples=4, parallel_sample=2)
    torch.manual_seed(321)
    m = TMDM_model(net).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert sd["model.diffussion_model.lin1.embed.weight"].shape == (21, 128)
    g = torch.Generator().manual_seed(4)
    x = torch.sigmoid((torch.randn(3, 100, 1, generator=g) * 0.2).cumsum(dim=1))
    noise = torch.randn(2, 20, 3 * 2, 150, 1, generator=g)
    outs, _ = m.evaluation_step(x.cuda(), noise=noise)
    assert tuple(outs.shape) == (3, 100, 1, 4)
    fx_sd = {k[len("cond_pred_model."):]: v for k, v in sd.items() if k.startswith("cond_pred_model.")}
    fx_cfg = dict(m.configs.__dict__)
    with torch.no_grad():
        y0 = fx_oracle.ns_transformer(fx_sd, fx_cfg, x, vae=True)                       # [3,150,1]
        it = iter(noise.reshape(-1, 6, 150, 1))
        ref = tmdm_oracle.evaluation_step(sd, dict(net, beta_schedule="linear"), y0, draw=lambda like: next(it))
    _close(outs, ref, tol=2e-4)


def test_state_dict_roundtrip_and_repack(model8):
    m, _ = model8
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    a, _ = m.evaluation_step(g["window_scaled"].cuda(), noise=g["noise"])
    with torch.no_grad():
        m.model.diffussion_model.lin2.lin.weight.mul_(1.5)
    b, _ = m.evaluation_step(g["window_scaled"].cuda(), noise=g["noise"])
    assert not torch.equal(a, b), "changed parameters must be re-packed"
    m.load_state_dict(sd, strict=True)
    c, _ = m.evaluation_step(g["window_scaled"].cuda(), noise=g["noise"])
    assert torch.equal(a, c)


def test_real_data_gx_uncertainty_batched(tmp_path):
    """SURVEY 8f row 3 (real_data_analysis.run_model_uncertainty): all windows' g(x) in one launch == the per-window oracle."""
    from updgm_b200.nsdiff import NsDiff_model_variants
    U = _U()
    net_param, _ = load_wo_fx_checkpoint()
    net = dict(net_param, device=torch.device("cuda:0"), dataset_nf=1, windows=100, seq_len=100, pred_len=100,
               rolling_length=50, scaler_type="StandardScaler")
    torch.manual_seed(77)
    m = NsDiff_model_variants(net, "cond_var").eval()
    m.scaler_mean.fill_(0.4)
    m.scaler_std.fill_(1.7)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(21)
    series = (torch.randn(3, 3000, 1, generator=g) * 0.1).cumsum(dim=1) * 1.7 + 0.4
    tdata = torch.arange(3000) * 0.1
    times, values = U.real_data_gx_uncertainty(m, series, tdata, windows=100, sampling_t=1.0, sample_window_step=20,
                                               pred_dim=2, cache_path=tmp_path / "gx.pt")
    sub = series[:, ::10, :]
    W = (300 - 100) // 20 + 1
    assert len(values) == W and len(times) == W and float(times[0]) == pytest.approx(float(tdata[::10][99]))
    cached = U._load_tensor_list(tmp_path / "gx.pt")
    assert len(cached) == W and tuple(cached[0].shape) == (3, 100)
    for w in (0, W - 1):
        x = (sub[:, 20 * w: 20 * w + 100, :] - 0.4) / 1.7
        ref = sigma_oracle.sigma_estimation(sd, x, 50, 100).squeeze(-1)          # [Node, O]
        _close(cached[w], ref, tol=2e-5)
        assert values[w] == pytest.approx(float(ref.mean(dim=-1)[2]), rel=1e-4)
    m2, _ = U.load_model_from_dir(CKPT_DIR, device=torch.device("cuda:0"))         # F = 2: not a scalar per window
    with pytest.raises(TypeError):
        U.real_data_gx_uncertainty(m2, torch.zeros(1, 4000, 2), torch.arange(4000) * 0.1, 200, 1.0, 50, 0, tmp_path / "x.pt")


def test_distributed_sweep_single_rank_group_equals_plain_sweep(model8):
    """The N > 1 driver (bench.py under torchrun) on a one-rank gloo group: same cache and MPV list as sample_sweep, and the
    Philox window counter advances by the whole sweep."""
    import socket
    import torch.distributed as dist
    U = _U()
    m, _ = model8
    g = torch.Generator().manual_seed(6)
    stacked = (torch.randn(5, 1, 200, 2, generator=g) * 0.1).cumsum(dim=2)
    torch.manual_seed(78)
    m._windows_drawn = 0
    ref = U.sample_sweep(m, stacked)
    ref_stats = {k: v.clone() for k, v in ref.upd_stats.get("raw", ref.upd_stats["scaled"]).items()}
    ref = ref.clone()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
    try:
        m._windows_drawn = 0
        cache, (w0, w1), stats = U.distributed_sweep(m, stacked, device=torch.device("cuda:0"))
    finally:
        dist.destroy_process_group()
    assert (w0, w1) == (0, 5) and m._windows_drawn == 5
    assert torch.equal(cache, ref)
    assert torch.equal(stats["mpv"], ref_stats["mpv"]) and torch.equal(stats["pred_mean"], ref_stats["pred_mean"])
    assert tuple(stats["mpv_f"].shape) == (5, 2)


@pytest.mark.parametrize("K", [100, 7])
def test_slbp_extras_batched_on_device_against_oracle(K):
    """the intrinsic dimension of every window from ONE batched centred-Gram launch + one batched
    eigenvalue solve (integer, exact against the oracle's (O*F)^2 covariance form, diffusion_model_uncertainy.py:686-698),
    and the prediction error of every window from one launch on the Welford means (:542-549)."""
    U = _U()
    torch.manual_seed(40 + K)
    W, O, F = 9, 200, 2
    # K trajectories with a window-dependent number of dominant directions + noise
    cache = torch.empty(W, 1, K, O, F)
    for w in range(W):
        r = 1 + w % 5
        basis = torch.randn(r, O * F)
        cache[w, 0] = (torch.randn(K, r) * torch.linspace(3.0, 1.0, r) @ basis + 0.3 * torch.randn(K, O * F)).view(K, O, F)
    elems = [cache[w, 0].permute(1, 2, 0) for w in range(W)]                      # [O, F, K], as in an SLBP cache
    targets = [torch.randn(O, F) for _ in range(W)]
    mpv, dims = U.summarize_slbp_sampling_for_fig6(elems, pred_dim=1)
    assert dims == [mpv_oracle.intrinsic_dimension(e.numpy()) for e in elems]
    assert len(set(dims)) > 1
    mpv2, err = U.summarize_slbp_sensitivity(elems, targets, model=None, pred_dim=1)
    for w in range(W):
        assert float(err[w]) == pytest.approx(mpv_oracle.slbp_prediction_error(elems[w].numpy(), targets[w].numpy(), 1), rel=2e-6)
        assert float(mpv2[w]) == pytest.approx(mpv_oracle.slbp_mpv(elems[w].numpy(), 1), rel=2e-5)
        assert float(mpv[w]) == pytest.approx(float(mpv2[w]), rel=1e-6)
