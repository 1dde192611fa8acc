"""The CPU oracle against fixtures produced by the reference itself (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import mpv_oracle, nsdiff_oracle, sigma_oracle, tmdm_oracle, windows_oracle

torch.set_num_threads(1)


class Replay:
    def __init__(self, noise):
        self.noise = list(noise)

    def __call__(self, like):
        z = self.noise.pop(0)
        assert z.shape == like.shape
        return z


def test_schedule_tables_bit_exact():
    g = load_golden("nsdiff_schedule_T20_linear.npz")
    s = nsdiff_oracle.nsdiff_schedule("linear", 20, 1e-4, 0.02)
    for k, v in g.items():
        assert torch.equal(s[k], v), k


def test_schedule_known_answers_from_baseline_md():
    # BASELINE.md "Known-answer values recorded from the reference"
    s = nsdiff_oracle.nsdiff_schedule("linear", 20, 1e-4, 0.02)
    np.testing.assert_allclose(s["betas_tilde"][[0, 1, 2, 18, 19]].numpy(),
                               [1.0001659393e-04, 1.2459754944e-03, 3.4332275391e-03, 1.6435050964e-01,
                                1.8066406250e-01], rtol=1e-7)
    np.testing.assert_allclose(s["alphas_cumprod_sum"][[0, 19]].numpy(), [0.9998999834, 17.3912906647], rtol=1e-7)
    np.testing.assert_allclose(s["one_minus_alphas_bar_sqrt"][[0, 19]].numpy(),
                               [1.0000829585e-02, 4.2804551125e-01], rtol=1e-7)
    assert s["betas_tilde_m_1"][0] == 1.0 and s["betas_bar_m_1"][0] == 1.0


def test_denoiser_matches_reference(wo_fx):
    _, sd = wo_fx
    g = load_golden("denoiser_wo_fx.npz")
    for t in (0, 1, 10, 19):
        eps, sig = nsdiff_oracle.denoiser_forward(sd, g["y_t"], g["y_0_hat"], g["gx"], t)
        assert torch.equal(eps, g["eps_t%d" % t])
        assert torch.equal(sig, g["sigma_t%d" % t])


@pytest.mark.parametrize("name", ["psample_loop_wo_fx.npz", "psample_loop_wo_fx_fx.npz"])
def test_p_sample_loop_matches_reference(wo_fx, name):
    _, sd = wo_fx
    g = load_golden(name)
    sched = nsdiff_oracle.nsdiff_schedule()
    draw = Replay(g["noise"])
    seq = nsdiff_oracle.p_sample_loop(sd, sched, g["y_0_hat"], g["gx"], g["y_0_hat"], 20, draw)
    assert not draw.noise, "exactly T draws per chunk"
    assert len(seq) == 21
    assert torch.equal(torch.stack(seq), g["seq"])


def test_p_sample_loop_random_weights_F1():
    g = load_golden("psample_loop_randF1.npz")
    seq = nsdiff_oracle.p_sample_loop(g["sd"], nsdiff_oracle.nsdiff_schedule(), g["y_0_hat"], g["gx"],
                                      g["y_0_hat"], 20, Replay(g["noise"]))
    assert torch.equal(torch.stack(seq), g["seq"])


def test_sigma_estimation_matches_reference(wo_fx):
    net_param, sd = wo_fx
    g = load_golden("sigma_estimation_wo_fx.npz")
    gx = sigma_oracle.sigma_estimation(sd, g["x"], net_param["rolling_length"], net_param["pred_len"])
    assert torch.equal(gx, g["gx"])


def test_evaluation_step_matches_reference(wo_fx):
    net_param, sd = wo_fx
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    net_param = dict(net_param, n_z_samples=8, parallel_sample=4)
    scaled = (g["window_raw"] - g["scaler_mean"]) / g["scaler_std"]
    assert torch.equal(scaled.unsqueeze(0), g["window_scaled"])
    draw = Replay(g["noise"].reshape(-1, 4, 200, 2))
    outs = nsdiff_oracle.evaluation_step(sd, net_param, g["window_scaled"], draw=draw)
    assert outs.shape == (1, 200, 2, 8)
    assert outs.stride() == g["outs"].permute(0, 3, 1, 2).contiguous().permute(0, 2, 3, 1).stride()
    assert torch.equal(outs, g["outs"])


def test_mpv_reductions_match_reference():
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    outs = g["outs"].numpy()
    elem = outs[0]
    assert mpv_oracle.slbp_mpv(elem, 0) == pytest.approx(float(g["slbp_mpv_dim0"]), rel=2e-6)
    assert mpv_oracle.slbp_mpv(elem, 1) == pytest.approx(float(g["fig6_mpv_dim1"]), rel=2e-6)
    assert mpv_oracle.intrinsic_dimension(elem) == int(g["fig6_intrinsic_dim"])
    mean, std = g["scaler_mean"].numpy(), g["scaler_std"].numpy()
    pm, mpv = mpv_oracle.network_mpv(outs, mean, std)
    assert pm == pytest.approx(float(g["net_pred_mean"]), rel=2e-6)
    assert mpv == pytest.approx(float(g["net_ews"]), rel=2e-6)
    pm, mpv = mpv_oracle.network_mpv(outs)
    assert pm == pytest.approx(float(g["net_pred_mean_nomodel"]), rel=1e-5, abs=1e-8)
    assert mpv == pytest.approx(float(g["net_ews_nomodel"]), rel=2e-6)
    tgt = (g["target_raw"] - g["scaler_mean"]) / g["scaler_std"]
    assert mpv_oracle.slbp_prediction_error(elem, tgt.numpy(), 0) == pytest.approx(float(g["slbp_err_dim0"]), rel=2e-6)
    gpm, gews = mpv_oracle.gx_ews(g["gx"].numpy(), pred_dim=1)
    assert gpm == pytest.approx(float(g["g_pred_mean"]), rel=2e-6)
    assert gews == pytest.approx(float(g["g_ews_dim1"]), rel=2e-6)
    assert mpv_oracle.gx_ews(g["gx"].numpy(), pred_dim=0)[1] == pytest.approx(float(g["gx_fig6_dim0"]), rel=2e-6)


def test_tmdm_loop_matches_reference():
    g = load_golden("tmdm_loop_randF1.npz")
    sched = tmdm_oracle.tmdm_schedule("linear", 20, 1e-4, 0.02)
    assert torch.equal(sched["alphas"], g["alphas"])
    assert torch.equal(sched["one_minus_alphas_bar_sqrt"], g["one_minus_alphas_bar_sqrt"])
    seq = tmdm_oracle.p_sample_loop(g["sd"], sched, g["y_0_hat"], g["y_0_hat"], 20, Replay(g["noise"]))
    assert torch.equal(torch.stack(seq), g["seq"])


def test_window_bookkeeping_bit_exact():
    with open(os.path.join(GOLDEN, "windows.json")) as f:
        cases = json.load(f)
    seen = set()
    for c in cases:
        seen.add(c["kind"])
        if c["kind"] == "interval":
            assert windows_oracle.sampling_interval_from_t(c["sampling_t"]) == c["interval"]
        elif c["kind"] == "infer_step":
            assert windows_oracle.infer_sample_window_step_from_cache(*c["args"]) == c["step"]
        elif c["kind"] == "count":
            assert windows_oracle.sliding_window_count(*c["args"]) == c["count"]
        elif c["kind"] == "network":
            n, t, f = c["shape"]
            series = np.arange(n * t, dtype=np.float32).reshape(n, t, f)
            idx = windows_oracle.sample_indices(t, c["sampling_t"])
            assert len(idx) == c["sampled_len"]
            wins, tps = windows_oracle.build_sliding_windows(series[:, idx, :], (np.arange(t) * 0.1)[idx],
                                                             c["windows"], c["step"])
            assert len(wins) == c["n_windows"] and len(tps) == c["n_time_points"]
            assert [float(w[0, 0, 0]) for w in wins[:4]] + [float(wins[-1][0, 0, 0])] == c["first_elems"]
            assert float(wins[-1][2, -1, 0]) == c["last_elem_of_last"]
            assert [float(x) for x in tps[:3]] == c["time_points_head"] and float(tps[-1]) == c["time_points_tail"]
        elif c["kind"] == "slbp":
            t, f = c["shape"]
            raw = np.arange(t * f, dtype=np.float32).reshape(t, f)
            ins, tgts, tps = windows_oracle.build_slbp_windows(raw, np.arange(t), c["windows"], c["pred_len"],
                                                               c["sampling_t"], c["step"])
            assert len(ins) == c["n_inputs"] and len(tgts) == c["n_targets"] and len(tps) == c["n_time_points"]
            assert [float(w[0, 0]) for w in ins[:3]] == c["input_first"]
            assert [float(w[0, 0]) for w in tgts[:3]] == c["target_first"]
            assert [int(x) for x in tps[:3]] == c["time_points_head"]
    assert seen == {"interval", "infer_step", "count", "network", "slbp"}
