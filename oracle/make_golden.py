"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py

Every array written here is an input to, or an output of, reference code imported from
/root/reference through oracle/ref_harness.py.  The GPU box has no reference tree, so these
fixtures are what pins both the oracle (CPU tests) and the CUDA path (GPU tests).
"""
import json
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def npz(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().contiguous().numpy()
        out[k] = v
    path = os.path.join(GOLD, name)
    np.savez(path, **out)
    print("wrote", path, "{:.1f} KB".format(os.path.getsize(path) / 1024))


def sd_arrays(sd, prefix="sd:"):
    return {prefix + k: v for k, v in sd.items()}


def main():
    os.makedirs(GOLD, exist_ok=True)
    dmu = ref_harness.analysis_module()
    torch.set_num_threads(1)  # fixtures must not depend on the thread count of the build box
    cpu = torch.device("cpu")

    # ---- shipped checkpoint: copied as data (bench config[0] / smoke / loader tests use it) ----
    for rel in ("ews_results/NsDiff_machine/wo_fx/model_trained",
                "ews_results/NsDiff_machine/wo_fx/model_trained.yaml",
                # training YAMLs only (their checkpoints are absent from the reference tree): they define the
                # architectures of BASELINE configs 2 and 3, instantiated with seeded random weights
                "ews_results/model_compare/NsDiff/biomass/model_trained.yaml",
                "ews_results/model_compare/TMDM/neuronal/model_trained.yaml"):
        dst = os.path.join(GOLD, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref_harness.REFERENCE_ROOT, rel), dst)
        os.chmod(dst, 0o644)

    model, net_param = dmu.load_model_from_dir(
        os.path.join(ref_harness.REFERENCE_ROOT, "ews_results/NsDiff_machine/wo_fx"), device=cpu)
    net = model.model  # NsDiff_net
    T = net.num_timesteps

    # ---- A. schedule tables (NsDiff_net.py:92-137) ----
    npz("nsdiff_schedule_T20_linear.npz",
        betas=net.betas, alphas=net.alphas, one_minus_alphas_bar_sqrt=net.one_minus_alphas_bar_sqrt,
        alphas_cumprod=net.alphas_cumprod, alphas_cumprod_sum=net.alphas_cumprod_sum,
        alphas_hat=net.alphas_hat, alphas_cumprod_prev=net.alphas_cumprod_prev,
        alphas_cumprod_sum_prev=net.alphas_cumprod_sum_prev, betas_tilde=net.betas_tilde,
        betas_bar=net.betas_bar, betas_tilde_m_1=net.betas_tilde_m_1, betas_bar_m_1=net.betas_bar_m_1)

    sched_args = (net.alphas, net.one_minus_alphas_bar_sqrt, net.alphas_cumprod, net.alphas_cumprod_sum,
                  net.alphas_cumprod_prev, net.alphas_cumprod_sum_prev, net.betas_tilde, net.betas_bar,
                  net.betas_tilde_m_1, net.betas_bar_m_1)
    from models.Diffusion_model.NsDiff import nsdiff_utils as nu

    # ---- F. g(x) on seeded windows (g_backbone.py:49-72) ----
    g = torch.Generator().manual_seed(11)
    xg = torch.randn(3, 200, 2, generator=g).cumsum(dim=1) * 0.1
    with torch.no_grad():
        gx3 = model.cond_pred_model_g(xg)
    npz("sigma_estimation_wo_fx.npz", x=xg, gx=gx3)

    # ---- B. denoiser (denoise.py:35-51) ----
    g = torch.Generator().manual_seed(12)
    y_t = torch.randn(3, 50, 2, generator=g)
    y0h = torch.randn(3, 50, 2, generator=g) * 0.3
    gxs = torch.rand(3, 50, 2, generator=g) * 0.07 + 0.01
    den = {"y_t": y_t, "y_0_hat": y0h, "gx": gxs}
    with torch.no_grad():
        for t in (0, 1, 10, 19):
            e, s = net(y_t, y0h, gxs, torch.tensor([t]))
            den["eps_t%d" % t] = e
            den["sigma_t%d" % t] = s
    npz("denoiser_wo_fx.npz", **den)

    # ---- C/D. p_sample / p_sample_t_1to0 / p_sample_loop with recorded noise ----
    g = torch.Generator().manual_seed(13)
    y0h = torch.zeros(6, 40, 2)
    gxl = torch.rand(6, 40, 2, generator=g) * 0.06 + 0.01
    tape = ref_harness.NoiseTape()
    torch.manual_seed(1234)
    with torch.no_grad(), tape.patched():
        seq = nu.p_sample_loop(net, y0h, gxl, y0h, T, *sched_args)
    npz("psample_loop_wo_fx.npz", y_0_hat=y0h, gx=gxl, noise=torch.stack(tape.draws), seq=torch.stack(seq))

    # same with a non-zero condition mean (f(x) present in the full NsDiff model)
    y0h2 = torch.randn(6, 40, 2, generator=g) * 0.5
    tape = ref_harness.NoiseTape()
    with torch.no_grad(), tape.patched():
        seq2 = nu.p_sample_loop(net, y0h2, gxl, y0h2, T, *sched_args)
    npz("psample_loop_wo_fx_fx.npz", y_0_hat=y0h2, gx=gxl, noise=torch.stack(tape.draws), seq=torch.stack(seq2))

    # ---- E. evaluation_step on a raw SLBP-like window, K=8, S=4 + the reductions ----
    model8, np8 = dmu.load_model_from_dir(
        os.path.join(ref_harness.REFERENCE_ROOT, "ews_results/NsDiff_machine/wo_fx"), device=cpu,
        infer_params={"n_z_samples": 8, "parallel_sample": 4})
    g = torch.Generator().manual_seed(14)
    ar = torch.zeros(400, 2)
    e = torch.randn(400, 2, generator=g) * 0.1
    for i in range(1, 400):
        ar[i] = 0.99 * ar[i - 1] + e[i]
    raw = ar * model8.scaler_std + model8.scaler_mean          # [400,2] raw units
    win_raw = raw[:200]
    tape = ref_harness.NoiseTape()
    with torch.no_grad(), tape.patched():
        scaled = model8.scaler_transform(win_raw).clone().unsqueeze(0)
        outs, _ = model8.evaluation_step(scaled)                 # [1,200,2,8]
        gx1 = model8.cond_pred_model_g(scaled)
    noise = torch.stack(tape.draws).reshape(2, 20, 4, 200, 2)    # [chunk, draw, B*S, O, F]
    slbp_elem = outs.squeeze(0)                                  # [O,F,K] SLBP cache element
    mpv_slbp, err_slbp = dmu.summarize_slbp_sensitivity([slbp_elem], [raw[200:400]], model=model8, device=cpu,
                                                        pred_dim=0)
    mpv6, dim6 = dmu.summarize_slbp_sampling_for_fig6([slbp_elem], pred_dim=1)
    pm_net, ews_net = dmu.summarize_pred_future_list([outs], model=model8)
    pm_net_nomodel, ews_net_nomodel = dmu.summarize_pred_future_list([outs], model=None)
    g_pm, g_ews = dmu.summarize_nsdiff_g_list([gx1], pred_dim=1)
    gx_fig6 = dmu.summarize_slbp_gx_for_fig6([gx1], pred_dim=0)
    npz("evalstep_wo_fx_k8s4.npz", window_raw=win_raw, window_scaled=scaled, noise=noise, outs=outs, gx=gx1,
        target_raw=raw[200:400],
        scaler_mean=model8.scaler_mean, scaler_std=model8.scaler_std,
        slbp_mpv_dim0=np.asarray(mpv_slbp[0]), slbp_err_dim0=np.asarray(err_slbp[0]),
        fig6_mpv_dim1=np.asarray(mpv6[0]), fig6_intrinsic_dim=np.asarray(dim6[0]),
        net_pred_mean=np.asarray(pm_net[0]), net_ews=np.asarray(ews_net[0]),
        net_pred_mean_nomodel=np.asarray(pm_net_nomodel[0]), net_ews_nomodel=np.asarray(ews_net_nomodel[0]),
        g_pred_mean=np.asarray(g_pm[0]), g_ews_dim1=np.asarray(g_ews[0]), gx_fig6_dim0=np.asarray(gx_fig6[0]))

    # ---- I. random-weight NsDiff denoiser, F=1 (BASELINE config 2 shape), B=3 rows, K=4, S=2 ----
    from models.Diffusion_model.NsDiff.NsDiff_net import NsDiff_net
    cfg = types.SimpleNamespace(diffusion_steps=20, diffusion_schedule="linear", beta_start=1e-4,
                                beta_end=0.02, dataset_nf=1)
    torch.manual_seed(123)
    net1 = NsDiff_net(cfg, cpu)
    sd1 = {"model." + k: v for k, v in net1.state_dict().items()}
    g = torch.Generator().manual_seed(15)
    y0h = torch.randn(6, 30, 1, generator=g) * 0.5 + 0.2          # rows = B*S = 3*2
    gx1f = torch.rand(6, 30, 1, generator=g) * 0.5 + 0.05
    s1 = (net1.alphas, net1.one_minus_alphas_bar_sqrt, net1.alphas_cumprod, net1.alphas_cumprod_sum,
          net1.alphas_cumprod_prev, net1.alphas_cumprod_sum_prev, net1.betas_tilde, net1.betas_bar,
          net1.betas_tilde_m_1, net1.betas_bar_m_1)
    tape = ref_harness.NoiseTape()
    with torch.no_grad(), tape.patched():
        seq = nu.p_sample_loop(net1, y0h, gx1f, y0h, 20, *s1)
    npz("psample_loop_randF1.npz", y_0_hat=y0h, gx=gx1f, noise=torch.stack(tape.draws), seq=torch.stack(seq),
        **sd_arrays(sd1))

    # ---- H. TMDM sampler with random weights (tmdm_model.py / tmdm_diffusion_utils.py) ----
    from models.Diffusion_model.TMDM import tmdm_diffusion_utils as tu
    from models.Diffusion_model.TMDM.tmdm_model import ConditionalGuidedModel as TmdmDenoiser
    dcfg = types.SimpleNamespace(diffusion=types.SimpleNamespace(timesteps=20),
                                 model=types.SimpleNamespace(cat_x=True, cat_y_pred=True))
    torch.manual_seed(321)
    tden = TmdmDenoiser(dcfg, types.SimpleNamespace(enc_in=1))

    class _Shim(torch.nn.Module):  # TMDM.forward (TMDM.py:94-98) minus the dead DataEmbedding of x
        def __init__(self, d):
            super().__init__()
            self.diffussion_model = d

        def forward(self, x, x_mark, y, y_t, y_0_hat, t):
            if t.shape[0] != y_t.shape[0]:
                pass
            return self.diffussion_model(None, y_t, y_0_hat, t)

    shim = _Shim(tden)
    betas = tu.make_beta_schedule("linear", 20, 1e-4, 0.02).float()
    alphas = 1.0 - betas
    om = torch.sqrt(1 - alphas.cumprod(dim=0))
    g = torch.Generator().manual_seed(16)
    y0t = torch.randn(5, 24, 1, generator=g) * 0.4
    tape = ref_harness.NoiseTape()
    with torch.no_grad(), tape.patched():
        tseq = tu.p_sample_loop(shim, None, None, y0t, y0t, 20, alphas, om)
    npz("tmdm_loop_randF1.npz", y_0_hat=y0t, noise=torch.stack(tape.draws), seq=torch.stack(tseq),
        alphas=alphas, one_minus_alphas_bar_sqrt=om,
        **sd_arrays({"model." + k: v for k, v in shim.state_dict().items()}))

    # ---- G. window / step bookkeeping (integer) ----
    cases = []
    series = torch.arange(3 * 100000, dtype=torch.float32).reshape(3, 100000, 1)
    tdata = np.arange(100000) * 0.1
    ss, st = dmu.sample_time_series(series, tdata, sampling_t=10)
    wins, tps = dmu.build_sliding_windows(ss, st, windows=100, sample_window_step=5)
    cases.append({"kind": "network", "shape": [3, 100000, 1], "sampling_t": 10, "windows": 100, "step": 5,
                  "n_windows": len(wins), "sampled_len": int(ss.shape[1]),
                  "first_elems": [float(w[0, 0, 0]) for w in wins[:4]] + [float(wins[-1][0, 0, 0])],
                  "last_elem_of_last": float(wins[-1][2, -1, 0]),
                  "time_points_head": [float(x) for x in tps[:3]], "time_points_tail": float(tps[-1]),
                  "n_time_points": int(len(tps))})
    for st_ in (0.05, 0.1, 0.3, 0.7, 1, 10, 100, 2.9999):
        cases.append({"kind": "interval", "sampling_t": st_, "interval": dmu.sampling_interval_from_t(st_)})
    for args in ((1000, 100, 181, 10), (10000, 200, 981, 5), (10000, 200, 981, 10), (1000, 100, 19, 5),
                 (1000, 100, 1, 5), (50, 100, 3, 5), (1000, 100, 0, 5), (1000, 100, 226, 5), (997, 100, 300, 7)):
        cases.append({"kind": "infer_step", "args": list(args),
                      "step": int(dmu.infer_sample_window_step_from_cache(*args))})
    for args in ((1000, 100, 5), (99, 100, 5), (100, 100, 5), (10000, 200, 10)):
        cases.append({"kind": "count", "args": list(args), "count": int(dmu.sliding_window_count(*args))})
    raw = torch.arange(2 * 1000000, dtype=torch.float32).reshape(1000000, 2)
    ins, tgts, tps = dmu.build_slbp_sensitivity_windows(raw, np.arange(1000000), 200, 200, 100, 10)
    cases.append({"kind": "slbp", "shape": [1000000, 2], "windows": 200, "pred_len": 200, "sampling_t": 100,
                  "step": 10, "n_inputs": len(ins), "n_targets": len(tgts),
                  "input_first": [float(w[0, 0]) for w in ins[:3]], "target_first": [float(w[0, 0]) for w in tgts[:3]],
                  "time_points_head": [int(x) for x in tps[:3]], "n_time_points": int(len(tps))})
    with open(os.path.join(GOLD, "windows.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote windows.json")


if __name__ == "__main__":
    main()
