"""Full-size parity harness: the oracle's evaluation_step on WHOLE windows of the BASELINE workloads with recorded
noise, and the comparison BASELINE.json's north_star states literally.  TEST INFRASTRUCTURE ONLY: imported by
``tests/test_gpu_parity_full.py`` and by ``bench.py``'s ``cpu_baseline`` leg (which emits the same check as the bench
line's ``parity`` object); never by the product package.

Tolerances (north_star: "rel 1e-4 on MPV, rel 1e-3 per trajectory value"), reference algorithm:
``models/Diffusion_model/NsDiff/nsdiff_utils.py:271-284`` (p_sample_loop), ``NsDiff_model.py:180-268, 404-495``
(evaluation_step) and ``evaluation_and_analysis/diffusion_model_uncertainy.py:286-303`` (MPV).
"""
import time

import numpy as np
import torch

from . import fx_oracle, mpv_oracle, nsdiff_oracle, sigma_oracle, tmdm_oracle

REL_PER_VALUE = 1e-3       # |gpu - ref| <= REL_PER_VALUE * |ref| + FLOOR_OF_RMS * rms(ref)
FLOOR_OF_RMS = 1e-4        # absolute floor for values near zero (a relative bound alone is undefined there)
REL_MPV = 1e-4


class RecordingDraw:
    """torch.randn_like stand-in: seeded draws, kept in call order (the order SURVEY A.4 documents)."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.draws = []

    def __call__(self, like):
        z = torch.randn(like.shape, generator=self.gen, dtype=like.dtype)
        self.draws.append(z)
        return z

    def noise_tensor(self, n_chunks, T):
        """[1, K/S, T, B*S, O, F]: the layout upd_*_sample's validation mode reads."""
        assert len(self.draws) == n_chunks * T, (len(self.draws), n_chunks, T)
        return torch.stack(self.draws).reshape(1, n_chunks, T, *self.draws[0].shape)


def nsdiff_window_reference(sd, net, x, seed=7, with_fx=True, variant_adds_eps=False, threads=None):
    """One whole window [B, L, F] through the oracle: f(x) (fx_oracle), g(x) (sigma_oracle), every chunk of the reverse
    loop with recorded draws.  -> dict(ref [B,O,F,K], noise [1,K/S,T,B*S,O,F], y0 [B,O,F], gx [B,O,F], seconds)."""
    if threads:
        torch.set_num_threads(threads)
    L, O, T = net["windows"], net["pred_len"], net["diffusion_steps"]
    S = int(net["parallel_sample"])
    n_chunks = int(net["n_z_samples"]) // S
    draw = RecordingDraw(seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        if with_fx:
            fx_sd = {k[len("cond_pred_model."):]: v for k, v in sd.items() if k.startswith("cond_pred_model.")}
            y0 = fx_oracle.ns_transformer(fx_sd, dict(net, seq_len=L, label_len=L // 2), x)[:, -O:, :]
        else:
            y0 = torch.zeros(x.shape[0], O, x.shape[2])
        gx = sigma_oracle.sigma_estimation(sd, x, net["rolling_length"], O)
        if variant_adds_eps:
            gx = gx + nsdiff_oracle.EPS
        ref = nsdiff_oracle.evaluation_step(sd, net, x, y_0_hat=y0, gx=gx, draw=draw, variant_adds_eps=False)
    return {"ref": ref, "noise": draw.noise_tensor(n_chunks, T), "y0": y0, "gx": gx, "seconds": time.perf_counter() - t0}


def tmdm_window_reference(sd, net, x, y0=None, seed=7, threads=None):
    """One whole TMDM window: condition mean (fx_oracle with the VAE head, unless ``y0`` [B, L/2+O, F] is given), the
    chunked reverse loop with recorded draws.  -> dict(ref [B,O,F,K], noise [1,K/S,T,B*S,L/2+O,F], y0, seconds)."""
    if threads:
        torch.set_num_threads(threads)
    T, K = net["diffusion_steps"], int(net["n_z_samples"])
    S = min(int(net["parallel_sample"]), K)
    draw = RecordingDraw(seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        if y0 is None:
            fx_sd = {k[len("cond_pred_model."):]: v for k, v in sd.items() if k.startswith("cond_pred_model.")}
            y0 = fx_oracle.ns_transformer(fx_sd, dict(net), x, vae=True)
        ref = tmdm_oracle.evaluation_step(sd, dict(net, beta_schedule=net.get("beta_schedule", "linear")), y0, draw=draw)
    return {"ref": ref, "noise": draw.noise_tensor(K // S, T), "y0": y0, "seconds": time.perf_counter() - t0}


def compare(got, ref):
    """``got`` / ``ref``: [B, O, F, K] (any device).  The north-star comparison, as numbers and a verdict."""
    got = got.detach().cpu().double()
    ref = ref.detach().cpu().double()
    rms = float(ref.pow(2).mean().sqrt())
    d = (got - ref).abs()
    bound = REL_PER_VALUE * ref.abs() + FLOOR_OF_RMS * rms
    worst = float((d / bound).max())
    _, mpv_ref = mpv_oracle.network_mpv(ref.numpy())
    _, mpv_got = mpv_oracle.network_mpv(got.numpy())
    mpv_rel = abs(float(mpv_got) - float(mpv_ref)) / abs(float(mpv_ref))
    return {"values": int(ref.numel()), "max_abs_err_over_rms": float(d.max()) / rms,
            "worst_err_over_bound": worst, "per_value_ok": bool(worst <= 1.0),
            "per_value_bound": "|d| <= {:g}*|ref| + {:g}*rms(ref)".format(REL_PER_VALUE, FLOOR_OF_RMS),
            "mpv_ref": float(mpv_ref), "mpv_rel_err": mpv_rel, "mpv_ok": bool(mpv_rel <= REL_MPV), "mpv_bound": REL_MPV,
            "finite": bool(np.isfinite(got.numpy()).all())}
