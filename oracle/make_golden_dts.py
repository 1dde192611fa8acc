"""Generate tests/golden/dts_*.npz by running the UNMODIFIED reference DiffusionTS (build container only).

    python oracle/make_golden_dts.py

The reference's DiffusionTS package imports without third-party stand-ins (torch, einops, scipy, tqdm).  No
DiffusionTS checkpoint ships with the reference, so weights are the deterministic
``diffusionts_oracle.synth_state_dict`` (numpy RandomState keyed by parameter name) loaded into the reference module;
fixtures store the seed and the key/shape list, not the weights.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_harness, diffusionts_oracle as dto  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

SMALL = dict(dataset_nf=2, windows=24, pred_len=24, n_z_samples=4, parallel_sample=2, diffusion_steps=10,
             d_model=32, timesteps=100, n_heads=4, n_layer_enc=2, n_layer_dec=2, beta_schedule="cosine",
             loss_type="l2", infill_coef=0.1, infill_learning_rate=0.05, eta=0.0, scaler_type="StandardScaler")
# ews_results/model_compare/DiffusionTS/SIS/model_trained.yaml (BASELINE config 4)
YAML = dict(dataset_nf=1, windows=100, pred_len=100, n_z_samples=100, parallel_sample=10, diffusion_steps=100,
            d_model=64, timesteps=100, n_heads=4, n_layer_enc=3, n_layer_dec=6, beta_schedule="cosine",
            loss_type="l2", infill_coef=0.1, infill_learning_rate=0.05, eta=0.0, scaler_type="StandardScaler",
            attn_pd=0.0, resid_pd=0.0, mlp_hidden_times=4, kernel_size=None, padding_size=None, use_ff=True,
            reg_weight=None)


class Tape:
    """Records (or replays) every torch.randn / torch.randn_like draw of the reference, in order."""

    def __init__(self):
        self.draws = []

    def _get(self, shape):
        z = torch.empty(tuple(shape)).normal_()
        self.draws.append(z.clone())
        return z

    def __enter__(self):
        self._rl, self._r = torch.randn_like, torch.randn
        torch.randn_like = lambda t, *a, **k: self._get(t.shape)
        torch.randn = lambda *a, **k: self._get(a[0] if len(a) == 1 and not isinstance(a[0], int) else a)
        return self

    def __exit__(self, *exc):
        torch.randn_like, torch.randn = self._rl, self._r


def float_param_shapes(model):
    return {k: tuple(v.shape) for k, v in model.state_dict().items() if k.startswith("model.model.")}


def build_reference(cfg, seed):
    from models.Diffusion_model.DiffusionTS.DiffusionTS_model import DiffusionTS_model
    m = DiffusionTS_model(dict(cfg, device="cpu")).eval()
    shapes = float_param_shapes(m)
    sd = dto.synth_state_dict(shapes, seed)
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(not k.startswith("model.model.") for k in missing.missing_keys), missing.missing_keys
    return m, shapes


def save(name, **arrays):
    out = {k: (v.detach().cpu().contiguous().numpy() if isinstance(v, torch.Tensor) else v) for k, v in arrays.items()}
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **out)
    print("wrote", path, "{:.1f} KB".format(os.path.getsize(path) / 1024))


def main():
    ref_harness.activate()
    torch.set_num_threads(1)
    # ---------- A: whole evaluation_step of a small architecture (kernel-size-1 combine branch, F=2, 2 chunks) ----------
    m, shapes = build_reference(SMALL, seed=11)
    torch.manual_seed(5)
    batch = torch.tanh(torch.randn(2, 24, 2).cumsum(1) * 0.2)
    with Tape() as tape:
        outs, _ = m.evaluation_step(batch)
    save("dts_small_evalstep.npz", cfg=json.dumps(SMALL), seed=11, keys=json.dumps({k: list(v) for k, v in shapes.items()}),
         batch=batch, outs=outs.contiguous(), n_draws=len(tape.draws),
         **{"z%03d" % i: z for i, z in enumerate(tape.draws)},
         **{"tab:" + k: getattr(m.model, k) for k in dto.schedule_buffers(100)})

    # ---------- B: BASELINE config-4 architecture: forward passes and single sampling steps ----------
    m, shapes = build_reference(YAML, seed=23)
    net = m.model
    torch.manual_seed(7)
    R = 3
    hist = torch.tanh(torch.randn(R, 100, 1).cumsum(1) * 0.1)
    target = torch.cat([hist, torch.zeros(R, 100, 1)], dim=1)
    mask = m.gt_mask.expand(R, -1, -1).clone()
    arrays = dict(cfg=json.dumps(YAML), seed=23, keys=json.dumps({k: list(v) for k, v in shapes.items()}),
                  target=target)
    for t in (0, 50, 99):
        x = torch.randn(R, 200, 1) * 0.8
        tc = torch.full((R,), t, dtype=torch.long)
        with torch.no_grad():
            trend, season = net.model(x, tc)
        arrays.update({"fwd%d:x" % t: x, "fwd%d:trend" % t: trend, "fwd%d:season" % t: season})
    kw = {"coef": 0.1, "learning_rate": 0.05}
    for (time, time_next) in ((99, 98), (80, 79), (50, 49), (3, 2)):
        img = torch.randn(R, 200, 1)
        tc = torch.full((R,), time, dtype=torch.long)
        key = "step%d:" % time
        arrays[key + "img_in"] = img.clone()
        with torch.no_grad():
            pred_noise, x_start = net.model_predictions(img, tc, clip_x_start=True)
            alpha, alpha_next = net.alphas_cumprod[time], net.alphas_cumprod[time_next]
            sigma = net.eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            pred_mean = x_start * alpha_next.sqrt() + c * pred_noise
            with Tape() as tape:
                noise = torch.randn_like(img)
                img1 = pred_mean + sigma * noise
                arrays[key + "ddim"] = img1.clone()
                # gradient of the refinement loss at the DDIM point, through the reference network
                p = torch.nn.Parameter(img1.clone())
                with torch.enable_grad():
                    xs = net.output(x=p, t=tc)
                    loss = 0.1 * ((pred_mean - p) ** 2 / 1.).mean(dim=0).sum() + \
                        ((xs[mask] - target[mask]) ** 2).mean(dim=0).sum()
                    loss.backward()
                arrays[key + "grad"] = p.grad.clone()
                img2 = net.langevin_fn(sample=img1, mean=pred_mean, sigma=sigma, t=tc, tgt_embs=target,
                                       partial_mask=mask, **kw)
                arrays[key + "langevin"] = img2.clone()
                target_t = net.q_sample(target, t=tc)
                img2[mask] = target_t[mask]
        arrays.update({key + "x_start": x_start, key + "pred_noise": pred_noise, key + "img_out": img2,
                       key + "n_draws": len(tape.draws)})
        arrays.update({key + "z%d" % i: z for i, z in enumerate(tape.draws)})
    save("dts_yaml_steps.npz", **arrays)


if __name__ == "__main__":
    main()
