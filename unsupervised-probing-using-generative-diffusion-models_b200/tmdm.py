"""TMDM model object with the reference's surface (models/Diffusion_model/TMDM/tmdm_adapter.py:24-155),
backed by the fused CUDA sampler (kind TMDM).  Parameter containers only; no CPU path."""
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib, kernels, schedules
from .fx_encoder import DataEmbedding, NsTransformer
from .nsdiff import ConditionalLinearParams


class TmdmDenoiserParams(nn.Module):
    """ConditionalGuidedModel of TMDM (tmdm_model.py:23-37): cat(y_t, y_0_hat) -> 3 x 128 -> F;
    the step-embedding tables have timesteps + 1 rows (:26)."""

    def __init__(self, timesteps, enc_in):
        super().__init__()
        n_steps = timesteps + 1
        self.lin1 = ConditionalLinearParams(enc_in * 2, 128, n_steps)
        self.lin2 = ConditionalLinearParams(128, 128, n_steps)
        self.lin3 = ConditionalLinearParams(128, 128, n_steps)
        self.lin4 = nn.Linear(128, enc_in)


class TmdmNet(nn.Module):
    """TMDM (TMDM.py:22-98).  ``enc_embedding`` is kept for state-dict shape only: with cat_x=cat_y_pred=True
    (tmdm.yml:21-22) the denoiser ignores the embedded x, so sampling never evaluates it."""

    def __init__(self, configs, device):
        super().__init__()
        self.args, self.device = configs, device
        self.num_timesteps = configs.timesteps
        tab = schedules.tmdm_tables(getattr(configs, "beta_schedule", "linear"), configs.timesteps,
                                    getattr(configs, "beta_start", 1e-4), getattr(configs, "beta_end", 0.02))
        self.tables = tab
        for k, v in tab.items():
            setattr(self, k, v.to(device))
        self.diffussion_model = TmdmDenoiserParams(configs.timesteps, configs.enc_in)
        self.enc_embedding = DataEmbedding(configs.enc_in, configs.CART_input_x_embed_dim)


class TMDM_model(nn.Module):
    def __init__(self, net_param):
        super().__init__()
        self.device = net_param["device"]
        self.dataset_nf = net_param["dataset_nf"]
        self.windows = net_param["windows"]
        self.pred_len = net_param["pred_len"]
        self.seq_len = net_param["seq_len"] = self.windows
        self.label_len = net_param["label_len"] = net_param.get("label_len", self.windows // 2)
        self.diffusion_steps = net_param.get("diffusion_steps", 100)
        self.n_z_samples = net_param.get("n_z_samples", 100)
        self.parallel_sample = net_param.get("parallel_sample", min(10, self.n_z_samples))
        self.scaler = net_param.get("scaler_type", None)
        self.k_z = net_param.get("k_z", 0.01)
        for key, val in (("enc_in", self.dataset_nf), ("dec_in", self.dataset_nf), ("c_out", self.dataset_nf),
                         ("features", "M"), ("embed", "fixed"), ("freq", "h"), ("dropout", 0.05),
                         ("output_attention", False), ("d_model", 64)):
            net_param.setdefault(key, val)
        net_param.setdefault("CART_input_x_embed_dim", net_param["d_model"])
        for key, val in (("factor", 3), ("n_heads", 4), ("d_ff", 128), ("activation", "gelu"), ("e_layers", 2),
                         ("d_layers", 1), ("p_hidden_dims", [64, 64]), ("p_hidden_layers", 2)):
            net_param.setdefault(key, val)
        net_param.setdefault("d_z", net_param["d_model"])
        net_param.setdefault("k_cond", 1.0)
        net_param.setdefault("timesteps", self.diffusion_steps)
        net_param.setdefault("diffusion_config_dir", os.path.join(os.path.dirname(__file__), "tmdm.yml"))
        self.configs = SimpleNamespace(**net_param)
        self.register_buffer("scaler_mean", torch.zeros(self.dataset_nf))
        self.register_buffer("scaler_std", torch.ones(self.dataset_nf))
        self.model = TmdmNet(self.configs, self.device)
        self.cond_pred_model = NsTransformer(self.configs, vae=True)
        self.sampler_impl = kernels.IMPL_TCGEN05      # the library picks the kernel for (kind, F, T), include/upd_b200.h
        self._packed = None
        self._packed_key = None
        self._windows_drawn = 0
        self.to(self.device)

    def scaler_fit(self, data):
        data_std = data.std(axis=0)
        data_std[data_std == 0] = 1
        self.scaler_mean = data.mean(axis=0)
        self.scaler_std = data_std

    def scaler_transform(self, data):
        return (data - self.scaler_mean) / self.scaler_std

    def scaler_inverse_transform(self, data):
        return (data * self.scaler_std) + self.scaler_mean

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def packed_weights(self):
        w = self.model.diffussion_model.lin1.lin.weight
        key = (w.device,) + tuple(p._version for p in self.model.diffussion_model.parameters())
        if self._packed is None or self._packed_key != key:
            sd = {"model." + k: v for k, v in self.model.state_dict().items()}
            rows = schedules.stack_rows(self.model.tables, schedules.TMDM_ROWS)
            self._packed = kernels.pack_denoiser(sd, kernels.KIND_TMDM, self.dataset_nf, self.diffusion_steps, rows,
                                                 w.device)
            self._packed_key = key
        return self._packed

    @_lib.on_device
    def condition(self, batch_x):
        """Condition mean over label_len + pred_len positions (tmdm_adapter.py:123-124)."""
        dev = self.model.diffussion_model.lin1.lin.weight.device
        batch_x = batch_x.to(dev, torch.float32).contiguous()
        dec_inp = torch.cat([batch_x[:, -self.label_len:, :],
                             torch.zeros(batch_x.size(0), self.pred_len, self.dataset_nf, device=dev)], dim=1)
        _, y0, _, _ = self.cond_pred_model(batch_x, None, dec_inp, None)
        return y0.contiguous()

    @_lib.on_device
    def sample_windows(self, windows, noise=None, seed=None, window_base=None, y_0_hat=None):
        """windows [W,B,L(+O),F] scaled -> trajectories [W*B, K, pred_len, F] on the device."""
        W, B = windows.shape[0], windows.shape[1]
        S = min(int(self.parallel_sample), int(self.n_z_samples))
        if self.n_z_samples % S != 0:
            raise ValueError("n_z_samples must be divisible by parallel_sample")
        K = int(self.n_z_samples)
        x = windows.reshape(W * B, windows.shape[2], windows.shape[3])[:, :self.windows, :]
        Lr = self.label_len + self.pred_len
        with torch.no_grad():
            y0 = self.condition(x) if y_0_hat is None else y_0_hat.to(self.scaler_mean.device, torch.float32).contiguous()
            if seed is None:
                seed = torch.initial_seed()
            if window_base is None:
                window_base = self._windows_drawn
                self._windows_drawn += W
            traj = kernels.tmdm_sample(self.packed_weights(), y0, W, B, K, S, Lr, self.dataset_nf, self.diffusion_steps,
                                       seed=seed, window_base=window_base,
                                       noise=None if noise is None else noise.to(y0.device, torch.float32).contiguous(),
                                       impl=self.sampler_impl)
        return traj[:, :, -self.pred_len:, :].contiguous()

    def evaluation_step(self, batch, noise=None, y_0_hat=None):
        """tmdm_adapter.py:116-155 -> (outs [B,O,F,K] cpu, batch_y or None)."""
        if batch.shape[1] - self.windows >= self.pred_len:
            batch_y = batch[:, self.windows:self.windows + self.pred_len, :].to(self.device)
        else:
            batch_y = None
        traj = self.sample_windows(batch.unsqueeze(0), noise=None if noise is None else noise.unsqueeze(0), y_0_hat=y_0_hat)
        return traj.cpu().permute(0, 2, 3, 1), batch_y

    def training_step(self, batch):
        raise NotImplementedError("training is outside the accelerated hot path (SURVEY section 8: out of scope)")
