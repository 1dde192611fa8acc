"""NsDiff model objects with the reference's attribute surface, backed by the CUDA hot path.

Mirrors models/Diffusion_model/NsDiff/NsDiff_model.py: classes ``NsDiff_model`` (:16-268) and
``NsDiff_model_variants`` (:271-495) keep their constructor arguments, parameter names (so the
reference's checkpoints load with ``strict=True``), ``scaler_*`` helpers, ``cond_pred_model_g`` and
``evaluation_step``.  The nn.Modules here are parameter containers: the arithmetic runs in the
kernels of csrc/ through the C ABI, and there is no CPU path.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib, kernels, schedules

EPS = 10e-8  # NsDiff_model.py:37
FX_ROWS_PER_CALL = 4096


class ConditionalLinearParams(nn.Module):
    """Parameters of ConditionalLinear (denoise.py:6-20): ``lin`` + per-step ``embed`` table."""

    def __init__(self, num_in, num_out, n_steps):
        super().__init__()
        self.lin = nn.Linear(num_in, num_out)
        self.embed = nn.Embedding(n_steps, num_out)
        self.embed.weight.data.uniform_()


class GuidedDenoiserParams(nn.Module):
    """Parameters of ConditionalGuidedModel (denoise.py:23-33)."""

    def __init__(self, diff_steps, enc_in):
        super().__init__()
        self.lin1 = ConditionalLinearParams(enc_in * 3, 128, diff_steps)
        self.lin2 = ConditionalLinearParams(128, 128, diff_steps)
        self.lin3 = ConditionalLinearParams(128, 128, diff_steps)
        self.lin4 = nn.Linear(128, enc_in)
        self.sigma_lin = nn.Linear(128, enc_in)


class NsDiffNet(nn.Module):
    """NsDiff_net (NsDiff_net.py:77-172): the denoiser under the reference's attribute name
    ``diffussion_model`` plus the schedule tables as plain attributes (not buffers)."""

    def __init__(self, configs, device):
        super().__init__()
        self.args = configs
        self.device = device
        self.num_timesteps = configs.diffusion_steps
        tab = schedules.nsdiff_tables(configs.diffusion_schedule, configs.diffusion_steps,
                                      configs.beta_start, configs.beta_end)
        self.tables = tab
        for k, v in tab.items():
            setattr(self, k, v.to(device))
        self.alphas_tilde = self.alphas_cumprod_sum
        self.diffussion_model = GuidedDenoiserParams(configs.diffusion_steps, configs.dataset_nf)


class SigmaEstimation(nn.Module):
    """g(x) (g_backbone.py:19-72): same ``mlp.{0,2,3,5,6}`` parameters; forward = one CUDA kernel."""

    def __init__(self, seq_len, pred_len, enc_in, hidden_size=512, kernel_size=24):
        super().__init__()
        self.pred_len, self.seq_len, self.enc_in = pred_len, seq_len, enc_in
        self.hidden_size, self.kernel_size = hidden_size, kernel_size
        self.mlp = nn.Sequential(
            nn.Linear(seq_len - kernel_size, hidden_size), nn.ReLU(), nn.LayerNorm([enc_in, hidden_size]),
            nn.Linear(hidden_size, hidden_size), nn.ReLU(), nn.LayerNorm([enc_in, hidden_size]),
            nn.Linear(hidden_size, pred_len))

    def _weights(self):
        m = self.mlp
        return [t.detach().contiguous() for t in (m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[3].weight,
                                                  m[3].bias, m[5].weight, m[5].bias, m[6].weight, m[6].bias)]

    def forward(self, x_enc, add_eps=0.0):
        if x_enc.dim() != 3:
            raise ValueError("x_enc must be a 3D tensor with shape (B, T, N)")
        if self.kernel_size < 1 or self.kernel_size > x_enc.shape[1]:
            raise ValueError("window_size must be between 1 and T (got window_size={}, T={})".format(
                self.kernel_size, x_enc.shape[1]))
        x = x_enc.to(self.mlp[0].weight.device, torch.float32).contiguous()
        return kernels.sigma_estimation(self._weights(), x, self.kernel_size, self.pred_len, add_eps)


class _NsDiffBase(nn.Module):
    """Shared by NsDiff_model and NsDiff_model_variants."""

    variant_adds_eps = False

    def _init_common(self, net_param):
        self.device = net_param["device"]
        self.dataset_nf = net_param["dataset_nf"]
        self.windows = net_param["windows"]
        self.pred_len = net_param["pred_len"]
        self.rolling_length = net_param["rolling_length"]
        self.seq_len = net_param["seq_len"] = self.windows
        self.label_len = net_param["label_len"] = self.windows // 2
        self.diffusion_steps = net_param["diffusion_steps"]
        self.EPS = EPS
        self.configs = SimpleNamespace(**net_param)
        self.scaler = net_param["scaler_type"]
        self.register_buffer("scaler_mean", torch.zeros(self.dataset_nf))
        self.register_buffer("scaler_std", torch.zeros(self.dataset_nf))
        self.sampler_impl = kernels.IMPL_TCGEN05      # the library picks the kernel for (kind, F, T), include/upd_b200.h
        self._packed = None
        self._packed_key = None
        self._windows_drawn = 0

    # ---- reference helpers (NsDiff_model.py:99-113) ----
    def scaler_fit(self, data):
        data_std = data.std(axis=0)
        data_std[data_std == 0] = 1
        self.scaler_mean = data.mean(axis=0)
        self.scaler_std = data_std

    def scaler_transform(self, data):
        return (data - self.scaler_mean) / self.scaler_std

    def scaler_inverse_transform(self, data):
        return (data * self.scaler_std) + self.scaler_mean

    # ---- device-side state ----
    def packed_weights(self):
        """Packed denoiser blob on the model's device; rebuilt if the parameters changed."""
        w = self.model.diffussion_model.lin1.lin.weight
        key = (w.device,) + tuple(p._version for p in self.model.diffussion_model.parameters())
        if self._packed is None or self._packed_key != key:
            sd = {"model." + k: v for k, v in self.model.state_dict().items()}
            rows = schedules.stack_rows(self.model.tables, schedules.NSDIFF_ROWS)
            self._packed = kernels.pack_denoiser(sd, kernels.KIND_NSDIFF, self.dataset_nf, self.diffusion_steps, rows,
                                                 w.device)
            self._packed_key = key
        return self._packed

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    @_lib.on_device
    def condition(self, batch_x):
        """f(x) and g(x) once per window row: -> (y_0_hat [R,O,F] or None, gx [R,O,F])."""
        dev = self.model.diffussion_model.lin1.lin.weight.device
        batch_x = batch_x.to(dev, torch.float32).contiguous()
        y0 = None
        if getattr(self, "cond_pred_model", None) is not None:
            parts = []
            for r0 in range(0, batch_x.size(0), FX_ROWS_PER_CALL):      # bounds the encoder's activation memory
                xb = batch_x[r0:r0 + FX_ROWS_PER_CALL]
                dec_inp = torch.cat([xb[:, -self.label_len:, :],
                                     torch.zeros(xb.size(0), self.pred_len, self.dataset_nf, device=dev)], dim=1)
                parts.append(self.cond_pred_model(xb, dec_inp)[0])
            y0 = torch.cat(parts).contiguous()
        if getattr(self, "cond_pred_model_g", None) is not None:
            gx = self.cond_pred_model_g(batch_x, add_eps=EPS if self.variant_adds_eps else 0.0)
        else:
            gx = torch.ones(batch_x.shape[0], self.pred_len, self.dataset_nf, device=dev)
        return y0, gx

    @_lib.on_device
    def sample_windows(self, windows, noise=None, seed=None, window_base=None):
        """Batched hot path: ``windows`` [W, B, L(+O), F] already scaled -> trajectories
        [W*B, K, O, F] on the device (K = (n_z_samples // parallel_sample) * parallel_sample).
        ``noise`` (validation mode) is laid out [W, K/S, T, B*S, O, F] as the reference draws it."""
        W, B = windows.shape[0], windows.shape[1]
        S = int(self.configs.parallel_sample)
        K = (int(self.configs.n_z_samples) // S) * S
        if K <= 0:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors (n_z_samples // parallel_sample == 0: the reference's chunk loop is empty, NsDiff_model.py:227-247)")
        x = windows.reshape(W * B, windows.shape[2], windows.shape[3])[:, :self.windows, :]
        with torch.no_grad():
            y0, gx = self.condition(x)
            if seed is None:
                seed = torch.initial_seed()
            if window_base is None:
                window_base = self._windows_drawn
                self._windows_drawn += W
            return kernels.nsdiff_sample(self.packed_weights(), y0, gx, W, B, K, S, self.pred_len, self.dataset_nf,
                                         self.diffusion_steps, seed=seed, window_base=window_base,
                                         noise=None if noise is None else noise.to(gx.device, torch.float32).contiguous(),
                                         impl=self.sampler_impl)

    def evaluation_step(self, batch, noise=None):
        """NsDiff_model.py:180-268 / :404-495: batch [B, L or L+O, F] (scaled) ->
        (outs [B,O,F,K] on the CPU, a permuted view of contiguous [B,K,O,F]; batch_y or None)."""
        if batch.shape[1] - self.windows >= self.pred_len:
            batch_y = batch[:, self.windows:, :].to(self.device)
            assert batch_y.size(1) == self.pred_len, "pred_len is not equal to the length of the prediction"
        else:
            batch_y = None
        if noise is not None:
            noise = noise.unsqueeze(0)
        traj = self.sample_windows(batch.unsqueeze(0), noise=noise)
        preds = traj.cpu()
        outs = preds.permute(0, 2, 3, 1)
        assert (outs.shape[1], outs.shape[2]) == (self.pred_len, self.dataset_nf)
        return outs, batch_y

    def training_step(self, batch):
        raise NotImplementedError("training is outside the accelerated hot path (SURVEY section 8: out of scope)")


class NsDiff_model(_NsDiffBase):
    """NsDiff_model.py:16-97.  ``train_model_select`` in {'NsDiff_model','pretrain_f','pretrain_g'}."""

    variant_adds_eps = False   # the base class feeds g(x) without EPS at inference (:223)

    def __init__(self, net_param, train_model_select, pretrain_f_path="results/pre_model_F",
                 pretrain_g_path="results/pre_model_G"):
        super().__init__()
        self._init_common(net_param)
        self.load_pretrain = net_param["load_pretrain"]
        self.freeze_pretrain = net_param["freeze_pretrain"] if "freeze_pretrain" in net_param else False
        from .fx_encoder import NsTransformer
        if train_model_select == "NsDiff_model":
            self.model = NsDiffNet(self.configs, self.device)
            self.cond_pred_model = NsTransformer(self.configs)
            rolling = self.rolling_length
            if self.load_pretrain:
                pre = torch.load(pretrain_g_path + "/model_trained", map_location="cpu", weights_only=True)
                rolling = pre["net_param"]["rolling_length"]
            self.cond_pred_model_g = SigmaEstimation(self.windows, self.pred_len, self.dataset_nf, 512, rolling)
            if self.load_pretrain:
                sd = {k.replace("cond_pred_model_g.", ""): v for k, v in pre["state_dict"].items()}
                sd.pop("scaler_mean")
                sd.pop("scaler_std")
                self.cond_pred_model_g.load_state_dict(sd, strict=True)
        elif train_model_select == "pretrain_f":
            self.cond_pred_model = NsTransformer(self.configs)
        elif train_model_select == "pretrain_g":
            self.cond_pred_model_g = SigmaEstimation(self.windows, self.pred_len, self.dataset_nf, 512,
                                                     self.rolling_length)
        else:
            raise ValueError("train_model_select should be in ['NsDiff_model', 'pretrain_f', 'pretrain_g']")
        self.to(self.device)


class NsDiff_model_variants(_NsDiffBase):
    """NsDiff_model.py:271-328.  ``train_model_select`` in Guassian / cond_mean / cond_var / wo_UANS."""

    variant_adds_eps = True    # :450 adds EPS to g(x)

    def __init__(self, net_param, train_model_select):
        super().__init__()
        self._init_common(net_param)
        self.train_model_select = train_model_select
        if train_model_select not in ("Guassian", "cond_mean", "cond_var", "wo_UANS"):
            raise ValueError("train_model_select should be in Guassian/cond_mean/cond_var")
        self.model = NsDiffNet(self.configs, self.device)
        self.cond_pred_model = None
        self.cond_pred_model_g = None
        if train_model_select in ("cond_mean", "wo_UANS"):
            from .fx_encoder import NsTransformer
            self.cond_pred_model = NsTransformer(self.configs)
        if train_model_select in ("cond_var", "wo_UANS"):
            self.cond_pred_model_g = SigmaEstimation(self.windows, self.pred_len, self.dataset_nf, 512,
                                                     self.rolling_length)
        self.to(self.device)
