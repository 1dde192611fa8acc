"""Host-side orchestration of the graph samplers (NsDiff_spatial, DiffSTG) on the CPU: the C-ABI kernels are replaced by oracle-based stand-ins INSIDE THIS
TEST (monkeypatched, restored afterwards), so tile order, chunk / window batching, the shared-CSR replica convention, draw
order and Philox row keys are checked against the reference-made fixture without a GPU.  The kernels themselves are
checked on the GPU (tests/test_gpu_nsx.py, tests/test_gpu_stg.py); the product has no CPU path."""
import contextlib
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import diffusionts_oracle as dto, diffstg_oracle as so, nsdiff_oracle as nso, nsdiff_spatial_oracle as nsx
from oracle import sigma_oracle

torch.set_num_threads(1)


class _StandInLib:
    """torch-CPU stand-ins with the C-ABI argument order (tensors instead of pointers)."""

    def upd_stg_gated_aggregate(self, kqvs, rowptr, col, bias, N, V, C, relu, out, st):
        k, q, v, s = kqvs.split(C, dim=1)
        rp = rowptr.long()
        dst = torch.repeat_interleave(torch.arange(V), rp[1:] - rp[:-1])
        par = so.duplicate_edge_index(N // V, torch.stack([col.long()[:dst.numel()], dst]), V)
        o = torch.zeros(N, C).index_add_(0, par[1], torch.sigmoid(k[par[1]] + q[par[0]]) * v[par[0]]) + s
        if bias is not None:
            o = o + bias
        out.copy_(torch.relu(o) if relu else o)
        return 0

    def upd_nsx_step(self, e, w4, b4, ws, bs, y, yT, gx, z, sched, steps, t, N, DH, T, nf, out, eo, so_, st):
        eps = F.linear(e.transpose(1, 2), w4, b4)
        sig = F.softplus(F.linear(F.softplus(e.transpose(1, 2)), ws, bs))
        if eo is not None:
            eo.copy_(eps)
        if so_ is not None:
            so_.copy_(sig)
        if y is not None:
            assert (z is None) == (t == 0)
            sc = nso.nsdiff_schedule("linear", steps, 1e-4, 0.02)
            sy0, noise = nso._sigma_y0_and_noise(sc, t, gx, sig)
            y0 = nso._y0_reparam(sc, t, y, yT, eps, noise)
            if t > 0:
                g0, g1, g2 = nso._gammas(sc, t, gx, sy0)
                y0 = g0 * y0 + g1 * y + g2 * yT + torch.sqrt(sig) * z
            out.copy_(y0)
        return 0

    def upd_stg_conv1d(self, x, w, b, N, CI, CO, Tin, k, stride, pad, transposed, y, st):
        fn = F.conv_transpose1d if transposed else F.conv1d
        y.copy_(fn(x, w, b, stride=stride, padding=pad))
        return 0

    def upd_stg_posterior(self, xt, pred, z, n, a, b, c, out, st):
        a, b, c = (torch.tensor(v, dtype=torch.float32) for v in (a, b, c))
        out.copy_(a * (xt - b * pred) + c * (z if z is not None else pred))
        return 0

    def upd_gauss_fill(self, z, rows, elems, seed, base, draw, st):      # keyed per (seed, row_base + row, draw) like the kernel
        for r in range(rows):
            g = torch.Generator().manual_seed((seed * 1315423911 + (base + r) * 2654435761 + draw) % (2 ** 63))
            z[r].copy_(torch.randn(z[r].shape, generator=g))
        return 0


def _front(self, b, x, t, c_in, c_out, T_in, as_operand):
    if isinstance(x, tuple):
        x = torch.cat(x, dim=1)
    N = x.shape[0]
    h = F.conv1d(F.pad(x, (2, 0)), b["tcn1.w"], b["tcn1.b_step"][t])
    h = F.conv1d(F.pad(h, (2, 0)), b["tcn2.w"], b["tcn2.b"])
    var, mu = torch.var_mean(h, dim=1, unbiased=False, keepdim=True)
    hn = ((h - mu) * torch.rsqrt(var + 1e-5) * b["norm_w"][None, :, None] + b["norm_b"][None, :, None]).reshape(N, -1)
    sc = None if c_in == c_out else torch.matmul(b["sc_w2"], x).reshape(N, -1)
    return hn, sc


@pytest.fixture()
def cpu_stand_ins(monkeypatch):
    from updgm_b200 import _lib, diffstg, nsdiff_spatial as ns
    monkeypatch.setattr(_lib, "lib", lambda: _StandInLib())
    monkeypatch.setattr(_lib, "ptr", lambda t: t)
    monkeypatch.setattr(_lib, "stream_ptr", lambda d: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda d: torch.device("cpu"))
    monkeypatch.setattr(_lib, "check", lambda rc, name: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(diffstg.PreparedUGnet, "_front", _front)
    orig_init = diffstg.PreparedUGnet.__init__

    def init(self, *a, **k):
        orig_init(self, *a, **k)
        for b in self.blocks.values():
            if "down_w3" in b:
                b["down_w3"] = b["kqvs_w3"] = b["up_w3"] = None          # fp32 library GEMMs instead of the fp16 split path
    monkeypatch.setattr(diffstg.PreparedUGnet, "__init__", init)
    monkeypatch.setattr(ns.SigmaEstimation, "forward", lambda self, x, add_eps=0.0: sigma_oracle.sigma_estimation(
        {"cond_pred_model_g." + k: v for k, v in self.state_dict().items()}, x, self.kernel_size, self.pred_len) + add_eps)
    return ns, diffstg


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().pow(2).mean().sqrt())


def test_spatial_sampler_orchestration_matches_reference_fixture(cpu_stand_ins):
    from updgm_b200.fx_encoder import PositionalEmbedding
    ns, diffstg = cpu_stand_ins
    g = np.load("{}/nsx_small_evalstep.npz".format(GOLDEN))
    cfg, shapes, seed = json.loads(str(g["cfg"])), json.loads(str(g["keys"])), int(g["seed"])
    sd = dto.synth_state_dict(shapes, seed)
    for k in shapes:
        if ".net.0." in k:
            sd[k] = sd[k.replace(".net.0.", ".conv.")]
    fx_sd = {nsx.FX + k: v for k, v in dto.synth_state_dict(json.loads(str(g["fx_keys"])), seed + 1).items()}
    for k in fx_sd:
        if k.endswith("position_embedding.pe"):
            fx_sd[k] = PositionalEmbedding(cfg["d_model"]).pe.clone()
    sd.update(fx_sd)
    sd["scaler_mean"], sd["scaler_std"] = torch.zeros(cfg["dataset_nf"]), torch.ones(cfg["dataset_nf"])
    m = ns.NsDiff_model_spatial(dict(cfg, device="cpu"), "NsDiff_model").eval()
    m.load_state_dict(sd, strict=True)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V, T, S = x.shape[0], cfg["diffusion_steps"], cfg["parallel_sample"]
    # the denoiser on S replicas in the duplicated-edge layout
    for t in (0, 1, T - 1):
        y, y0, gx = [torch.from_numpy(g["den%d:%s" % (t, n)]) for n in ("y", "y0", "gx")]
        eps, sig = m.denoise(y, y0, gx, t, ei, V)
        assert _rel(eps, torch.from_numpy(g["den%d:eps" % t])) < 1e-5 and _rel(sig, torch.from_numpy(g["den%d:sig" % t])) < 1e-5
    # whole evaluation_step with the reference's recorded draws (f(x) and g(x) included)
    draws = [torch.from_numpy(g["z%03d" % i]) for i in range(int(g["n_draws"]))]
    noise = [draws[c * T:(c + 1) * T] for c in range(len(draws) // T)]
    outs, by = m.evaluation_step(diffstg.GraphData(x=x, edge_index=ei, num_nodes=V), noise=noise)
    ref = torch.from_numpy(g["outs"])
    assert by is None and tuple(outs.shape) == tuple(ref.shape) and _rel(outs, ref) < 1e-5
    # Philox mode: results do not depend on how chunks / windows are cut into launches, and a later window sampled alone
    # reproduces its rows (row keys are global)
    wins = torch.stack([x, x.flip(0), x * 0.5], 0)
    a = m.sample_windows(wins, ei, V, seed=5, window_base=7)
    m.rows_per_launch = V * S
    b = m.sample_windows(wins, ei, V, seed=5, window_base=7)
    c = m.sample_windows(wins[1:], ei, V, seed=5, window_base=8)
    assert _rel(b, a) < 1e-5 and _rel(c, a[V:]) < 1e-5 and float(a.var(dim=1).mean()) > 0
    # the counter of windows drawn advances when no base is given
    m._windows_drawn = 0
    m.sample_windows(wins[:2], ei, V, seed=5)
    assert m._windows_drawn == 2


def test_diffstg_sampler_orchestration_matches_reference_fixture(cpu_stand_ins):
    """DiffSTG.sample_windows: (window, round, parallel sample) -> graph replicas, step plan, draw order, output layout."""
    ns, diffstg = cpu_stand_ins
    g = np.load("{}/stg_small_evalstep.npz".format(GOLDEN))
    cfg, shapes, seed = json.loads(str(g["cfg"])), json.loads(str(g["keys"])), int(g["seed"])
    sd = dto.synth_state_dict(shapes, seed)
    for k in shapes:
        if ".net.0." in k:
            sd[k] = sd[k.replace(".net.0.", ".conv.")]
    sd["scaler_mean"], sd["scaler_std"] = torch.zeros(cfg["F"]), torch.ones(cfg["F"])
    m = diffstg.DiffSTG(dict(cfg, device="cpu")).eval()
    m.load_state_dict(sd, strict=True)
    x, ei = torch.from_numpy(g["x"]), torch.from_numpy(g["edge_index"])
    V = x.shape[0]
    draws = [torch.from_numpy(g["z%03d" % i]) for i in range(int(g["n_draws"]))]
    per = m.draws_per_round()
    noise = [draws[r * per:(r + 1) * per] for r in range(cfg["sequential_sampling"])]
    outs, truth = m.evaluation_step(diffstg.GraphData(x=x, edge_index=ei, num_nodes=V), noise=noise)
    ref = torch.from_numpy(g["outs"])
    assert truth is None and tuple(outs.shape) == tuple(ref.shape) and _rel(outs, ref) < 1e-5
    wins = torch.stack([x, x.flip(0), x * 0.5], 0)
    a = m.sample_windows(wins, ei, V, seed=5, window_base=7)
    m.rows_per_launch = V
    b = m.sample_windows(wins, ei, V, seed=5, window_base=7)
    c = m.sample_windows(wins[1:], ei, V, seed=5, window_base=8)
    assert _rel(b, a) < 1e-5 and _rel(c, a[V:]) < 1e-5 and float(a.var(dim=1).mean()) > 0


@pytest.mark.parametrize("d,L,fT,B", [(6, 20, 4, 3), (4, 16, 10, 2), (8, 9, 2, 1)])
def test_dense_time_conv_maps_equal_the_bridge_convolutions(d, L, fT, B):
    """The NsDiff_spatial f(x) bridge runs its (1, T+1) Conv2d / ConvTranspose2d (mu_backbone.py:203-206) as dense
    matrices on upd_gemm3; the re-indexing itself is host logic and is checked here in float64 against
    F.conv1d / F.conv_transpose1d (bit-level agreement is not expected: the summation order differs)."""
    import torch.nn.functional as F
    from updgm_b200.nsdiff_spatial import dense_time_conv_maps
    torch.manual_seed(d * 100 + L)
    down = torch.nn.Conv2d(d, d, (1, L + 1), (1, 1), (0, fT // 2)).double()
    up = torch.nn.ConvTranspose2d(d, d, (1, L + 1), (1, 1), (0, fT // 2)).double()
    x = torch.randn(B, L, d, dtype=torch.float64)
    Dn, Up = dense_time_conv_maps(down.weight.detach()[:, :, 0, :], up.weight.detach()[:, :, 0, :], L, fT)
    assert tuple(Dn.shape) == (fT * d, L * d) and tuple(Up.shape) == (L * d, fT * d)
    h = F.conv1d(x.transpose(1, 2), down.weight[:, :, 0, :], down.bias, padding=fT // 2)
    assert h.shape[-1] == fT
    s_ref = h.transpose(1, 2).reshape(B, fT * d)
    s = x.reshape(B, L * d) @ Dn.t() + down.bias.detach().repeat(fT)
    assert float((s - s_ref).abs().max()) < 1e-12
    o_ref = F.conv_transpose1d(s_ref.reshape(B, fT, d).transpose(1, 2), up.weight[:, :, 0, :], up.bias,
                               padding=fT // 2).transpose(1, 2)
    o = (s_ref @ Up.t() + up.bias.detach().repeat(L)).view(B, L, d)
    assert tuple(o_ref.shape) == (B, L, d) and float((o - o_ref).abs().max()) < 1e-12
