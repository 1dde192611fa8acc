"""GPU parity tests proper: the CUDA path, called through the C ABI, against the reference-made
golden fixtures and against the CPU oracle on seeded inputs.

Tolerances (stated per north star): trajectories rel 1e-3 per value (+ an absolute floor of
1e-4 x rms for values near zero), MPV rel 1e-4.  Observed errors are ~1e-6 x rms; the tighter
`TIGHT` bounds below are what the kernels are actually held to.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, load_wo_fx_checkpoint
from oracle import mpv_oracle, nsdiff_oracle, sigma_oracle, tmdm_oracle

pytestmark = pytest.mark.gpu

TIGHT_RMS = 2e-5      # max |d| / rms(ref)
MPV_RTOL = 1e-5


def _k():
    from updgm_b200 import kernels, schedules
    return kernels, schedules


def _dev():
    return torch.device("cuda:0")


def _assert_traj(out, ref, what, tight=None):
    tight = TIGHT_RMS if tight is None else tight
    out, ref = out.double().cpu(), ref.double().cpu()
    assert torch.isfinite(out).all(), what
    rms = ref.pow(2).mean().sqrt()
    d = (out - ref).abs()
    assert (d <= 1e-3 * ref.abs() + 1e-4 * rms).all(), "{}: beyond the stated tolerance, max|d|/rms={:.3e}".format(
        what, float(d.max() / rms))
    assert float(d.max() / rms) <= tight, "{}: max|d|/rms={:.3e}".format(what, float(d.max() / rms))
    return float(d.max() / rms)


def _pack_ns(sd, F, T=20, schedule="linear"):
    kernels, schedules = _k()
    tab = schedules.nsdiff_tables(schedule, T, 1e-4, 0.02)
    return kernels.pack_denoiser(sd, kernels.KIND_NSDIFF, F, T, schedules.stack_rows(tab, schedules.NSDIFF_ROWS), _dev())


IMPLS = [pytest.param(1, id="simt"), pytest.param(0, id="tcgen05"), pytest.param(2, id="tcgen05x2"), pytest.param(4, id="tcgen05ws")]


# ---------------------------------------------------------------- tcgen05 descriptor known-answer
def _selftest_umma(a, b, mode=0, flags=0):
    """D = A @ B^T through the tcgen05 path of the sampler (A via TMEM, B via shared memory): the tests-only library
    libupd_selftest.so (csrc/selftest_umma.cu, built by _build.build_selftest(); not part of the product ABI)."""
    import ctypes
    from updgm_b200 import _build
    L = ctypes.CDLL(_build.build_selftest())
    L.upd_selftest_umma.restype = ctypes.c_int
    L.upd_selftest_umma.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
    d = torch.zeros((128, 128), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = L.upd_selftest_umma(a.data_ptr(), b.data_ptr(), d.data_ptr(), a.shape[1], mode, flags,
                                 torch.cuda.current_stream(a.device).cuda_stream)
    assert rc == 0, rc
    return d


@pytest.mark.parametrize("mode,K", [(0, 128), (1, 8), (1, 16)])
def test_umma_selftest(mode, K):
    kernels, _ = _k()
    g = torch.Generator().manual_seed(5 + K)
    a = torch.randn(128, K, generator=g)
    b = torch.randn(128, K, generator=g)
    # distinct structure per row/column so a transposed or permuted operand cannot pass by accident
    a += torch.arange(128).float().view(-1, 1) * 0.01
    b += torch.arange(K).float().view(1, -1) * 0.02
    ref = a.double() @ b.double().t()
    d = _selftest_umma(a.to(_dev()), b.to(_dev()), mode=mode, flags=0).cpu().double()
    scale = (a.double().abs() @ b.double().abs().t())
    err = ((d - ref).abs() / scale).max().item()
    assert err < 4e-6, "tcgen05 3-pass contraction off: {:.3e}".format(err)


# ---------------------------------------------------------------- sampler vs reference fixtures
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", ["psample_loop_wo_fx.npz", "psample_loop_wo_fx_fx.npz"])
def test_nsdiff_loop_golden_wo_fx(impl, name):
    kernels, _ = _k()
    _, sd = load_wo_fx_checkpoint()
    g = load_golden(name)
    packed = _pack_ns(sd, 2)
    R, O, F = g["gx"].shape
    dev = _dev()
    y0 = g["y_0_hat"].to(dev)
    out = kernels.nsdiff_sample(packed, y0 if name.endswith("_fx.npz") else None, g["gx"].to(dev), n_win=1, B=R, K=1,
                                S=1, O=O, F=F, T=20, noise=g["noise"].to(dev).contiguous(), impl=impl)
    _assert_traj(out.reshape(R, O, F), g["seq"][-1], name)


@pytest.mark.parametrize("impl", IMPLS)
def test_nsdiff_loop_golden_random_F1(impl):
    kernels, _ = _k()
    g = load_golden("psample_loop_randF1.npz")
    packed = _pack_ns(g["sd"], 1)
    R, O, F = g["gx"].shape
    dev = _dev()
    out = kernels.nsdiff_sample(packed, g["y_0_hat"].to(dev), g["gx"].to(dev), 1, R, 1, 1, O, F, 20,
                                noise=g["noise"].to(dev).contiguous(), impl=impl)
    _assert_traj(out.reshape(R, O, F), g["seq"][-1], "randF1")


@pytest.mark.parametrize("impl", IMPLS)
def test_evaluation_step_golden_chunked_noise_layout(impl):
    """K=8 in chunks of S=4: noise tensor in the reference's own consumption order."""
    kernels, _ = _k()
    net_param, sd = load_wo_fx_checkpoint()
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    dev = _dev()
    packed = _pack_ns(sd, 2)
    gx = (g["gx"] + 1e-7).to(dev)                     # NsDiff_model.py:450
    out = kernels.nsdiff_sample(packed, None, gx, 1, 1, 8, 4, 200, 2, 20, noise=g["noise"].to(dev).contiguous(), impl=impl)
    outs = out.reshape(1, 8, 200, 2).permute(0, 2, 3, 1)
    _assert_traj(outs, g["outs"], "evalstep")


@pytest.mark.parametrize("impl", IMPLS)
def test_tmdm_loop_golden(impl):
    kernels, schedules = _k()
    g = load_golden("tmdm_loop_randF1.npz")
    tab = schedules.tmdm_tables("linear", 20, 1e-4, 0.02)
    dev = _dev()
    packed = kernels.pack_denoiser(g["sd"], kernels.KIND_TMDM, 1, 20, schedules.stack_rows(tab, schedules.TMDM_ROWS), dev)
    R, Lr, F = g["y_0_hat"].shape
    out = kernels.tmdm_sample(packed, g["y_0_hat"].to(dev), 1, R, 1, 1, Lr, F, 20, noise=g["noise"].to(dev).contiguous(),
                              impl=impl)
    _assert_traj(out.reshape(R, Lr, F), g["seq"][-1], "tmdm")


# ---------------------------------------------------------------- sampler vs oracle, seeded, multi-window
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("F", [1, 2, 3, 4])
def test_nsdiff_vs_oracle_multiwindow(impl, F):
    """n_win=3 windows x B=2 rows x K=6 (S=3) x O=37 (ragged tile tail), every feature count."""
    kernels, schedules = _k()
    torch.manual_seed(100 + F)
    T, n_win, B, K, S, O = 20, 3, 2, 6, 3, 37
    sd = _random_ns_weights(F, T)
    sched = nsdiff_oracle.nsdiff_schedule("linear", T, 1e-4, 0.02)
    y0 = torch.randn(n_win * B, O, F) * 0.5
    gx = torch.rand(n_win * B, O, F) * 0.3 + 0.02
    noise = torch.randn(n_win, K // S, T, B * S, O, F)
    ref = torch.empty(n_win * B, K, O, F)
    for w in range(n_win):
        for c in range(K // S):
            it = iter(noise[w, c])
            y0t = nsdiff_oracle.tile_rows(y0[w * B:(w + 1) * B], S)
            gxt = nsdiff_oracle.tile_rows(gx[w * B:(w + 1) * B], S)
            seq = nsdiff_oracle.p_sample_loop(sd, sched, y0t, gxt, y0t, T, lambda like: next(it))
            ref[w * B:(w + 1) * B, c * S:(c + 1) * S] = seq[-1].reshape(B, S, O, F)
    dev = _dev()
    packed = _pack_ns(sd, F, T)
    if impl == 4 and F > 2:
        # the warp-specialised kernel is built for F <= 2 only (shared memory of its head-sum exchange); the C ABI says so
        with pytest.raises(RuntimeError, match="unsupported"):
            kernels.nsdiff_sample(packed, y0.to(dev), gx.to(dev), n_win, B, K, S, O, F, T, noise=noise.to(dev), impl=impl)
        return
    out = kernels.nsdiff_sample(packed, y0.to(dev), gx.to(dev), n_win, B, K, S, O, F, T, noise=noise.to(dev), impl=impl)
    _assert_traj(out, ref, "F=%d" % F)


def _random_ns_weights(F, T):
    P = nsdiff_oracle.DENOISER_PREFIX
    sd = {}
    dims = {"lin1": 3 * F, "lin2": 128, "lin3": 128}
    for name, k in dims.items():
        sd[P + name + ".lin.weight"] = (torch.rand(128, k) * 2 - 1) / k ** 0.5
        sd[P + name + ".lin.bias"] = (torch.rand(128) * 2 - 1) / k ** 0.5
        sd[P + name + ".embed.weight"] = torch.rand(T, 128)
    for name in ("lin4", "sigma_lin"):
        sd[P + name + ".weight"] = (torch.rand(F, 128) * 2 - 1) / 128 ** 0.5
        sd[P + name + ".bias"] = (torch.rand(F) * 2 - 1) / 128 ** 0.5
    return sd

@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("F", [1, 2])
@pytest.mark.parametrize("gain", [6.0, 200.0])
def test_nsdiff_large_weights_exercise_guard_and_softplus_tails(impl, F, gain):
    """Hidden-layer weights and biases scaled by `gain`: base-2 pre-activations reach +-20 (gain 6) and +-100s (gain 200,
    where the bound upd_denoiser_pack derives from the row norms exceeds 120, so the ex2 overflow guard of layers 2-3
    stays in).  Both tails of the softplus -- the two-MUFU form, its guarded variant and the one-MUFU form with the
    polynomial lg2 (csrc/sampler_math.cuh) -- against the oracle's F.softplus (denoise.py:47-51).  The two-pass
    contractions perturb an activation by 2^-12 relative whatever the weight scale, so the error stays a fixed fraction
    of the trajectory rms; the bound is the north star's, with 2e-4 x rms as the regression line."""
    kernels, schedules = _k()
    torch.manual_seed(700 + F)
    T, n_win, B, K, S, O = 20, 2, 2, 4, 2, 70
    sd = _random_ns_weights(F, T)
    P = nsdiff_oracle.DENOISER_PREFIX
    for name in ("lin1", "lin2", "lin3"):
        sd[P + name + ".lin.weight"] = sd[P + name + ".lin.weight"] * gain
        sd[P + name + ".lin.bias"] = sd[P + name + ".lin.bias"] * gain
    sched = nsdiff_oracle.nsdiff_schedule("linear", T, 1e-4, 0.02)
    y0 = torch.randn(n_win * B, O, F) * 0.5
    gx = torch.rand(n_win * B, O, F) * 0.3 + 0.02
    noise = torch.randn(n_win, K // S, T, B * S, O, F)
    ref = torch.empty(n_win * B, K, O, F)
    for w in range(n_win):
        for c in range(K // S):
            it = iter(noise[w, c])
            y0t = nsdiff_oracle.tile_rows(y0[w * B:(w + 1) * B], S)
            gxt = nsdiff_oracle.tile_rows(gx[w * B:(w + 1) * B], S)
            seq = nsdiff_oracle.p_sample_loop(sd, sched, y0t, gxt, y0t, T, lambda like: next(it))
            ref[w * B:(w + 1) * B, c * S:(c + 1) * S] = seq[-1].reshape(B, S, O, F)
    dev = _dev()
    packed = _pack_ns(sd, F, T)
    out = kernels.nsdiff_sample(packed, y0.to(dev), gx.to(dev), n_win, B, K, S, O, F, T, noise=noise.to(dev), impl=impl)
    err = _assert_traj(out, ref, "gain=%g F=%d" % (gain, F), tight=2e-4)
    print("large weights, gain %g, F=%d, impl %d: max|d|/rms %.2e" % (gain, F, impl, err))


@pytest.mark.parametrize("tc_impl", [0, 2, 4])
def test_tc_matches_simt_bitwise_structure_large(tc_impl):
    """Full-size tile coverage: 5 windows x K=100 x O=200 (100k rows; 782 tiles = several rotations of the persistent
    grid for both tcgen05 orchestrations) -- every tensor-core implementation agrees with the FFMA kernel."""
    kernels, _ = _k()
    _, sd = load_wo_fx_checkpoint()
    dev = _dev()
    packed = _pack_ns(sd, 2)
    torch.manual_seed(3)
    gx = (torch.rand(5, 200, 2) * 0.06 + 0.01).to(dev)
    a = kernels.nsdiff_sample(packed, None, gx, 5, 1, 100, 100, 200, 2, 20, seed=11, window_base=7, impl=tc_impl)
    b = kernels.nsdiff_sample(packed, None, gx, 5, 1, 100, 100, 200, 2, 20, seed=11, window_base=7, impl=1)
    _assert_traj(a, b, "tc vs simt")


# ---------------------------------------------------------------- Philox mode properties
@pytest.mark.parametrize("impl", IMPLS)
def test_philox_split_invariance_and_seed(impl):
    """Same (seed, global window) -> identical samples however the sweep is cut into launches."""
    kernels, _ = _k()
    _, sd = load_wo_fx_checkpoint()
    dev = _dev()
    packed = _pack_ns(sd, 2)
    torch.manual_seed(4)
    gx = (torch.rand(4, 50, 2) * 0.06 + 0.01).to(dev)
    full = kernels.nsdiff_sample(packed, None, gx, 4, 1, 16, 4, 50, 2, 20, seed=99, window_base=10, impl=impl)
    lo = kernels.nsdiff_sample(packed, None, gx[:1].contiguous(), 1, 1, 16, 16, 50, 2, 20, seed=99, window_base=10, impl=impl)
    hi = kernels.nsdiff_sample(packed, None, gx[1:].contiguous(), 3, 1, 16, 8, 50, 2, 20, seed=99, window_base=11, impl=impl)
    assert torch.equal(full[:1], lo) and torch.equal(full[1:], hi)
    other = kernels.nsdiff_sample(packed, None, gx, 4, 1, 16, 4, 50, 2, 20, seed=100, window_base=10, impl=impl)
    assert not torch.equal(full, other)


def test_philox_statistics_match_reference_distribution():
    """MPV from in-kernel Philox noise agrees with the oracle driven by torch noise (K=2000)."""
    kernels, _ = _k()
    net_param, sd = load_wo_fx_checkpoint()
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    dev = _dev()
    packed = _pack_ns(sd, 2)
    gx1 = g["gx"][:, :40] + 1e-7                                  # [1,40,2]
    K = 2000
    out = kernels.nsdiff_sample(packed, None, gx1.to(dev).contiguous(), 1, 1, K, K, 40, 2, 20, seed=2024)
    torch.manual_seed(0)
    sched = nsdiff_oracle.nsdiff_schedule()
    gxt = nsdiff_oracle.tile_rows(gx1, K)
    ref = nsdiff_oracle.p_sample_loop(sd, sched, torch.zeros_like(gxt), gxt, torch.zeros_like(gxt), 20, torch.randn_like)[-1]
    v_gpu = out.reshape(K, 40, 2).var(dim=0, unbiased=False).mean().item()
    v_ref = ref.var(dim=0, unbiased=False).mean().item()
    m_gpu = out.mean().item()
    assert abs(v_gpu - v_ref) / v_ref < 0.03, (v_gpu, v_ref)       # ~ 1/sqrt(K*40*2/corr) sampling error
    assert abs(m_gpu) < 0.02


# ---------------------------------------------------------------- g(x)
def test_sigma_estimation_golden():
    kernels, _ = _k()
    net_param, sd = load_wo_fx_checkpoint()
    dev = _dev()
    ws = [sd["cond_pred_model_g.mlp.%s" % k].to(dev).contiguous() for k in
          ("0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias", "5.weight", "5.bias", "6.weight", "6.bias")]
    g = load_golden("sigma_estimation_wo_fx.npz")
    gx = kernels.sigma_estimation(ws, g["x"].to(dev), 100, 200).cpu()
    np.testing.assert_allclose(gx.numpy(), g["gx"].numpy(), rtol=2e-5, atol=1e-7)
    g2 = load_golden("evalstep_wo_fx_k8s4.npz")
    gx2 = kernels.sigma_estimation(ws, g2["window_scaled"].to(dev), 100, 200, add_eps=1e-7).cpu()
    np.testing.assert_allclose(gx2.numpy(), (g2["gx"] + 1e-7).numpy(), rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("rows,Lw,R,F,O", [(1, 100, 50, 1, 100), (7, 100, 50, 1, 100), (5, 60, 20, 3, 33), (9, 200, 100, 4, 200)])
def test_sigma_estimation_vs_oracle(rows, Lw, R, F, O):
    kernels, _ = _k()
    torch.manual_seed(rows * 7 + F)
    H = 512
    n_in = Lw - R
    P = "cond_pred_model_g."
    sd = {P + "mlp.0.weight": torch.randn(H, n_in) / n_in ** 0.5, P + "mlp.0.bias": torch.randn(H) * 0.1,
          P + "mlp.2.weight": torch.rand(F, H) + 0.5, P + "mlp.2.bias": torch.randn(F, H) * 0.1,
          P + "mlp.3.weight": torch.randn(H, H) / H ** 0.5, P + "mlp.3.bias": torch.randn(H) * 0.1,
          P + "mlp.5.weight": torch.rand(F, H) + 0.5, P + "mlp.5.bias": torch.randn(F, H) * 0.1,
          P + "mlp.6.weight": torch.randn(O, H) / H ** 0.5, P + "mlp.6.bias": torch.randn(O) * 0.1}
    x = torch.randn(rows, Lw, F).cumsum(dim=1) * 0.2
    ref = sigma_oracle.sigma_estimation(sd, x, R, O)
    dev = _dev()
    ws = [sd[P + "mlp.%s" % k].to(dev).contiguous() for k in
          ("0.weight", "0.bias", "2.weight", "2.bias", "3.weight", "3.bias", "5.weight", "5.bias", "6.weight", "6.bias")]
    gx = kernels.sigma_estimation(ws, x.to(dev), R, O).cpu()
    np.testing.assert_allclose(gx.numpy(), ref.numpy(), rtol=5e-5, atol=1e-6)


# ---------------------------------------------------------------- MPV reduction
def test_mpv_reduce_golden():
    kernels, _ = _k()
    g = load_golden("evalstep_wo_fx_k8s4.npz")
    dev = _dev()
    traj = g["outs"].permute(0, 3, 1, 2).contiguous().to(dev)        # [B=1,K,O,F]
    r = kernels.mpv_reduce(traj, 1, 1, want_var=True, want_mean=True)
    assert r["mpv"].item() == pytest.approx(float(g["net_ews_nomodel"]), rel=MPV_RTOL)
    assert r["pred_mean"].item() == pytest.approx(float(g["net_pred_mean_nomodel"]), rel=1e-4, abs=1e-7)
    assert r["mpv_f"][0, 0].item() == pytest.approx(float(g["slbp_mpv_dim0"]), rel=MPV_RTOL)
    assert r["mpv_f"][0, 1].item() == pytest.approx(float(g["fig6_mpv_dim1"]), rel=MPV_RTOL)
    scale = torch.stack([g["scaler_mean"], g["scaler_std"]]).to(dev).contiguous()
    r2 = kernels.mpv_reduce(traj, 1, 1, scale=scale)
    assert r2["mpv"].item() == pytest.approx(float(g["net_ews"]), rel=MPV_RTOL)
    assert r2["pred_mean"].item() == pytest.approx(float(g["net_pred_mean"]), rel=MPV_RTOL)
    tgt = ((g["target_raw"] - g["scaler_mean"]) / g["scaler_std"]).to(dev)
    err = (r["mean"].reshape(200, 2) - tgt).abs().mean(dim=0)[0].item()
    assert err == pytest.approx(float(g["slbp_err_dim0"]), rel=1e-5)


@pytest.mark.parametrize("n_win,B,K,O,F", [(1, 1, 1, 1, 1), (3, 2, 7, 5, 3), (4, 3, 100, 100, 1), (2, 1, 100, 200, 2),
                                           (5, 4, 33, 37, 2), (2, 2, 9, 3, 4)])
def test_mpv_reduce_vs_oracle_ragged(n_win, B, K, O, F):
    kernels, _ = _k()
    torch.manual_seed(n_win * 100 + K)
    traj = torch.randn(n_win * B, K, O, F) * torch.rand(n_win * B, 1, O, F) + torch.randn(n_win * B, 1, O, F) * 3
    dev = _dev()
    mean = torch.randn(F)
    std = torch.rand(F) + 0.5
    for scale in (None, torch.stack([mean, std])):
        r = kernels.mpv_reduce(traj.to(dev), n_win, B, scale=None if scale is None else scale.to(dev).contiguous(), want_var=True)
        for w in range(n_win):
            elem = traj[w * B:(w + 1) * B].permute(0, 2, 3, 1).numpy()          # [B,O,F,K]
            pm, mpv = mpv_oracle.network_mpv(elem, *( (None, None) if scale is None else (mean.numpy(), std.numpy())))
            assert r["mpv"][w].item() == pytest.approx(mpv, rel=MPV_RTOL, abs=1e-9)
            assert r["pred_mean"][w].item() == pytest.approx(pm, rel=1e-4, abs=1e-5)
            if scale is None:
                for f in range(F):
                    want = np.asarray(elem, np.float64).var(axis=-1)[:, :, f].mean()
                    assert r["mpv_f"][w, f].item() == pytest.approx(want, rel=MPV_RTOL, abs=1e-9)


def test_mpv_linearity_property_full_size():
    """BASELINE config-1 window size (K=100, O=200, F=2) x 64 windows: var(a*x+b) = a^2 var(x)."""
    kernels, _ = _k()
    dev = _dev()
    torch.manual_seed(8)
    traj = torch.randn(64, 100, 200, 2, device=dev)
    r1 = kernels.mpv_reduce(traj, 64, 1)
    r2 = kernels.mpv_reduce(traj * 3.0 + 5.0, 64, 1)
    torch.testing.assert_close(r2["mpv"], r1["mpv"] * 9.0, rtol=2e-5, atol=0)
    torch.testing.assert_close(r2["pred_mean"], r1["pred_mean"] * 3.0 + 5.0, rtol=2e-5, atol=1e-5)
    ref = traj.double().var(dim=1, unbiased=False).mean(dim=(1, 2)).float()
    torch.testing.assert_close(r1["mpv"], ref, rtol=MPV_RTOL, atol=0)


# ---------------------------------------------------------------- error behaviour of the C ABI
def test_c_abi_rejects_bad_arguments():
    kernels, _ = _k()
    _, sd = load_wo_fx_checkpoint()
    dev = _dev()
    packed = _pack_ns(sd, 2)
    gx = torch.rand(1, 10, 2, device=dev)
    with pytest.raises(RuntimeError, match="bad argument"):
        kernels.nsdiff_sample(packed, None, gx, 1, 1, 0, 1, 10, 2, 20)
    with pytest.raises(RuntimeError, match="unsupported"):
        kernels.nsdiff_sample(packed, None, gx, 1, 1, 1, 1, 10, 5, 20)
    with pytest.raises(RuntimeError, match="bad argument"):
        kernels.nsdiff_sample(packed, None, gx, 1, 1, 6, 4, 10, 2, 20, noise=torch.zeros(4, device=dev))


# ------------------------------------------------------------------------------------------ f(x) glue kernels
def _a3_decode(a3, K):
    """[hi | lo | hi | 1 1 0..] -> (hi + lo as fp32, tail)."""
    a = a3.float()
    assert torch.equal(a3[:, :K], a3[:, 2 * K:3 * K])
    return a[:, :K] + a[:, K:2 * K], a[:, 3 * K:]


@pytest.mark.parametrize("K,act", [(512, 0), (256, 2), (256, 1), (64, 0)])
def test_fx_split_operand(K, act):
    from updgm_b200 import fx_encoder
    torch.manual_seed(K + act)
    x = torch.randn(1000, K, device=_dev()) * 3
    a3 = fx_encoder.a3_split(x, act=act)
    ref = {0: x, 1: torch.relu(x), 2: torch.nn.functional.gelu(x)}[act]
    val, tail = _a3_decode(a3, K)
    assert float((val - ref).abs().max() / ref.abs().max()) < 2e-6          # 22 mantissa bits
    assert torch.equal(tail, torch.tensor([1., 1., 0, 0, 0, 0, 0, 0], device=_dev()).expand(1000, 8))
    # head merge: attention output [B,H,L,dk] -> rows (b,l), columns (h,j)
    B, H, L, dk = 3, 8, 50, K // 8
    o = torch.randn(B, H, L, dk, device=_dev())
    val, _ = _a3_decode(fx_encoder.a3_split(o, heads=(B, H, L)), K)
    assert float((val - o.transpose(1, 2).reshape(B * L, K)).abs().max()) < 2e-6 * float(o.abs().max())


@pytest.mark.parametrize("K", [64, 128, 512])
def test_fx_add_layernorm_split(K):
    from updgm_b200 import fx_encoder
    torch.manual_seed(K)
    dev = _dev()
    x, r = torch.randn(777, K, device=dev), torch.randn(777, K, device=dev)
    ln1, ln2 = torch.nn.LayerNorm(K).to(dev), torch.nn.LayerNorm(K).to(dev)
    for ln in (ln1, ln2):
        ln.weight.data.uniform_(0.5, 1.5)
        ln.bias.data.uniform_(-0.5, 0.5)
    with torch.no_grad():
        y, a3 = fx_encoder.add_ln_split(x, r, ln1)
        ref = ln1(x + r)
        assert float((y - ref).abs().max()) < 5e-6
        val, _ = _a3_decode(a3, K)
        assert float((val - ref).abs().max()) < 5e-6
        y2, a32 = fx_encoder.add_ln_split(x, None, ln1, ln2, want_y=False)
        assert y2 is None
        val, _ = _a3_decode(a32, K)
        assert float((val - ln2(ln1(x))).abs().max()) < 1e-5


def test_fx_compensated_gemm_accuracy():
    """One fp16 GEMM on the split operand reproduces the fp32 linear layer to ~1e-6 (bias inside the GEMM)."""
    from updgm_b200 import fx_encoder
    torch.manual_seed(1)
    lin = fx_encoder.SLinear(512, 264).to(_dev())
    x = torch.randn(4096, 512, device=_dev())
    with torch.no_grad():
        y = fx_encoder.gemm3(fx_encoder.a3_split(x), lin.w3(), 264)
        ref = (x.double() @ lin.weight.double().t() + lin.bias.double())
    assert tuple(y.shape) == (4096, 264) and float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-5


@pytest.mark.parametrize("Lq,S,causal,use_delta", [(100, 100, False, True), (150, 150, True, False), (150, 100, False, True),
                                                   (7, 33, False, False), (130, 192, False, True)])
def test_fx_tcgen05_attention_against_fp64(Lq, S, causal, use_delta):
    """upd_fx_attention == softmax(scale*(tau*QK^T + delta)) V in float64, read back from the split operand it emits."""
    import ctypes
    import math
    from updgm_b200 import _lib
    torch.manual_seed(Lq * 1000 + S)
    dev = _dev()
    B, H, dk = 3, 8, 64
    d = H * dk
    qkv = torch.randn(B * max(Lq, S), 3 * d, device=dev)          # fused-projection style buffer: strided heads
    tau = torch.rand(B, device=dev) * 1.5 + 0.5
    scale = 1.0 / math.sqrt(dk)
    pitch = (S + 15) // 16 * 16
    dbuf = torch.zeros(B, pitch, device=dev)
    delta = torch.randn(B, S, device=dev)
    dbuf[:, :S] = delta * scale
    a3 = torch.full((B * Lq, 3 * d + 8), float("nan"), dtype=torch.float16, device=dev)
    q_rows = qkv[: B * Lq]
    kv_rows = qkv[: B * S]
    rc = _lib.lib().upd_fx_attention(
        _lib.ptr(qkv), 3 * d, ctypes.c_void_p(qkv.data_ptr() + 4 * d), ctypes.c_void_p(qkv.data_ptr() + 8 * d), 3 * d,
        _lib.ptr(tau), ctypes.c_void_p(dbuf.data_ptr()) if use_delta else None, pitch, B, H, Lq, S, dk, int(causal),
        scale, _lib.ptr(a3), _lib.stream_ptr(dev))
    _lib.check(rc, "upd_fx_attention")
    q = q_rows[:, :d].double().view(B, Lq, H, dk).transpose(1, 2)
    k = kv_rows[:, d:2 * d].double().view(B, S, H, dk).transpose(1, 2)
    v = kv_rows[:, 2 * d:].double().view(B, S, H, dk).transpose(1, 2)
    sc = (q @ k.transpose(-1, -2)) * tau.double().view(B, 1, 1, 1)
    if use_delta:
        sc = sc + delta.double().view(B, 1, 1, S)
    sc = sc * scale
    if causal:
        sc = sc.masked_fill(torch.ones(Lq, S, dtype=torch.bool, device=dev).triu(1), float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * Lq, d)
    val, tail = _a3_decode(a3, d)
    assert torch.isfinite(val).all()
    assert float((val.double() - ref).abs().max() / ref.abs().max()) < 2e-5
    assert torch.equal(tail, torch.tensor([1., 1., 0, 0, 0, 0, 0, 0], device=dev).expand(B * Lq, 8))


@pytest.mark.parametrize("impl", [4])
def test_warp_specialised_kernel_full_bench_shape_against_two_tile(impl):
    """BASELINE config-2 shape at a size where every SM runs many tile pairs (30 windows x 100 rows x K=100 x O=100 =
    30 M rows): the warp-specialised kernel reproduces the two-tile one (same Philox noise) to reordering error."""
    kernels, schedules = _k()
    dev = _dev()
    g = load_golden("psample_loop_randF1.npz")
    tab = schedules.nsdiff_tables("linear", 20, 1e-4, 0.02)
    packed = kernels.pack_denoiser(g["sd"], 0, 1, 20, schedules.stack_rows(tab, schedules.NSDIFF_ROWS), dev)
    torch.manual_seed(8)
    n_win, B, K, O = 30, 100, 100, 100
    gx = torch.rand(n_win * B, O, 1, device=dev) * 0.3 + 0.05
    y0 = torch.randn(n_win * B, O, 1, device=dev)
    a = kernels.nsdiff_sample(packed, y0, gx, n_win, B, K, 10, O, 1, 20, seed=5, window_base=3, impl=2)
    b = kernels.nsdiff_sample(packed, y0, gx, n_win, B, K, 10, O, 1, 20, seed=5, window_base=3, impl=impl)
    rms = float(a.pow(2).mean().sqrt())
    assert torch.isfinite(b).all()
    assert float((a - b).abs().max()) / rms < 2e-5


@pytest.mark.parametrize("Lq,S,causal,use_delta", [(100, 100, False, True), (150, 150, True, False), (150, 100, False, True)])
def test_fx_attention_head_size_16_against_fp64(Lq, S, causal, use_delta):
    """upd_fx_attention_hs16 (TMDM's condition encoder: 4 heads of 16) == softmax(scale*(tau*QK^T + delta)) V in float64."""
    import ctypes
    import math
    from updgm_b200 import _lib
    torch.manual_seed(Lq + 7 * S)
    dev = _dev()
    B, H, dk = 5, 4, 16
    d = H * dk
    qkv = torch.randn(B * max(Lq, S), 3 * d, device=dev)
    tau = torch.rand(B, device=dev) * 1.5 + 0.5
    scale = 1.0 / math.sqrt(dk)
    pitch = (S + 15) // 16 * 16
    dbuf = torch.zeros(B, pitch, device=dev)
    delta = torch.randn(B, S, device=dev)
    dbuf[:, :S] = delta * scale
    o = torch.empty(B * Lq, d, device=dev)
    rc = _lib.lib().upd_fx_attention_hs16(
        _lib.ptr(qkv), 3 * d, ctypes.c_void_p(qkv.data_ptr() + 4 * d), ctypes.c_void_p(qkv.data_ptr() + 8 * d), 3 * d,
        _lib.ptr(tau), ctypes.c_void_p(dbuf.data_ptr()) if use_delta else None, pitch, B, H, Lq, S, int(causal), scale,
        _lib.ptr(o), _lib.stream_ptr(dev))
    _lib.check(rc, "upd_fx_attention_hs16")
    q = qkv[: B * Lq, :d].double().view(B, Lq, H, dk).transpose(1, 2)
    k = qkv[: B * S, d:2 * d].double().view(B, S, H, dk).transpose(1, 2)
    v = qkv[: B * S, 2 * d:].double().view(B, S, H, dk).transpose(1, 2)
    sc = (q @ k.transpose(-1, -2)) * tau.double().view(B, 1, 1, 1)
    if use_delta:
        sc = sc + delta.double().view(B, 1, 1, S)
    sc = sc * scale
    if causal:
        sc = sc.masked_fill(torch.ones(Lq, S, dtype=torch.bool, device=dev).triu(1), float("-inf"))
    ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B * Lq, d)
    assert float((o.double() - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("M,n_out,K,with_add", [(1000, 512, 512, False), (129, 1, 512, False), (4096, 1536, 512, False),
                                                 (777, 64, 64, True), (2500, 320, 80, True), (128, 256, 256, False),
                                                 (300, 207, 200, False), (65, 40, 16, True)])
def test_gemm3_tcgen05_against_fp64(M, n_out, K, with_add):
    """upd_gemm3 (warp-specialised tcgen05 GEMM, TMA operands, TMEM accumulators) == A3 W3^T (+ addend) in float64 on
    the same fp16 operands, for every tile width (64 / 128 / 256), ragged M / N / K tails and the addend path; and the
    error-compensated operand reproduces x W^T + b to fp32 accuracy."""
    from updgm_b200 import fx_encoder
    dev = _dev()
    torch.manual_seed(M + n_out)
    x = torch.randn(M, K, device=dev)
    lin = torch.nn.Linear(K, n_out).to(dev)
    cache = fx_encoder._W3Cache()
    w3 = cache.get([(lin.weight, lin.bias)])
    a3 = fx_encoder.a3_split(x)
    add = torch.randn(M, n_out, device=dev) if with_add else None
    y = fx_encoder.gemm3(a3, w3, n_out, addend=add)
    ref = a3.double() @ w3.double().t()[:, :n_out] + (0 if add is None else add.double())
    assert torch.isfinite(y).all()
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-5      # fp32 accumulation over 3K+8 terms
    full = x.double() @ lin.weight.double().t() + lin.bias.double() + (0 if add is None else add.double())
    assert float((y.double() - full).abs().max() / full.abs().max()) < 1e-5


def test_long_step_counts_run_on_the_ffma_kernel_instead_of_failing():
    """include/upd_b200.h: the tcgen05 samplers keep three [T,128] step-embedding tables next to the 173 KB weight image
    in shared memory, which fits T <= ~40 (every shipped YAML has T = 20).  A longer schedule is not an error for the
    default implementation: UPD_IMPL_TCGEN05 runs it on the FFMA kernel; the forced tcgen05 variants say "unsupported"."""
    kernels, _ = _k()
    torch.manual_seed(9)
    T, F, O, K = 56, 1, 30, 4
    sd = _random_ns_weights(F, T)
    dev = _dev()
    packed = _pack_ns(sd, F, T)
    gx = (torch.rand(3, O, F) * 0.3 + 0.05).to(dev)
    a = kernels.nsdiff_sample(packed, None, gx, 3, 1, K, 2, O, F, T, seed=5, impl=0)
    b = kernels.nsdiff_sample(packed, None, gx, 3, 1, K, 2, O, F, T, seed=5, impl=1)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    for impl in (2, 4):
        with pytest.raises(RuntimeError, match="unsupported"):
            kernels.nsdiff_sample(packed, None, gx, 3, 1, K, 2, O, F, T, seed=5, impl=impl)
    # and F = 3 (outside the warp-specialised kernel) takes the two-tile tcgen05 kernel under the default
    sd3 = _random_ns_weights(3, 20)
    p3 = _pack_ns(sd3, 3, 20)
    gx3 = (torch.rand(2, O, 3) * 0.3 + 0.05).to(dev)
    c = kernels.nsdiff_sample(p3, None, gx3, 2, 1, K, 2, O, 3, 20, seed=6, impl=0)
    d = kernels.nsdiff_sample(p3, None, gx3, 2, 1, K, 2, O, 3, 20, seed=6, impl=2)
    assert torch.equal(c, d)
