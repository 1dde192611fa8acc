"""Full-size parity against the ORACLE (not CUDA against CUDA): whole windows of BASELINE configs 1, 2 and 3 -- every
(row, sample, position) of a window as shipped, all T reverse steps, recorded noise -- through the implementation the
host actually selects, held to the tolerances BASELINE.json's north_star states literally:

    per trajectory value   |gpu - ref| <= 1e-3 * |ref| + 1e-4 * rms(ref)      (rel 1e-3; the floor covers values near 0)
    MPV of the window      |gpu - ref| <= 1e-4 * |ref|                          (diffusion_model_uncertainy.py:286-303)

Reference: nsdiff_utils.py:271-284 (p_sample_loop), NsDiff_model.py:180-268 / :404-495 (evaluation_step),
tmdm_adapter.py:116-155.  The same check is emitted by bench.py as the bench line's `parity` object.
"""
import os

import pytest
import torch
import yaml

from conftest import GOLDEN, load_wo_fx_checkpoint
from oracle import parity

pytestmark = pytest.mark.gpu


def _report(tag, c):
    return ("{}: {} values, max|d|/rms {:.2e}, worst |d|/bound {:.3f} ({}), MPV rel err {:.2e} (bound {:g})"
            .format(tag, c["values"], c["max_abs_err_over_rms"], c["worst_err_over_bound"], c["per_value_bound"],
                    c["mpv_rel_err"], c["mpv_bound"]))


def _assert_parity(tag, got, ref):
    c = parity.compare(got, ref)
    msg = _report(tag, c)
    print(msg)
    assert c["finite"], msg
    assert c["per_value_ok"], msg
    assert c["mpv_ok"], msg
    return c


def _bench_series(nodes=100, length=1000):
    g = torch.Generator().manual_seed(0)
    e = torch.randn(nodes, length, generator=g) * 0.1
    x = torch.zeros(nodes, length)
    for t in range(1, length):
        x[:, t] = 0.99 * x[:, t - 1] + e[:, t]
    return (x + 5.0).unsqueeze(-1)


def test_config2_full_window_sampler_and_model_against_oracle():
    """BASELINE config 2 (NsDiff, biomass YAML architecture, seeded random weights): one whole window = B 100 rows x K 100
    samples x O 100 positions = 1e6 denoiser rows x T 20 steps, S = 10 -> 10 chunks of recorded draws (2e7 normals)."""
    import updgm_b200  # noqa: F401
    from updgm_b200 import kernels
    from updgm_b200.nsdiff import NsDiff_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "NsDiff", "biomass", "model_trained.yaml")))
    dev = torch.device("cuda:0")
    net = dict(cfg["net"], device=dev)
    torch.manual_seed(123)
    m = NsDiff_model(net, "NsDiff_model").eval()
    m.scaler_std.fill_(1.0)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    x = _bench_series()[:, 400:500, :].contiguous()                                # window 80 of the bench sweep
    r = parity.nsdiff_window_reference(sd, dict(cfg["net"]), x, seed=11, with_fx=True, threads=os.cpu_count())
    B, O, F, K = r["ref"].shape
    assert (B, O, F, K) == (100, 100, 1, 100)
    # (1) the sampler alone at full size, same f(x) / g(x) as the oracle, the implementation the host selects
    traj = kernels.nsdiff_sample(m.packed_weights(), r["y0"].to(dev), r["gx"].to(dev), 1, B, K, 10, O, F, 20,
                                 noise=r["noise"].to(dev), impl=m.sampler_impl)
    _assert_parity("config 2, sampler (impl {})".format(m.sampler_impl), traj.permute(0, 2, 3, 1), r["ref"])
    # (2) the whole evaluation_step (f(x) and g(x) on the GPU too; f(x) is parity-unpinned: checked against its restatement)
    outs, _ = m.evaluation_step(x.to(dev), noise=r["noise"][0])
    _assert_parity("config 2, evaluation_step", outs, r["ref"])


def test_config1_full_windows_shipped_checkpoint_against_oracle():
    """BASELINE config 1 (NsDiff SLBP, the shipped wo_fx checkpoint: denoiser + g(x), no f(x), F = 2): whole windows =
    K 100 x O 200 = 20 000 rows x 20 steps each, S = 100 (one chunk), four windows of a synthetic scaled series."""
    import updgm_b200  # noqa: F401
    from updgm_b200 import uncertainty as U
    net_param, sd = load_wo_fx_checkpoint()
    dev = torch.device("cuda:0")
    m, _ = U.load_model_from_dir(os.path.join(GOLDEN, "ews_results", "NsDiff_machine", "wo_fx"), device=dev)
    g = torch.Generator().manual_seed(3)
    e = torch.randn(1000, 2, generator=g) * 0.1
    s = torch.zeros(1000, 2)
    for t in range(1, 1000):
        s[t] = 0.99 * s[t - 1] + e[t]
    net = {k: v for k, v in net_param.items() if k != "device"}
    for w, start in enumerate((0, 250, 500, 800)):
        x = s[start:start + 200].unsqueeze(0).contiguous()                          # [1, 200, 2], scaled units
        r = parity.nsdiff_window_reference(sd, net, x, seed=20 + w, with_fx=False, variant_adds_eps=True, threads=os.cpu_count())
        assert tuple(r["ref"].shape) == (1, 200, 2, 100)
        outs, _ = m.evaluation_step(x.to(dev), noise=r["noise"][0])
        _assert_parity("config 1, window {} (impl {})".format(w, m.sampler_impl), outs, r["ref"])


def test_config3_full_window_tmdm_against_oracle():
    """BASELINE config 3 (TMDM, neuronal YAML architecture, seeded random weights): one whole window = B 100 x K 100 x
    (label 50 + pred 100) positions = 1.5e6 rows x 20 steps; the sampler is fed the oracle's condition mean."""
    import updgm_b200  # noqa: F401
    from updgm_b200.tmdm import TMDM_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "TMDM", "neuronal", "model_trained.yaml")))
    dev = torch.device("cuda:0")
    torch.manual_seed(321)
    m = TMDM_model(dict(cfg["net"], device=dev)).eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    x = torch.sigmoid((torch.randn(100, 100, 1, generator=g) * 0.2).cumsum(dim=1))
    r = parity.tmdm_window_reference(sd, dict(m.configs.__dict__, device="cpu"), x, seed=31, threads=os.cpu_count())
    assert tuple(r["ref"].shape) == (100, 100, 1, 100)
    traj = m.sample_windows(x.unsqueeze(0).to(dev), noise=r["noise"], y_0_hat=r["y0"])
    _assert_parity("config 3, sampler (impl {})".format(m.sampler_impl), traj.permute(0, 2, 3, 1), r["ref"])
    outs, _ = m.evaluation_step(x.to(dev), noise=r["noise"][0])
    _assert_parity("config 3, evaluation_step", outs, r["ref"])
