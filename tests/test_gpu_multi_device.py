"""A model that lives on a GPU which is NOT the current device: every launch (f(x), g(x), the samplers, the reductions)
must go to the model's device and stream.  Needs two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os

import pytest
import torch
import yaml

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")


def _nsdiff(device):
    from updgm_b200.nsdiff import NsDiff_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "NsDiff", "biomass", "model_trained.yaml")))
    net = dict(cfg["net"], device=torch.device(device), n_z_samples=6, parallel_sample=3)
    torch.manual_seed(123)
    m = NsDiff_model(net, "NsDiff_model").eval()
    m.scaler_std.fill_(1.0)
    return m


def test_nsdiff_model_on_a_non_current_device_gives_the_same_trajectories():
    _need_two()
    import updgm_b200  # noqa: F401
    from updgm_b200 import uncertainty as U
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(4, 100, 1, generator=g) * 0.1).cumsum(dim=1) + 5.0
    noise = torch.randn(2, 20, 4 * 3, 100, 1, generator=g)
    torch.cuda.set_device(0)
    m0, m1 = _nsdiff("cuda:0"), _nsdiff("cuda:1")
    m1.load_state_dict({k: v.to("cuda:1") for k, v in m0.state_dict().items()})
    outs0, _ = m0.evaluation_step(x.to("cuda:0"), noise=noise)
    assert torch.cuda.current_device() == 0
    outs1, _ = m1.evaluation_step(x.to("cuda:1"), noise=noise)        # current device is still cuda:0
    assert torch.cuda.current_device() == 0
    assert torch.equal(outs0.cpu(), outs1.cpu())
    # the batched sweep (Philox noise): same seed and window base -> identical caches on either device
    wins = x.unsqueeze(0).repeat(3, 1, 1, 1)
    m0._windows_drawn = m1._windows_drawn = 0
    torch.manual_seed(7)                                              # the sweep's Philox seed is torch.initial_seed()
    c0 = U.sample_sweep(m0, wins, device=torch.device("cuda:0"))
    c1 = U.sample_sweep(m1, wins, device=torch.device("cuda:1"))
    assert torch.equal(torch.as_tensor(c0), torch.as_tensor(c1))


def test_tmdm_model_on_a_non_current_device():
    _need_two()
    from updgm_b200.tmdm import TMDM_model
    cfg = yaml.safe_load(open(os.path.join(GOLDEN, "ews_results", "model_compare", "TMDM", "neuronal", "model_trained.yaml")))
    torch.manual_seed(321)
    m0 = TMDM_model(dict(cfg["net"], device=torch.device("cuda:0"), n_z_samples=4, parallel_sample=2)).eval()
    m1 = TMDM_model(dict(cfg["net"], device=torch.device("cuda:1"), n_z_samples=4, parallel_sample=2)).eval()
    m1.load_state_dict({k: v.to("cuda:1") for k, v in m0.state_dict().items()})
    g = torch.Generator().manual_seed(4)
    x = torch.sigmoid((torch.randn(3, 100, 1, generator=g) * 0.2).cumsum(dim=1))
    torch.cuda.set_device(0)
    o0, _ = m0.evaluation_step(x.to("cuda:0"))
    o1, _ = m1.evaluation_step(x.to("cuda:1"))
    assert torch.cuda.current_device() == 0
    assert tuple(o0.shape) == tuple(o1.shape) and torch.isfinite(o1).all()
