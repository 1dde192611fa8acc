import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> dict of torch tensors; keys starting with 'sd:' are gathered into out['sd']."""
    z = np.load(os.path.join(GOLDEN, name))
    out, sd = {}, {}
    for k in z.files:
        v = torch.from_numpy(np.asarray(z[k]))
        if k.startswith("sd:"):
            sd[k[3:]] = v
        else:
            out[k] = v
    if sd:
        out["sd"] = sd
    return out


def load_wo_fx_checkpoint():
    path = os.path.join(GOLDEN, "ews_results", "NsDiff_machine", "wo_fx", "model_trained")
    with open(path, "rb") as f:
        state = torch.load(f, map_location="cpu", weights_only=False)
    return state["net_param"], state["state_dict"]


@pytest.fixture(scope="session")
def wo_fx():
    return load_wo_fx_checkpoint()
