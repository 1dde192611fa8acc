"""Run the UNMODIFIED reference in the build container (test infrastructure, never shipped).

/root/reference is read-only and only exists in the build container; the GPU box never sees
it.  The reference's hot path needs torch_geometric / torch_timeseries / igraph / matplotlib /
torch_sparse only at import time; ``oracle/_stubs`` holds arithmetic-free stand-ins so the
real reference modules import and run on the CPU.  Used by ``oracle/make_golden.py`` (fixture
generator) and by CPU tests that skip when the reference is absent.
"""
import contextlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("UPD_REFERENCE_ROOT", "/root/reference")
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models", "Diffusion_model"))


def activate():
    """Put the stubs and the reference on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("reference tree not present at {}".format(REFERENCE_ROOT))
    for p in (REFERENCE_ROOT, STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)


def analysis_module():
    activate()
    from evaluation_and_analysis import diffusion_model_uncertainy as dmu
    return dmu


class NoiseTape:
    """Record or replay every ``torch.randn_like`` draw the reference makes."""

    def __init__(self, replay=None):
        self.draws = []
        self._replay = list(replay) if replay is not None else None

    def __call__(self, like, *a, **k):
        if self._replay is not None:
            z = self._replay.pop(0)
            assert z.shape == like.shape, (z.shape, like.shape)
            z = z.to(like.dtype)
        else:
            z = self._orig(like, *a, **k)
        self.draws.append(z.clone())
        return z

    @contextlib.contextmanager
    def patched(self):
        self._orig = torch.randn_like
        torch.randn_like = self
        try:
            yield self
        finally:
            torch.randn_like = self._orig
