"""B200-native (sm_100a) uncertainty-inference hot path: conditional reverse-diffusion sampling and
the MPV / gx reductions, behind the reference's own model-loading and inference API.

Importable as ``updgm_b200`` (the directory name carries the reference's full name and is not a
valid Python identifier; ``updgm_b200/__init__.py`` at the repo root aliases it).
"""
from . import _build, _lib, kernels, schedules  # noqa: F401

__all__ = ["_build", "_lib", "kernels", "schedules"]
