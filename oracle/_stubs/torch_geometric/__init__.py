"""Import-only stub (no arithmetic): lets the read-only reference import in this container."""
from . import data, nn, utils, loader  # noqa: F401
