"""Mirror of fitburst.routines for the hot path (profiles and derivative maps backed by the CUDA kernels)."""
from . import derivative, profile  # noqa: F401
