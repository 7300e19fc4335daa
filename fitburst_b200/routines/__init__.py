"""Mirror of fitburst.routines for the hot path (derivative maps backed by the CUDA kernels)."""
from . import derivative  # noqa: F401
