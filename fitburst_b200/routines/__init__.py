"""Mirror of fitburst.routines for the hot path (profiles, spectra, ISM delays, resampling and the
derivative maps, all backed by the CUDA kernels behind include/fitburst_b200.h)."""
from . import derivative, ism, manipulate, profile, spectrum  # noqa: F401
