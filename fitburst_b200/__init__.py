"""
fitburst_b200: B200-native (sm_100a) implementation of fitburst's dynamic-spectrum model
evaluation and least-squares fitting hot path, behind the reference's own class surface:

    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.analysis.fitter import LSFitter

All arithmetic runs in hand-written CUDA kernels reached through the C ABI declared in
include/fitburst_b200.h; PyTorch is used only to own device buffers. There is no CPU fallback.
"""
from . import analysis, backend, routines, synthetic  # noqa: F401

__version__ = "0.1.0"
