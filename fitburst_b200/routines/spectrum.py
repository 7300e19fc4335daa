"""
Mirror of fitburst.routines.spectrum (reference routines/spectrum.py:11-43): the running power
law of the spectral energy distribution, evaluated by the same device arithmetic the model kernels
use (fb_spectrum_rpl). Same signature as the reference.
"""
from ._device import elementwise


def compute_spectrum_rpl(freqs, freq_ref: float, sp_idx: float, sp_run: float):
    """exp(sp_idx * log(freqs / freq_ref) + sp_run * log(freqs / freq_ref) ** 2)."""
    return elementwise(
        lambda lib, src, count, dst, stream: lib.fb_spectrum_rpl(
            src, count, float(freq_ref), float(sp_idx), float(sp_run), dst, stream),
        freqs)
