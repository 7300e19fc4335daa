"""
Mirror of fitburst.routines.ism (reference routines/ism.py:11-140): dispersion delay / smearing,
thin-screen scattering time and the per-channel scintillation amplitudes, through the C ABI
(fb_ism_times, fb_amplitude_per_channel). Same signatures as the reference.
"""
import numpy as np
import torch

from .. import _lib
from ._device import elementwise, require_gpu, stream_of, to_device


def compute_amplitude_per_channel(data, model):
    """
    Per-channel amplitudes of the model components from the observed spectrum of ONE channel
    (routines/ism.py:11-48): solves (P^T P) a = P^T d with data (num_time,), model (num_time,
    num_component). Raises numpy.linalg.LinAlgError for a singular system like np.linalg.solve.
    """
    lib, device = require_gpu()
    model = np.asarray(model, dtype=np.float64)
    num_time, num_component = model.shape
    data_dev, model_dev = to_device(np.asarray(data).reshape(num_time), device), to_device(model, device)
    out = torch.empty(num_component, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.fb_amplitude_per_channel(data_dev.data_ptr(), model_dev.data_ptr(), num_time, num_component,
                                                out.data_ptr(), stream_of(device)))
    return out.cpu().numpy()


def compute_time_dm_delay(dm_value: float, dm_const: float, dm_idx: float, freq1, freq2: float = np.inf):
    """dm_value * dm_const * (freq1 ** dm_idx - freq2 ** dm_idx) (routines/ism.py:50-80)."""
    return elementwise(
        lambda lib, src, count, dst, stream: lib.fb_ism_times(
            0, src, count, float(dm_value), float(dm_const), float(dm_idx), float(freq2), dst, stream),
        freq1)


def compute_time_dm_smear(dm_value: float, dm_const: float, dm_idx: float, freq, width_chan: float):
    """2 * dm_value * dm_const * width_chan * freq ** (dm_idx - 1) (routines/ism.py:82-112)."""
    return elementwise(
        lambda lib, src, count, dst, stream: lib.fb_ism_times(
            1, src, count, float(dm_value), float(dm_const), float(dm_idx), float(width_chan), dst, stream),
        freq)


def compute_time_scattering(freq, ref_freq: float, sc_time_ref: float, sc_idx: float):
    """sc_time_ref * (freq / ref_freq) ** sc_idx (routines/ism.py:114-140)."""
    return elementwise(
        lambda lib, src, count, dst, stream: lib.fb_ism_times(
            2, src, count, float(ref_freq), float(sc_time_ref), float(sc_idx), 0.0, dst, stream),
        freq)
