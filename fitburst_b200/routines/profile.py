"""
Mirror of fitburst.routines.profile (reference routines/profile.py:15-160): temporal profile
shapes evaluated by the CUDA kernel behind fb_profile_eval. Same signatures as the reference.
"""
import numpy as np
import torch

from .. import _lib


def _evaluate(time, toa, width, freq, ref_freq, sc_time_ref, sc_index, scattered, clamp, folded, normalize):
    if not torch.cuda.is_available():
        raise _lib.FitburstB200Error("no CUDA device visible: fitburst_b200 has no CPU fallback")
    time = np.asarray(time, dtype=np.float64)
    freq = np.asarray(freq, dtype=np.float64)
    shape = np.broadcast_shapes(freq.shape, time.shape)
    time_b = np.ascontiguousarray(np.broadcast_to(time, shape)).reshape(-1, shape[-1] if shape else 1)
    freq_b = np.ascontiguousarray(np.broadcast_to(freq, shape)).reshape(-1, shape[-1] if shape else 1)[:, 0].copy()
    device = torch.device("cuda")
    t_dev = torch.from_numpy(time_b.copy()).to(device)
    f_dev = torch.from_numpy(freq_b).to(device)
    out = torch.empty_like(t_dev)
    with torch.cuda.device(device):
        _lib.check(_lib.load_library().fb_profile_eval(
            t_dev.data_ptr(), f_dev.data_ptr(), time_b.shape[0], time_b.shape[1], float(toa), float(sc_time_ref),
            float(sc_index), float(width), float(ref_freq), int(scattered), int(clamp), int(folded),
            int(bool(normalize)), out.data_ptr(), torch.cuda.current_stream(device).cuda_stream))
    return out.cpu().numpy().reshape(shape)


def compute_profile_gaussian(values: float, mean: float, width: float, normalize: bool = False) -> float:
    """Gaussian profile exp(-0.5 ((values - mean) / width) ** 2) (routines/profile.py:15-49)."""
    values = np.asarray(values, dtype=np.float64)
    profile = _evaluate(values, mean, width, np.ones(values.shape[:-1] + (1,)) if values.ndim else np.ones(()),
                        1.0, 0.0, 0.0, False, False, False, False)
    if normalize:  # the reference divides by sqrt(2 pi) / width (routines/profile.py:45)
        profile = profile / (np.sqrt(2 * np.pi) / width)
    return profile


def compute_profile_pbf(time: float, toa: float, width: float, freq: float, ref_freq: float,
                        sc_time_ref: float, sc_index: float = -4., normalize: bool = False) -> float:
    """Gaussian convolved with a one-sided exponential (routines/profile.py:52-114)."""
    return _evaluate(time, toa, width, freq, ref_freq, sc_time_ref, sc_index, True, False, False, normalize)
