"""
Mirror of fitburst.routines.manipulate (reference routines/manipulate.py:10-148): block-mean
down-sampling on the device (fb_pre_downsample_2d, the kernel ReaderBaseClass.downsample uses)
and the sub-bin label arithmetic of the up-samplers. Same signatures as the reference.

The up-samplers produce O(len * factor) axis LABELS; they are host arithmetic in the reference's
own operation order (np.linspace: start + arange * step, last element forced to stop) -- exactly
what fb_plan_create restates in C for the sub-frequency table -- because the labels feed Python
callers, not kernels.
"""
import numpy as np
import torch

from .. import _lib
from ._device import require_gpu, stream_of, to_device


def _block_mean(array2d, factor_freq, factor_time):
    lib, device = require_gpu()
    num_freq, num_time = array2d.shape
    src = to_device(array2d, device)
    out = torch.empty((num_freq // factor_freq, num_time // factor_time), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.fb_pre_downsample_2d(src.data_ptr(), num_freq, num_time, int(factor_freq), int(factor_time),
                                            out.data_ptr(), stream_of(device)))
    return out.cpu().numpy()


def downsample_1d(array_orig, factor: int, boolean: bool = False):
    """Mean over consecutive groups of `factor` elements (routines/manipulate.py:10-41)."""
    array = np.asarray(array_orig)
    if len(array) % factor != 0:  # what np.reshape raises in the reference
        raise ValueError(f"cannot reshape array of size {len(array)} into shape ({len(array) // factor},{factor})")
    out = _block_mean(array.astype(np.float64).reshape(1, -1), 1, factor)[0]
    return out.astype(bool) if boolean else out


def downsample_2d(spectrum_orig, factor_freq: int, factor_time: int):
    """Block mean over (factor_freq, factor_time) cells (routines/manipulate.py:43-72)."""
    spectrum = np.asarray(spectrum_orig, dtype=np.float64)
    num_freq, num_time = spectrum.shape
    if num_freq % factor_freq != 0 or num_time % factor_time != 0:
        raise ValueError(f"cannot reshape array of size {spectrum.size} into shape "
                         f"({num_freq // factor_freq},{factor_freq},{num_time // factor_time},{factor_time})")
    return _block_mean(spectrum, factor_freq, factor_time)


def upsample_1d(array_orig, diff_orig: float, factor: int):
    """Sub-bin centre labels of a regularly sampled axis (routines/manipulate.py:75-107)."""
    diff_new = diff_orig / factor
    bound_lo = array_orig[0] - (diff_orig / 2) + (diff_new / 2)
    bound_hi = array_orig[-1] + (diff_orig / 2) - (diff_new / 2)
    return np.linspace(bound_lo, bound_hi, num=len(array_orig) * factor)


def upsample_orig(input_array, factor: int):
    """(len, factor) sub-bin labels of the 'original' fitburst (routines/manipulate.py:110-148)."""
    input_array = np.asarray(input_array, dtype=np.float64)
    res_orig = np.diff(input_array)[1]
    offsets = ((np.arange(factor, dtype=float) + 0.5) / factor - 0.5) * res_orig
    return input_array[:, None] + offsets[None, :]
