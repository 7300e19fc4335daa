"""
The data product of fitburst.utilities.plotting (reference utilities/plotting.py:19-90):
compute_downsampled_data(), the 2-D block means of spectrum / model / residuals that the summary
plots draw. The O(num_freq x num_time) work -- masking the flagged channels and the block means --
runs on the device (fb_pre_zero_channels, fb_pre_downsample_2d); drawing itself
(plot_summary_triptych) is matplotlib and is out of scope: the dictionary returned here is what the
reference's plotting functions consume.
"""
import numpy as np
import torch

from .. import _lib
from ..routines import manipulate as manip
from ..routines._device import require_gpu, stream_of


def _masked_block_mean(lib, device, spectrum, keep, factor_freq, factor_time):
    """downsample_2d(spectrum * keep[:, None]) with the spectrum given as array or device tensor."""
    if isinstance(spectrum, torch.Tensor):
        work = spectrum.to(device=device, dtype=torch.float64).clone()
    else:
        work = torch.from_numpy(np.ascontiguousarray(spectrum, dtype=np.float64)).to(device)
    num_freq, num_time = work.shape
    if num_freq % factor_freq != 0 or num_time % factor_time != 0:
        raise ValueError(f"cannot reshape array of size {num_freq * num_time} into shape "
                         f"({num_freq // factor_freq},{factor_freq},{num_time // factor_time},{factor_time})")
    out = torch.empty((num_freq // factor_freq, num_time // factor_time), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        if keep is not None:
            keep_dev = torch.from_numpy(np.ascontiguousarray(keep, dtype=np.uint8)).to(device)
            _lib.check(lib.fb_pre_zero_channels(work.data_ptr(), keep_dev.data_ptr(), num_freq, num_time,
                                                stream_of(device)))
        _lib.check(lib.fb_pre_downsample_2d(work.data_ptr(), num_freq, num_time, int(factor_freq), int(factor_time),
                                            out.data_ptr(), stream_of(device)))
    return out.cpu().numpy()


def compute_downsampled_data(times, freqs, spectrum_data, good_freq, spectrum_model=None, factor_freq: int = 1,
                             factor_time: int = 1) -> dict:
    """
    Down-sampled arrays and matrices for the plotting functions (utilities/plotting.py:19-90, same
    keys). spectrum_data / spectrum_model may be numpy arrays or device tensors. Like the reference
    the model entry requires a model: without one the reference raises UnboundLocalError at the
    dictionary; here "model" is None in that case.
    """
    # pylint: disable=too-many-arguments,too-many-locals
    lib, device = require_gpu()
    good_freq = np.asarray(good_freq)
    good_freq_downsampled = manip.downsample_1d(good_freq, factor_freq, boolean=True)
    # a flagged channel must contribute exactly 0 to the block mean (the reference multiplies by
    # the boolean mask); fb_pre_zero_channels writes zeros, so NaN / inf in flagged rows do not
    # propagate where `nan * False` would
    spectrum_data_downsampled = _masked_block_mean(lib, device, spectrum_data, good_freq.astype(bool), factor_freq,
                                                   factor_time)
    spectrum_model_downsampled = None
    residuals_downsampled = None
    if spectrum_model is not None:
        spectrum_model_downsampled = _masked_block_mean(lib, device, spectrum_model, None, factor_freq, factor_time)
        spectrum_model_downsampled *= good_freq_downsampled[:, None]
        residuals_downsampled = spectrum_data_downsampled - spectrum_model_downsampled
    freqs_downsampled = manip.downsample_1d(freqs, factor_freq)
    times_downsampled = manip.downsample_1d(times, factor_time)
    return {
        "data": spectrum_data_downsampled,
        "freqs": freqs_downsampled,
        "good_freq": good_freq_downsampled,
        "model": spectrum_model_downsampled,
        "num_freq": len(freqs_downsampled),
        "num_time": len(times_downsampled),
        "res_freq": freqs_downsampled[1] - freqs_downsampled[0],
        "res_time": times_downsampled[1] - times_downsampled[0],
        "residuals": residuals_downsampled,
        "times": times_downsampled,
    }
