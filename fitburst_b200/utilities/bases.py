"""
ReaderBaseClass: drop-in mirror of fitburst.utilities.bases.ReaderBaseClass (reference
utilities/bases.py:20-336) whose heavy methods -- preprocess_data, downsample, window_data -- run
as HBM-streaming CUDA kernels (fitburst_b200/csrc/fb_preprocess.cuh). Format-specific readers
inherit from it exactly like fitburst.backend.generic.DataReader inherits from the reference class
and define load_data().

`data_full` and `data_weights` stay numpy arrays to their users (the pipeline script sums them and
loops over rows, fitburst_pipeline.py:364-372); internally the device copy is the working copy
and the host copy is refreshed lazily on attribute access. window_data() can also hand the
windowed spectrum back as a device tensor (as_tensor=True), which LSFitter accepts directly, so
that the only host->device copy of a burst happens once.

The O(num_freq) bookkeeping (mask logic, dedispersion indices, axis labels) stays on the host in
the reference's own numpy expressions; the O(num_freq x num_time) work is on the device.
"""
import sys

import numpy as np
import torch

from .. import _lib
from ..backend import general
from .._say import say


class _MirroredArray:
    """A (num_freq, num_time) array with a host and a device copy and validity flags. The host
    side may also be a per-channel row pattern (set_rows) that is expanded only where a full array
    is asked for -- the generic format's weights are a channel mask, not num_freq x num_time values."""

    def __init__(self):
        self.host = None
        self.dev = None
        self.host_valid = False
        self.dev_valid = False
        self.rows = None  # (per-channel values, num_time) of a row-constant array not yet expanded

    def set_host(self, array):
        self.host = array
        self.host_valid = array is not None
        self.dev = None
        self.dev_valid = False
        self.rows = None

    def set_rows(self, values, num_time):
        self.set_host(None)
        self.rows = (np.asarray(values, dtype=np.float64), int(num_time))
        self.host_valid = True

    def _expand_rows(self):
        if self.rows is not None:
            values, num_time = self.rows
            self.host = np.repeat(values[:, None], num_time, axis=1)
            self.rows = None

    def get_host(self):
        if self.host_valid:
            self._expand_rows()
            # the caller may modify the array in place: the device copy can no longer be trusted
            self.dev_valid = False
            return self.host
        if self.dev_valid:
            self.host = self.dev.cpu().numpy()
            self.host_valid = True
            self.dev_valid = False
            return self.host
        return None

    def get_dev(self, device):
        if not self.dev_valid:
            if not self.host_valid:
                return None
            if self.rows is not None:
                values, num_time = self.rows
                self.dev = torch.from_numpy(values).to(device)[:, None].expand(len(values), num_time).contiguous()
            else:
                self.dev = upload_spectrum(self.host, device)
            self.dev_valid = True
        return self.dev

    def set_dev(self, tensor):
        self.dev = tensor
        self.dev_valid = True
        self.host = None
        self.host_valid = False
        self.rows = None


def upload_spectrum(array, device, out=None):
    """
    Host spectrum -> float64 device tensor on the current stream. float32 sources cross the host
    link as float32 (half the bytes) and are widened exactly by fb_widen_f32; page-locked sources
    (backend.generic.read_spectrum_pinned, batch.pinned_like) are copied asynchronously.
    """
    source = array if isinstance(array, torch.Tensor) else None
    if source is None:
        array = np.asarray(array)
        if array.dtype not in (np.float32, np.float64):
            array = array.astype(np.float64)
        source = torch.from_numpy(np.ascontiguousarray(array))
    source = source.contiguous()
    if out is None:
        out = torch.empty(source.shape, dtype=torch.float64, device=device)
    if source.dtype == torch.float64:
        out.copy_(source, non_blocking=True)
        return out
    staging = torch.empty(source.shape, dtype=torch.float32, device=device)
    staging.copy_(source, non_blocking=True)
    with torch.cuda.device(device):
        _lib.check(_lib.load_library().fb_widen_f32(staging.data_ptr(), staging.numel(), out.data_ptr(),
                                                    torch.cuda.current_stream(device).cuda_stream))
    return out


class ReaderBaseClass:
    """
    A base class for objects that read and pre-process data provided by user.
    """

    # pylint: disable=too-many-instance-attributes,too-many-arguments,dangerous-default-value

    def __init__(self, device=None):
        """Initializes key attributes to be set by all data readers (utilities/bases.py:27-45)."""
        self.burst_parameters = {}
        self._data_full = _MirroredArray()
        self._data_weights = _MirroredArray()
        self.dedispersion_idx = None
        self.freqs = None
        self.good_freq = None
        self.is_dedispersed = False
        self.num_freq = None
        self.num_time = None
        self.res_time = None
        self.res_freq = None
        self.times = None
        self.times_bin0 = None
        self._device = torch.device(device if device is not None else "cuda")

    data_full = property(lambda self: self._data_full.get_host(),
                         lambda self, value: self._data_full.set_host(value))
    data_weights = property(lambda self: self._data_weights.get_host(),
                            lambda self, value: self._data_weights.set_host(value))

    # ------------------------------------------------------------------ native plumbing
    def _stream(self):
        return torch.cuda.current_stream(self._device).cuda_stream

    def _require_gpu(self):
        if not torch.cuda.is_available():
            raise _lib.FitburstB200Error("no CUDA device visible: fitburst_b200 has no CPU fallback")
        return _lib.load_library()

    def _zero_channels(self, lib, tensor, keep):
        keep_dev = torch.from_numpy(np.ascontiguousarray(keep, dtype=np.uint8)).to(self._device)
        _lib.check(lib.fb_pre_zero_channels(tensor.data_ptr(), keep_dev.data_ptr(), tensor.shape[0],
                                            tensor.shape[1], self._stream()))

    # ------------------------------------------------------------------ reference surface
    def dedisperse(self, dm_value: float, arrival_time: float, ref_freq: float = np.inf,
                   dm_idx: float = general["constants"]["index_dispersion"],
                   dm_offset: float = None) -> None:
        """
        Computes the per-channel time-bin index of the dispersed arrival time
        (utilities/bases.py:47-127); O(num_freq) host bookkeeping.
        """
        dm_const = general["constants"]["dispersion"]

        def delay_of(dm):  # routines/ism.py:50-80
            return dm * dm_const * (self.freqs ** dm_idx - ref_freq ** dm_idx)

        self.dedispersion_idx = np.zeros(self.num_freq, dtype=int)
        if not self.is_dedispersed:
            delay = arrival_time + delay_of(dm_value)
            self.dedispersion_idx[:] = np.ceil((delay - self.times[0]) / (self.res_time))
        elif self.is_dedispersed and dm_offset is not None:
            idx_arrival_time = np.fabs(self.times - arrival_time).argmin()
            idx_1 = np.around((delay_of(dm_value) - self.times[0]) / (self.res_time))
            idx_2 = np.around((delay_of(dm_value + dm_offset) - self.times[0]) / (self.res_time))
            self.dedispersion_idx[:] = idx_arrival_time + (idx_2 - idx_1)

    def downsample(self, factor_freq: int = 1, factor_time: int = 1) -> None:
        """
        Downsamples the spectrum (and weights, axes, mask) by the given factors
        (utilities/bases.py:129-181; block means of routines/manipulate.py:10-72).
        """
        lib = self._require_gpu()
        if self.num_freq % factor_freq or self.num_time % factor_time:
            raise ValueError(f"cannot reshape array of size {self.num_freq * self.num_time} into shape "
                             f"({self.num_freq // factor_freq},{factor_freq},{self.num_time // factor_time},{factor_time})")
        new_nf, new_nt = self.num_freq // factor_freq, self.num_time // factor_time

        def block_mean(mirrored):
            src = mirrored.get_dev(self._device)
            out = torch.empty((new_nf, new_nt), dtype=torch.float64, device=self._device)
            with torch.cuda.device(self._device):
                _lib.check(lib.fb_pre_downsample_2d(src.data_ptr(), self.num_freq, self.num_time, factor_freq,
                                                    factor_time, out.data_ptr(), self._stream()))
            mirrored.set_dev(out)
            return out

        new_data = block_mean(self._data_full)
        self.freqs = np.sum(np.reshape(self.freqs, (new_nf, factor_freq)), axis=1) / factor_freq
        self.times = np.sum(np.reshape(self.times, (new_nt, factor_time)), axis=1) / factor_time
        self.num_freq, self.num_time = new_nf, new_nt
        self.res_freq *= factor_freq
        self.res_time *= factor_time
        if self.good_freq is not None:
            good = np.sum(np.reshape(self.good_freq, (new_nf, factor_freq)), axis=1) / factor_freq
            self.good_freq = good.astype(bool)
            with torch.cuda.device(self._device):
                self._zero_channels(lib, new_data, self.good_freq)
        if self._data_weights.host_valid or self._data_weights.dev_valid:
            block_mean_weights = self._data_weights
            # weights are averaged over the ORIGINAL shape: temporarily restore it for the call
            src = block_mean_weights.get_dev(self._device)
            out = torch.empty((new_nf, new_nt), dtype=torch.float64, device=self._device)
            with torch.cuda.device(self._device):
                _lib.check(lib.fb_pre_downsample_2d(src.data_ptr(), new_nf * factor_freq, new_nt * factor_time,
                                                    factor_freq, factor_time, out.data_ptr(), self._stream()))
            block_mean_weights.set_dev(out)

    def load_data(self) -> None:
        """To be defined by inheriting reader classes (utilities/bases.py:183-199)."""
        sys.exit("ERROR: load_data() must be defined for input data format!")

    def preprocess_data(self, apply_cut_variance: bool = False, apply_cut_skewness: bool = False,
                        normalize_variance: bool = True, remove_baseline: bool = False,
                        skewness_range: list = [-3., 3.], variance_range: list = [0.2, 0.8],
                        variance_weight: float = 1.) -> None:
        """
        Applies pre-fit routines for cleaning the raw dynamic spectrum: RFI masking on per-channel
        variance / skewness, optional baseline removal, normalisation (utilities/bases.py:201-298).
        Sets self.good_freq and zeroes flagged channels of data_full.
        """
        lib = self._require_gpu()
        nf, nt = self.num_freq, self.num_time
        data = self._data_full.get_dev(self._device)
        weights = self._data_weights.get_dev(self._device)
        stream = self._stream()
        with torch.cuda.device(self._device):
            moments = torch.empty((nf, 4), dtype=torch.float64, device=self._device)
            _lib.check(lib.fb_pre_channel_moments(data.data_ptr(), weights.data_ptr(), nf, nt, moments.data_ptr(), stream))
            mom = moments.cpu().numpy()
            mask_freq = mom[:, 0].copy()
            good_freq = mask_freq != 0
            for idx_freq in np.flatnonzero(good_freq & (mom[:, 2] == mom[:, 3])):  # bases.py:255-260
                say(f"ERROR: bad data value of {mom[idx_freq, 2]} in channel {idx_freq}!")
                good_freq[idx_freq] = False
            mean_spectrum = mom[:, 1].copy()
            mean_spectrum[good_freq] /= mask_freq[good_freq]
            if remove_baseline:
                good_dev = torch.from_numpy(good_freq.astype(np.uint8)).to(self._device)
                mean_dev = torch.from_numpy(mean_spectrum).to(self._device)
                _lib.check(lib.fb_pre_remove_baseline(data.data_ptr(), weights.data_ptr(), good_dev.data_ptr(),
                                                      mean_dev.data_ptr(), nf, nt, stream))
            sums = torch.empty((nf, 2), dtype=torch.float64, device=self._device)
            _lib.check(lib.fb_pre_power_sums(data.data_ptr(), nf, nt, sums.data_ptr(), stream))
            sums_host = sums.cpu().numpy()
        variance = sums_host[:, 0].copy()
        variance[good_freq] /= mask_freq[good_freq]
        skewness = sums_host[:, 1].copy()
        skewness[good_freq] /= mask_freq[good_freq]
        skewness_mean = np.mean(skewness[good_freq])
        skewness_std = np.std(skewness[good_freq])
        if normalize_variance:
            variance[good_freq] /= np.max(variance[good_freq])
        if apply_cut_variance:
            good_freq = np.logical_and(good_freq, (variance / variance_weight < variance_range[1]))
            good_freq = np.logical_and(good_freq, (variance / variance_weight > variance_range[0]))
        if apply_cut_skewness:
            good_freq = np.logical_and(good_freq, (skewness < skewness_mean + skewness_range[1] * skewness_std))
            good_freq = np.logical_and(good_freq, (skewness > skewness_mean + skewness_range[0] * skewness_std))
        with torch.cuda.device(self._device):
            self._zero_channels(lib, data, good_freq)
        self._data_full.set_dev(data)
        self.good_freq = good_freq
        num_freq_bad = self.num_freq - np.sum(self.good_freq)
        say(f"INFO: flagged and removed {num_freq_bad} out of {self.num_freq} channels!")

    def window_data(self, arrival_time: float, window: float = 0.08, as_tensor: bool = False) -> tuple:
        """
        Returns the subset of data centred on the (de-dispersed) arrival time of every channel
        (utilities/bases.py:300-336). With as_tensor=True the windowed spectrum is returned as a
        device tensor instead of a numpy array.
        """
        lib = self._require_gpu()
        idx_arrival_time = np.fabs(self.times - arrival_time).argmin()
        num_window_bins = np.around(window / self.res_time).astype(int)
        width = int(num_window_bins) * 2
        start = np.asarray(self.dedispersion_idx, dtype=np.int64) - int(num_window_bins)
        if np.any(start < 0) or np.any(start + width > self.num_time) or width < 1:
            # numpy slicing would wrap / truncate here and the reference's row assignment raises
            raise ValueError(f"could not broadcast input array into shape ({width},): window of +/-"
                             f"{int(num_window_bins)} bins leaves the spectrum for at least one channel")
        data = self._data_full.get_dev(self._device)
        start_dev = torch.from_numpy(start.astype(np.int32)).to(self._device)
        out = torch.empty((self.num_freq, width), dtype=torch.float64, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(lib.fb_pre_window(data.data_ptr(), self.num_freq, self.num_time, start_dev.data_ptr(), width,
                                         out.data_ptr(), self._stream()))
        times_windowed = self.times[idx_arrival_time - num_window_bins: idx_arrival_time + num_window_bins]
        return (out if as_tensor else out.cpu().numpy()), times_windowed
