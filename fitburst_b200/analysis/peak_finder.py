"""
FindPeak: mirror of fitburst.analysis.peak_finder.FindPeak (reference
analysis/peak_finder.py:39-203) -- initial guesses for the burst components from the band-averaged
time series. The O(num_freq x num_time) band average runs on the GPU (fb_pre_time_series) when the
spectrum is a device tensor (what window_data(as_tensor=True) returns); the O(num_time) peak scan
is scipy.signal.find_peaks on the host, exactly like the reference. The reference's plotting / csv
method is not mirrored (it needs matplotlib and an attribute that is never set).
"""
import numpy as np
import torch
from scipy.signal import find_peaks

from .. import _lib
from .._say import say


class FindPeak:
    """A Python class that finds peaks in the loaded data."""

    def __init__(self, data: float, time: float, freq: float, rms: float = None):
        """Same arguments as the reference constructor (analysis/peak_finder.py:47-71)."""
        self.data = data
        self.freq = freq
        self.time = np.asarray(time)
        self.rms = rms

    def _mean_over_channels(self):
        if isinstance(self.data, torch.Tensor) and self.data.is_cuda:
            lib = _lib.load_library()
            data = self.data.to(dtype=torch.float64).contiguous()
            num_freq, num_time = data.shape
            scratch = torch.empty(((num_freq + 63) // 64, num_time), dtype=torch.float64, device=data.device)
            out = torch.empty(num_time, dtype=torch.float64, device=data.device)
            with torch.cuda.device(data.device):
                _lib.check(lib.fb_pre_time_series(data.data_ptr(), num_freq, num_time, scratch.data_ptr(),
                                                  out.data_ptr(), torch.cuda.current_stream(data.device).cuda_stream))
            return out.cpu().numpy()
        return np.asarray(self.data).mean(axis=0)

    def find_peak(self, distance: int = 5):
        """
        Finds peak positions in the band-averaged time series and the approximate temporal width
        of each peak (analysis/peak_finder.py:72-130).
        """
        self.mean_intensity_freq = self._mean_over_channels()
        peaks_location = find_peaks(self.mean_intensity_freq, distance=distance)

        self.peak_times = self.time[peaks_location[0]]
        self.peak_mean_intensities = self.mean_intensity_freq[peaks_location[0]]

        self.rms_intensity = np.sqrt(np.mean((self.mean_intensity_freq) ** 2))
        say(f"The rms intensity is {self.rms_intensity} \n")

        if self.rms is not None:
            shifted_rms_intensity = self.rms * np.max(self.mean_intensity_freq)
            index_peaks_greater_rms = np.where(self.peak_mean_intensities > shifted_rms_intensity)
        else:
            index_peaks_greater_rms = np.where(self.peak_mean_intensities > self.rms_intensity)
        self.peaks_greater_rms = self.peak_mean_intensities[index_peaks_greater_rms]
        self.times_peaks_greater_rms = self.peak_times[index_peaks_greater_rms]

        # width of each burst; like the reference, only the value of the LAST accepted peak
        # survives the loop (analysis/peak_finder.py:105-124)
        for each_peak in index_peaks_greater_rms[0]:
            below_peak_index, above_peak_index = each_peak - 1, each_peak + 1
            if below_peak_index < 0:
                below_peak_index = 0
            if above_peak_index >= len(self.peak_times):
                above_peak_index = each_peak
            below_peak_time = self.peak_times[below_peak_index] * 1000
            above_peak_time = self.peak_times[above_peak_index] * 1000
            self.burst_widths = above_peak_time - below_peak_time
            self.time_of_arrivals = self.times_peaks_greater_rms * 1000

        say("   Times of Arrivals  Burst_Widths")
        for idx, toa in enumerate(np.atleast_1d(self.time_of_arrivals)):
            say(f"{idx}  {toa}  {self.burst_widths}")

    def get_parameters_dict(self, original_dict: dict, update_width=False):
        """
        Returns peak values (and optionally widths) in a dictionary that is compatible with
        fitburst (analysis/peak_finder.py:170-203).
        """
        mul_factor = len(self.time_of_arrivals)
        burst_parameters = {}
        for current_key in original_dict.keys():
            if current_key == "arrival_time":
                burst_parameters[current_key] = (self.time_of_arrivals / 1000.).tolist()
            elif current_key == "burst_width" and update_width:
                burst_parameters[current_key] = (self.burst_widths / 1000.).tolist()
            else:
                burst_parameters[current_key] = [original_dict[current_key][0]] * mul_factor
        return burst_parameters
