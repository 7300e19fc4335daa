"""
DataReader: drop-in mirror of fitburst.backend.generic.DataReader (reference
backend/generic.py:10-99) on top of the device ReaderBaseClass of this package: the ".npz generic
format" (docs/usage/formatting_data_generic.md: `data_full`, `metadata`, `burst_parameters`) is
parsed on the host exactly like the reference does, the spectrum then lives on the GPU for
preprocess_data / downsample / window_data and is handed to LSFitter without coming back.
"""
import os

import numpy as np

from ..utilities import bases


class DataReader(bases.ReaderBaseClass):
    """
    A child class of I/O and processing for generic data stored in a .npz file, inheriting the
    basic structure defined in ReaderBaseClass().
    """

    def __init__(self, fname, device=None):
        super().__init__(device=device)
        self.file_path = fname
        if not os.path.isfile(self.file_path):
            raise IOError(f"Data file not found: {self.file_path}")

    def load_data(self):
        """
        Loads `data_full`, `metadata` and `burst_parameters` from the .npz file and derives the
        axis labels and the weights array (backend/generic.py:28-99, same checks and messages).
        """
        unpacked_data_set = np.load(self.file_path, allow_pickle=True)

        expected_subfile_names = ["data_full", "metadata", "burst_parameters"]
        retrieved_subfile_names = unpacked_data_set.files
        if not all(f in retrieved_subfile_names for f in expected_subfile_names):
            raise AssertionError(
                f"Data file does not contain one of more of the following keys: "
                f"{expected_subfile_names}"
            )

        metadata = unpacked_data_set["metadata"].item()
        burst_parameters = unpacked_data_set["burst_parameters"].item()
        # each parameter has its values in a list (one entry per component), generic.py:53-57
        for key, value in burst_parameters.items():
            self.burst_parameters[key] = value if isinstance(value, list) else [value]

        self.data_full = unpacked_data_set["data_full"]
        self.num_freq, self.num_time = self.data_full.shape
        if self.num_freq != metadata["num_freq"]:
            raise AssertionError(
                "Data shape does not match recorded number of channels"
                f"({self.num_freq} != {metadata['num_freq']})"
            )
        if self.num_time != metadata["num_time"]:
            raise AssertionError(
                "Data shape does not match recorded number of time samples"
                f"({self.num_time} != {metadata['num_time']})"
            )

        # weights: 1 = usable, 0 = masked channel (generic.py:75-78)
        data_weights = np.ones((self.num_freq, self.num_time), dtype=float)
        data_weights[metadata["bad_chans"], :] = 0.
        self.data_weights = data_weights

        # time labels in seconds since the start of the spectrum (generic.py:80-86)
        self.res_time = metadata["res_time"]
        self.times = np.arange(self.num_time, dtype=np.float64) * self.res_time
        self.times_bin0 = metadata["times_bin0"]

        # channel-centre labels: leading edge + half a channel (generic.py:88-96)
        self.res_freq = metadata["res_freq"]
        freqs = np.arange(self.num_freq, dtype=np.float64) * self.res_freq
        freqs += metadata["freqs_bin0"]
        freqs += self.res_freq / 2.0
        self.freqs = freqs

        self.is_dedispersed = metadata["is_dedispersed"]
