"""
SpectrumModeler: drop-in mirror of fitburst.analysis.model.SpectrumModeler
(reference analysis/model.py:15-409) whose arithmetic runs in the sm_100a kernels behind
include/fitburst_b200.h.

Same constructor, attributes and methods as the reference class. Differences, all deliberate:
 * the four (Nf, Nt, Nc) per-component caches (analysis/model.py:110-124) are materialised
   lazily, on first access after a compute_model() call, instead of on every call (they are
   1.6 GB at 16384 x 1024 x 3 and the fit never needs them on the host);
 * with verbose=False the reference prints one blank line per component per call
   (analysis/model.py:259-264); this class prints nothing;
 * a missing CUDA device / native library raises instead of computing on the CPU.
"""
import sys

import numpy as np
import torch

from .. import _lib
from ..backend import general

_CACHE_NAMES = ("amplitude", "spectrum", "timediff", "timeprof")


class SpectrumModeler:
    """
    A Python structure that contains all information regarding parameters that
    describe and are used to compute models of dynamic spectra.
    """

    # pylint: disable=too-many-instance-attributes

    def __init__(self, freqs: float, times: float, dm_incoherent: float = 0.,
        factor_freq_upsample: int = 1, factor_time_upsample: int = 1, num_components: int = 1,
        is_dedispersed: bool = False, is_folded: bool = False, scintillation: bool = False,
        verbose: bool = False, device=None) -> None:
        """Same arguments as the reference constructor (analysis/model.py:23-26) plus `device`."""
        # pylint: disable=too-many-arguments
        self.dm_incoherent = dm_incoherent
        self.factor_freq_upsample = factor_freq_upsample
        self.factor_time_upsample = factor_time_upsample
        self.freqs = freqs
        self.is_dedispersed = is_dedispersed
        self.is_folded = is_folded
        self.num_components = num_components
        self.scintillation = scintillation
        self.times = times
        self.verbose = verbose

        self.num_freq = len(self.freqs)
        self.num_time = len(self.times)
        self.res_freq = np.fabs(self.freqs[1] - self.freqs[0])
        self.res_time = np.fabs(self.times[1] - self.times[0])

        # NOTE: 'ref_freq' is not listed here as it is always held fixed (analysis/model.py:90-102).
        self.parameters = [
            "amplitude",
            "arrival_time",
            "burst_width",
            "dm",
            "dm_index",
            "scattering_timescale",
            "scattering_index",
            "spectral_index",
            "spectral_running",
        ]
        for current_parameter in self.parameters:
            setattr(self, current_parameter, None)

        # native state
        self._device = torch.device(device if device is not None else "cuda")
        self._plan = None
        self._plan_key = None
        self._cache_host = {}
        self._cache_state = None  # (packed parameters, data on device) of the last compute_model
        self._model_dev = None

    # ------------------------------------------------------------------ native plumbing
    def _stream(self):
        return torch.cuda.current_stream(self._device).cuda_stream

    def _normalize_pbf(self) -> bool:
        return bool(general["flags"]["normalize_pbf"])

    def _ensure_plan(self):
        """(Re)creates the native plan when the configuration it was built for changed."""
        key = (
            self.num_freq, self.num_time, int(self.num_components), int(self.factor_freq_upsample),
            int(self.factor_time_upsample), bool(self.is_folded), bool(self.scintillation),
            self._normalize_pbf(), float(self.dm_incoherent), float(general["constants"]["dispersion"]),
        )
        if self._plan is not None and key == self._plan_key:
            return self._plan
        if not torch.cuda.is_available():
            raise _lib.FitburstB200Error("no CUDA device visible: fitburst_b200 has no CPU fallback")
        lib = _lib.load_library()
        self._destroy_plan()
        desc = _lib.PlanDesc(
            self.num_freq, self.num_time, int(self.num_components), int(self.factor_freq_upsample),
            int(self.factor_time_upsample), int(bool(self.is_folded)), int(bool(self.scintillation)),
            int(self._normalize_pbf()), float(self.dm_incoherent), float(general["constants"]["dispersion"]),
        )
        freqs = np.ascontiguousarray(self.freqs, dtype=np.float64)
        times = np.ascontiguousarray(self.times, dtype=np.float64)
        handle = _lib._P()
        with torch.cuda.device(self._device):
            _lib.check(lib.fb_plan_create(desc, _lib.dptr(freqs), _lib.dptr(times), handle))
        self._plan = handle
        self._plan_key = key
        return handle

    def _destroy_plan(self):
        if getattr(self, "_plan", None) is not None:
            try:
                _lib.load_library().fb_plan_destroy(self._plan)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._plan = None

    def __del__(self):
        self._destroy_plan()

    def _packed_parameters(self) -> np.ndarray:
        """Packs the ten parameter lists into the (10, Nc) block of include/fitburst_b200.h."""
        num_components = int(self.num_components)
        packed = np.zeros((_lib.FB_NUM_PARAMETERS, num_components), dtype=np.float64)
        for name, row in _lib.PARAMETER_INDEX.items():
            values = getattr(self, name)
            values = np.atleast_1d(np.asarray(values, dtype=np.float64))
            if name in ("dm", "dm_index", "scattering_timescale", "scattering_index"):
                packed[row, :] = values[0]  # element 0 is used for all components (model.py:155-159)
            else:
                packed[row, :] = values[:num_components]
        return packed

    def _push_parameters(self, packed=None):
        plan = self._ensure_plan()
        if packed is None:
            packed = self._packed_parameters()
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load_library().fb_set_parameters(plan, _lib.dptr(packed), self._stream()))
        return packed

    def _to_device(self, array):
        if array is None:
            return None
        if isinstance(array, torch.Tensor):
            return array.to(device=self._device, dtype=torch.float64).contiguous()
        if isinstance(array, np.ndarray) and array.dtype == np.float32 and array.ndim == 2:
            # float32 spectra cross the host link as float32 and are widened exactly on the device
            from ..utilities.bases import upload_spectrum
            return upload_spectrum(array, self._device)
        return torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64)).to(self._device)

    def compute_model_device(self, data_dev=None, packed=None) -> torch.Tensor:
        """compute_model() that keeps the (Nf, Nt) result on the device (torch.float64)."""
        plan = self._ensure_plan()
        packed = self._push_parameters(packed)
        if self._model_dev is None or self._model_dev.shape != (self.num_freq, self.num_time):
            self._model_dev = torch.empty((self.num_freq, self.num_time), dtype=torch.float64, device=self._device)
        data_ptr = data_dev.data_ptr() if data_dev is not None else None
        with torch.cuda.device(self._device):
            _lib.check(_lib.load_library().fb_model_eval(
                plan, data_ptr, self._model_dev.data_ptr(), None, None, None, None, self._stream()))
        self._cache_state = (packed.copy(), data_dev)
        self._cache_host = {}
        return self._model_dev

    # ------------------------------------------------------------------ reference surface
    def compute_model(self, data: float = None) -> float:
        """
        Computes the model dynamic spectrum based on model parameters (set as class
        attributes) and input values of times and frequencies (analysis/model.py:126-279).
        """
        if self.scintillation and data is None:
            sys.exit("ERROR: scintillation modelling is desired by data are missing!")
        if self.verbose:
            self._print_parameters()
        data_dev = self._to_device(data) if self.scintillation else None
        model_dev = self.compute_model_device(data_dev)
        return model_dev.cpu().numpy()

    def _print_parameters(self):
        """The per-component line of analysis/model.py:164-172,259-261."""
        for comp in range(int(self.num_components)):
            dm, amp, toa = self.dm[0], self.amplitude[comp], self.arrival_time[comp]
            sc_idx, sc_time, width = self.scattering_index[0], self.scattering_timescale[0], self.burst_width[comp]
            if self.scintillation:
                print(f"{dm:.5f} {toa:.5f} ", f"{sc_idx:.5f}  {sc_time:.5f}  {width:.5f}", end=" ")
                print()
            else:
                print(f"{dm:.5f}  {amp:.5f}  {toa:.5f}  ", f"{sc_idx:.5f}  {sc_time:.5f}  {width:.5f}", end=" ")
                print(f"{self.spectral_index[comp]:.5f}  {self.spectral_running[comp]:.5f}")

    def _materialise_caches(self):
        """Fills the four per-component caches for the state of the last compute_model() call."""
        shape = (self.num_freq, self.num_time, int(self.num_components))
        if self._cache_state is None:
            for name in _CACHE_NAMES:
                self._cache_host[name] = np.zeros(shape, dtype=float)  # analysis/model.py:110-124
            return
        packed, data_dev = self._cache_state
        plan = self._ensure_plan()
        current = self._packed_parameters() if self.amplitude is not None else packed
        self._push_parameters(packed)
        buffers = [torch.empty(shape, dtype=torch.float64, device=self._device) for _ in _CACHE_NAMES]
        scratch = torch.empty((self.num_freq, self.num_time), dtype=torch.float64, device=self._device)
        with torch.cuda.device(self._device):
            _lib.check(_lib.load_library().fb_model_eval(
                plan, data_dev.data_ptr() if data_dev is not None else None, scratch.data_ptr(),
                buffers[0].data_ptr(), buffers[1].data_ptr(), buffers[2].data_ptr(), buffers[3].data_ptr(),
                self._stream()))
        for name, buf in zip(_CACHE_NAMES, buffers):
            self._cache_host[name] = buf.cpu().numpy()
        self._push_parameters(current)

    def _get_cache(self, name):
        if name not in self._cache_host:
            self._materialise_caches()
        return self._cache_host[name]

    amplitude_per_component = property(
        lambda self: self._get_cache("amplitude"),
        lambda self, value: self._cache_host.__setitem__("amplitude", value))
    spectrum_per_component = property(
        lambda self: self._get_cache("spectrum"),
        lambda self, value: self._cache_host.__setitem__("spectrum", value))
    timediff_per_component = property(
        lambda self: self._get_cache("timediff"),
        lambda self, value: self._cache_host.__setitem__("timediff", value))
    timeprof_per_component = property(
        lambda self: self._get_cache("timeprof"),
        lambda self, value: self._cache_host.__setitem__("timeprof", value))

    def compute_profile(self, times: float, arrival_time: float, sc_time_ref: float, sc_index: float,
        width: float, freqs: float, ref_freq: float, is_folded: bool = False) -> float:
        """
        Returns the temporal profile, depending on input values of width and scattering timescale
        (analysis/model.py:281-352): pulse-broadening function with the -5 width clamp when any
        scattering timescale is positive, Gaussian otherwise; optional folding.
        """
        # pylint: disable=too-many-arguments
        from ..routines.profile import _evaluate
        freqs = np.asarray(freqs, dtype=np.float64)
        with np.errstate(all="ignore"):
            sc_time = sc_time_ref * (freqs / ref_freq) ** sc_index
        scattered = bool(np.any(sc_time > 0.0))
        return _evaluate(times, arrival_time, width, freqs, ref_freq, sc_time_ref, sc_index, scattered, True,
                         bool(is_folded), self._normalize_pbf())

    def get_parameters_dict(self) -> dict:
        """
        Returns model parameters as a dictionary, with keys set to the parameter names
        and values set to the Python list containing parameter values (analysis/model.py:354-381).
        """
        parameter_dict = {}
        for current_parameter in self.parameters:
            parameter_dict[current_parameter] = getattr(self, current_parameter)
        parameter_dict["ref_freq"] = getattr(self, "ref_freq")
        return parameter_dict

    def update_parameters(self, model_parameters: dict, global_parameters: list = ["dm", "scattering_timescale"]) -> None:
        """
        Overloads parameter values stored in object with those supplied by the user
        (analysis/model.py:383-409).
        """
        # pylint: disable=dangerous-default-value
        for current_parameter in model_parameters.keys():
            setattr(self, current_parameter, model_parameters[current_parameter])
            if len(model_parameters[current_parameter]) > 1:
                num_components = len(model_parameters[current_parameter])
                setattr(self, "num_components", num_components)
                if current_parameter in global_parameters:
                    setattr(self, current_parameter, [model_parameters[current_parameter][0]] * num_components)
