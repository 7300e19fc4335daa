"""
Mirror of fitburst.routines.derivative (reference routines/derivative.py:454-2836): the nine
first-derivative maps deriv_model_wrt_<name>(parameters, model, component=0[, add_all=True]) and
the 45 mixed partials deriv2_model_wrt_<a>_<b>(parameters, model, component=0), each returned as
an (Nf, Nt) numpy array, evaluated by the CUDA kernels behind fb_derivative_map /
fb_derivative2_map.

One deliberate difference: the reference functions combine the `parameters` dictionary with
whatever per-component caches the last compute_model() call left in `model`; here the maps are
evaluated consistently at `parameters` (the two coincide in every in-tree call site, where
`parameters` is model.get_parameters_dict()).
"""
import numpy as np
import torch

from .. import _lib

_NAMES = [
    "amplitude", "arrival_time", "burst_width", "dm", "dm_index",
    "scattering_timescale", "scattering_index", "spectral_index", "spectral_running",
]
_GLOBAL_CAPABLE = ("dm", "dm_index", "scattering_timescale", "scattering_index")

# order in which the reference names its 45 second derivatives (routines/derivative.py:938-2836)
_D2_ORDER = [
    "amplitude", "spectral_running", "spectral_index", "burst_width", "arrival_time",
    "dm", "dm_index", "scattering_timescale", "scattering_index",
]


def _packed_from_dict(parameters: dict, model) -> np.ndarray:
    num_components = int(model.num_components)
    packed = np.zeros((_lib.FB_NUM_PARAMETERS, num_components), dtype=np.float64)
    for name, row in _lib.PARAMETER_INDEX.items():
        values = np.atleast_1d(np.asarray(parameters[name], dtype=np.float64))
        if name in _GLOBAL_CAPABLE:
            packed[row, :] = values[0]
        else:
            packed[row, :] = values[:num_components]
    return packed


def _data_pointer(model):
    state = getattr(model, "_cache_state", None)
    if model.scintillation:
        if state is None or state[1] is None:
            raise _lib.FitburstB200Error("scintillation mode: call compute_model(data) first")
        return state[1].data_ptr()
    return None


def _first(name, parameters, model, component, add_all):
    plan = model._ensure_plan()
    current = model._packed_parameters()
    model._push_parameters(_packed_from_dict(parameters, model))
    out = torch.empty((model.num_freq, model.num_time), dtype=torch.float64, device=model._device)
    with torch.cuda.device(model._device):
        _lib.check(_lib.load_library().fb_derivative_map(
            plan, _data_pointer(model), _lib.PARAMETER_INDEX[name], int(component), int(bool(add_all)),
            out.data_ptr(), model._stream()))
    model._push_parameters(current)
    return out.cpu().numpy()


def _second(name1, name2, parameters, model, component):
    plan = model._ensure_plan()
    current = model._packed_parameters()
    model._push_parameters(_packed_from_dict(parameters, model))
    out = torch.empty((model.num_freq, model.num_time), dtype=torch.float64, device=model._device)
    with torch.cuda.device(model._device):
        _lib.check(_lib.load_library().fb_derivative2_map(
            plan, _data_pointer(model), _lib.PARAMETER_INDEX[name1], _lib.PARAMETER_INDEX[name2],
            int(component), out.data_ptr(), model._stream()))
    model._push_parameters(current)
    return out.cpu().numpy()


def _make_first(name):
    if name in _GLOBAL_CAPABLE:
        def fn(parameters: dict, model: object, component: int = 0, add_all: bool = True) -> float:
            return _first(name, parameters, model, component, add_all)
    else:
        def fn(parameters: dict, model: object, component: int = 0) -> float:
            return _first(name, parameters, model, component, False)
    fn.__name__ = f"deriv_model_wrt_{name}"
    fn.__doc__ = f"d model / d {name} as an (Nf, Nt) map (reference routines/derivative.py)."
    return fn


def _make_second(name1, name2):
    def fn(parameters: dict, model: object, component: int = 0) -> float:
        return _second(name1, name2, parameters, model, component)
    fn.__name__ = f"deriv2_model_wrt_{name1}_{name2}"
    fn.__doc__ = f"d2 model / d {name1} d {name2} as an (Nf, Nt) map (reference routines/derivative.py)."
    return fn


for _name in _NAMES:
    globals()[f"deriv_model_wrt_{_name}"] = _make_first(_name)

# the reference defines each unordered pair once, named in the order below
# (LSFitter.compute_hessian tries both orders, analysis/fitter.py:146-150).
_D2_PAIRS = [
    ("amplitude", ["amplitude", "spectral_running", "spectral_index", "burst_width", "arrival_time",
                   "dm", "dm_index", "scattering_timescale", "scattering_index"]),
    ("spectral_running", ["spectral_running", "spectral_index", "burst_width", "arrival_time", "dm",
                          "dm_index", "scattering_timescale", "scattering_index"]),
    ("spectral_index", ["spectral_index", "burst_width", "arrival_time", "dm", "dm_index",
                        "scattering_timescale", "scattering_index"]),
    ("burst_width", ["burst_width", "arrival_time", "scattering_timescale", "scattering_index", "dm",
                     "dm_index"]),
    ("arrival_time", ["arrival_time", "dm", "dm_index", "scattering_timescale", "scattering_index"]),
    ("dm", ["dm_index", "dm"]),
    ("scattering_timescale", ["scattering_timescale", "scattering_index", "dm", "dm_index"]),
    ("scattering_index", ["scattering_index", "dm", "dm_index"]),
    ("dm_index", ["dm_index"]),
]
for _first_name, _seconds in _D2_PAIRS:
    for _second_name in _seconds:
        globals()[f"deriv2_model_wrt_{_first_name}_{_second_name}"] = _make_second(_first_name, _second_name)
