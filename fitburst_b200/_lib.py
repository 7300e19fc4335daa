"""
ctypes binding of the C ABI declared in include/fitburst_b200.h.

The shared library is built in-tree (fitburst_b200/libfitburst_b200.so) by
`__graft_entry__.build()` / `python -m fitburst_b200.build`. There is NO CPU fallback: if the
library is missing or no CUDA device is visible, every product entry point raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfitburst_b200.so")

FB_MAX_COMPONENTS = 16
FB_NUM_PARAMETERS = 10
FB_MAX_FREE = 67

PARAMETER_INDEX = {
    "amplitude": 0,
    "arrival_time": 1,
    "burst_width": 2,
    "dm": 3,
    "dm_index": 4,
    "scattering_timescale": 5,
    "scattering_index": 6,
    "spectral_index": 7,
    "spectral_running": 8,
    "ref_freq": 9,
}


class FitburstB200Error(RuntimeError):
    """Raised when the native library reports a failure."""


class PlanDesc(ctypes.Structure):
    _fields_ = [
        ("num_freq", ctypes.c_int32),
        ("num_time", ctypes.c_int32),
        ("num_components", ctypes.c_int32),
        ("factor_freq_upsample", ctypes.c_int32),
        ("factor_time_upsample", ctypes.c_int32),
        ("is_folded", ctypes.c_int32),
        ("scintillation", ctypes.c_int32),
        ("normalize_pbf", ctypes.c_int32),
        ("dm_incoherent", ctypes.c_double),
        ("dispersion_constant", ctypes.c_double),
    ]


class FitResult(ctypes.Structure):
    _fields_ = [
        ("status", ctypes.c_int32),
        ("nfev", ctypes.c_int32),
        ("njev", ctypes.c_int32),
        ("num_free", ctypes.c_int32),
        ("cost", ctypes.c_double),
        ("optimality", ctypes.c_double),
        ("chisq_initial", ctypes.c_double),
        ("x", ctypes.c_double * FB_MAX_FREE),
        ("grad", ctypes.c_double * FB_MAX_FREE),
        ("uncertainty", ctypes.c_double * FB_MAX_FREE),
    ]


class FitOptions(ctypes.Structure):
    _fields_ = [
        ("ftol", ctypes.c_double),
        ("xtol", ctypes.c_double),
        ("gtol", ctypes.c_double),
        ("max_nfev", ctypes.c_int32),
        ("compute_hessian", ctypes.c_int32),
    ]


_P = ctypes.c_void_p
_D = ctypes.POINTER(ctypes.c_double)
_I32 = ctypes.c_int32

# name -> (restype, argtypes); must list every symbol declared in include/fitburst_b200.h
SIGNATURES = {
    "fb_last_error": (ctypes.c_char_p, []),
    "fb_version": (ctypes.c_int, []),
    "fb_plan_create": (ctypes.c_int, [ctypes.POINTER(PlanDesc), _D, _D, ctypes.POINTER(_P)]),
    "fb_plan_destroy": (ctypes.c_int, [_P]),
    "fb_set_parameters": (ctypes.c_int, [_P, _D, _P]),
    "fb_get_parameters": (ctypes.c_int, [_P, _D, _P]),
    "fb_model_eval": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "fb_profile_eval": (ctypes.c_int, [_P, _P, _I32, _I32, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_double, ctypes.c_double, _I32, _I32, _I32, _I32, _P, _P]),
    "fb_set_weights": (ctypes.c_int, [_P, _P, ctypes.POINTER(ctypes.c_uint8), _I32, _I32, _I32, _D, _D, _P]),
    "fb_load_weights": (ctypes.c_int, [_P, _D, _P]),
    "fb_set_free_parameters": (ctypes.c_int, [_P, ctypes.POINTER(_I32), ctypes.POINTER(_I32)]),
    "fb_normal_equations": (ctypes.c_int, [_P, _P, _D, _P]),
    "fb_normal_equations_dev": (ctypes.c_int, [_P, _P, _P, _P]),
    "fb_residuals": (ctypes.c_int, [_P, _P, _P, _P]),
    "fb_jacobian_dense": (ctypes.c_int, [_P, _P, _P, _P]),
    "fb_derivative_map": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "fb_hessian": (ctypes.c_int, [_P, _P, _D, _P]),
    "fb_derivative2_map": (ctypes.c_int, [_P, _P, _I32, _I32, _I32, _P, _P]),
    "fb_fit": (ctypes.c_int, [_P, _P, ctypes.POINTER(FitOptions), ctypes.POINTER(FitResult), _D, _D, _P]),
    "fb_trf_create": (ctypes.c_int, [_I32, ctypes.c_int64, _D, ctypes.POINTER(FitOptions), ctypes.POINTER(_P)]),
    "fb_trf_destroy": (ctypes.c_int, [_P]),
    "fb_trf_start": (ctypes.c_int, [_P, _D]),
    "fb_trf_propose": (ctypes.c_int, [_P, _D]),
    "fb_trf_update": (ctypes.c_int, [_P, _D]),
    "fb_trf_result": (ctypes.c_int, [_P, ctypes.POINTER(FitResult), _D]),
    "fb_fit_device_begin": (ctypes.c_int, [_P, ctypes.POINTER(FitOptions), ctypes.c_int64, _P]),
    "fb_fit_device_evaluate": (ctypes.c_int, [_P, _P, _P, _P]),
    "fb_fit_device_step": (ctypes.c_int, [_P, _P, _P]),
    "fb_fit_device_poll": (ctypes.c_int, [_P, ctypes.POINTER(_I32), ctypes.POINTER(_I32), _P]),
    "fb_fit_device_end": (ctypes.c_int, [_P, ctypes.POINTER(FitResult), _D, ctypes.POINTER(_I32), _P]),
    "fb_fit_batched": (ctypes.c_int, [_P, _I32, _P, _P, _P, ctypes.POINTER(FitOptions), ctypes.POINTER(FitResult), _P]),
    "fb_chisq_grid": (ctypes.c_int, [_P, _P, _D, _I32, _D, _I32, _P, _P]),
    "fb_pre_channel_moments": (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "fb_pre_remove_baseline": (ctypes.c_int, [_P, _P, _P, _P, _I32, _I32, _P]),
    "fb_pre_power_sums": (ctypes.c_int, [_P, _I32, _I32, _P, _P]),
    "fb_pre_zero_channels": (ctypes.c_int, [_P, _P, _I32, _I32, _P]),
    "fb_pre_downsample_2d": (ctypes.c_int, [_P, _I32, _I32, _I32, _I32, _P, _P]),
    "fb_pre_window": (ctypes.c_int, [_P, _I32, _I32, _P, _I32, _P, _P]),
    "fb_pre_time_series": (ctypes.c_int, [_P, _I32, _I32, _P, _P, _P]),
    "fb_upload_rows": (ctypes.c_int, [_P, _I32, _P, _I32, _I32, _I32, _P, _P]),
    "fb_spectrum_rpl": (ctypes.c_int, [_P, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double, _P, _P]),
    "fb_ism_times": (ctypes.c_int, [_I32, _P, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_double, _P, _P]),
    "fb_amplitude_per_channel": (ctypes.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "fb_widen_f32": (ctypes.c_int, [_P, ctypes.c_int64, _P, _P]),
    "fb_measure_fp64_peak": (ctypes.c_int, [_D, _P]),
    "fb_launch_count": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int64), _I32]),
}

_lib = None


def load_library():
    """Loads libfitburst_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FITBURST_B200_LIB", LIB_PATH)  # developer switch: A/B of kernel build variants
    if path != LIB_PATH and os.path.exists(path):
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib
    if not os.path.exists(LIB_PATH):
        raise FitburstB200Error(
            f"{LIB_PATH} is missing: build it with `python -m fitburst_b200.build` "
            "(fitburst_b200 has no CPU fallback)"
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        message = load_library().fb_last_error().decode("utf-8", "replace")
        if code == -4:
            raise np.linalg.LinAlgError(message)  # same exception type as routines/ism.py:46
        raise FitburstB200Error(f"fitburst_b200 error {code}: {message}")


def dptr(array: np.ndarray):
    """Host double pointer of a C-contiguous float64 numpy array."""
    assert array.dtype == np.float64 and array.flags["C_CONTIGUOUS"]
    return array.ctypes.data_as(_D)
