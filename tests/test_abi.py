"""The C-ABI library builds for sm_100a, loads without a GPU and exports every declared symbol."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fitburst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(native_lib):
    from fitburst_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(native_lib, name), f"{name} declared in include/fitburst_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert native_lib.fb_version() >= 100


def test_no_cpu_fallback_in_product_package():
    """Nothing under fitburst_b200/ may import the oracle (it is test infrastructure)."""
    for root, _, files in os.walk(os.path.join(ROOT, "fitburst_b200")):
        for name in files:
            if name.endswith(".py"):
                text = open(os.path.join(root, name)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(root, name)


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "fitburst_b200", "libfitburst_b200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_parameter_packing_order_matches_reference():
    from fitburst_b200 import _lib
    from fitburst_b200.analysis.model import SpectrumModeler
    import numpy as np
    model = SpectrumModeler.__new__(SpectrumModeler)
    order = ["amplitude", "arrival_time", "burst_width", "dm", "dm_index", "scattering_timescale",
             "scattering_index", "spectral_index", "spectral_running"]  # analysis/model.py:92-102
    assert [k for k, _ in sorted(_lib.PARAMETER_INDEX.items(), key=lambda kv: kv[1])][:9] == order


def test_argument_validation_without_gpu(native_lib):
    """Bad arguments are rejected with FB_ERR_INVALID before any CUDA call is made."""
    import ctypes
    import numpy as np
    from fitburst_b200 import _lib
    freqs = np.linspace(400.0, 800.0, 8)
    times = np.arange(16) * 1e-3
    handle = _lib._P()
    for bad in (dict(num_freq=1), dict(num_time=1), dict(num_components=0), dict(num_components=17),
                dict(factor_freq_upsample=0), dict(factor_time_upsample=0)):
        fields = dict(num_freq=8, num_time=16, num_components=1, factor_freq_upsample=1, factor_time_upsample=1)
        fields.update(bad)
        desc = _lib.PlanDesc(fields["num_freq"], fields["num_time"], fields["num_components"],
                             fields["factor_freq_upsample"], fields["factor_time_upsample"], 0, 0, 0, 0.0, 4149.377593360996)
        code = native_lib.fb_plan_create(ctypes.byref(desc), _lib.dptr(freqs), _lib.dptr(times), ctypes.byref(handle))
        assert code == -1, (bad, code)
        assert native_lib.fb_last_error()
    assert native_lib.fb_plan_create(None, None, None, None) == -1
    x0 = np.zeros(3)
    trf = _lib._P()
    assert native_lib.fb_trf_create(0, 10, _lib.dptr(x0), None, ctypes.byref(trf)) == -1
    assert native_lib.fb_trf_create(68, 10, _lib.dptr(x0), None, ctypes.byref(trf)) == -1


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device the model raises instead of computing."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fitburst_b200 import _lib
    from fitburst_b200.analysis.model import SpectrumModeler
    model = SpectrumModeler(np.linspace(400.0, 800.0, 8), np.arange(16) * 1e-3)
    model.update_parameters({"amplitude": [0.0], "arrival_time": [0.005], "burst_width": [1e-3], "dm": [0.0],
                             "dm_index": [-2.0], "ref_freq": [600.0], "scattering_timescale": [0.0],
                             "scattering_index": [-4.0], "spectral_index": [0.0], "spectral_running": [0.0]})
    with pytest.raises(_lib.FitburstB200Error):
        model.compute_model()
