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
