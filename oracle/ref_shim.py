"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path (fitburst_b200/).

Import shim for the UNMODIFIED reference (CHIMEFRB/fitburst, mounted read-only at
/root/reference in the build container). The reference's hot-path modules drag in
matplotlib and pytz transitively (fitburst/utilities/plotting.py:12-14,
fitburst/routines/times.py:7), neither of which is installed, so empty stub modules
are registered before the import. Nothing of the reference is copied.

/root/reference does not exist on the GPU box: this module is used only by
oracle/make_golden.py (to freeze known-answer fixtures into tests/golden/) and by
CPU-side tests that validate the numpy restatement in oracle/fitburst_oracle.py.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FITBURST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fitburst"))


def import_reference():
    """Returns the imported reference package `fitburst` (stubs matplotlib/pytz)."""
    if not reference_available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    try:  # pandas (imported by the reference's peak finder) must see the REAL absence of pytz
        import pandas  # noqa: F401
    except Exception:  # noqa: BLE001
        pass
    for name in ("pytz", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        mpl.gridspec = sys.modules["matplotlib.gridspec"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fitburst  # noqa: E402

    return fitburst


@contextlib.contextmanager
def quiet():
    """The reference prints one (blank) line per component per compute_model call."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
