"""Mirror of fitburst.analysis for the hot path (model, fitter)."""
from . import fitter, model  # noqa: F401
