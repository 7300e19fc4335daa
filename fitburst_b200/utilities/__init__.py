"""Mirror of fitburst.utilities for the data-preparation step in front of the hot path."""
from . import bases  # noqa: F401
