"""
Configuration constants of the hot path.

The reference loads backend/general.yaml at import time into the module dict
`fitburst.backend.general` and reads two values from it at call time: the dispersion constant
(analysis/model.py:186,195, routines/derivative.py:240) and the `normalize_pbf` flag
(analysis/model.py:337, routines/derivative.py:96,136,187). When this package is used inside a
fitburst installation the very same dict is used, so a user who flips the flag there gets the
same behaviour here; stand-alone, the reference's shipped values (general.yaml:7-12) apply.
"""

try:  # pragma: no cover - only inside a fitburst installation
    from fitburst.backend import general  # type: ignore
except Exception:  # noqa: BLE001 - any import problem means "stand-alone"
    general = {
        "constants": {
            "dispersion": 4149.377593360996,
            "index_dispersion": -2.0,
            "index_scattering": -4.0,
        },
        "flags": {"normalize_pbf": False},
    }
