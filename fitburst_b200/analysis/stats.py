"""
Mirror of fitburst.analysis.stats (reference analysis/stats.py:11-59) plus the batched form the
multi-burst callers need: the F-test that compares two nested best-fit models of the same burst
(e.g. the 1- vs 2-component fits BurstFitter / fit_stream produce for thousands of bursts).

Scalar special-function work on fit STATISTICS (a handful of numbers per burst), not on spectra:
it stays on the host with scipy.special.fdtr exactly like the reference.
"""
import numpy as np
from scipy.special import fdtr  # pylint: disable=no-name-in-module


def compute_test_f(chisq_1: float, chisq_2: float, num_fit_parameters_1: int, num_fit_parameters_2: int,
                   num_observations_1: int, num_observations_2: int) -> float:
    """F-test chance-coincidence probability (analysis/stats.py:11-59, same argument order)."""
    # pylint: disable=too-many-arguments
    delta_chisq = chisq_2 - chisq_1
    deg_freedom_1 = num_observations_1 - num_fit_parameters_1
    deg_freedom_2 = num_observations_2 - num_fit_parameters_2
    delta_deg_freedom = deg_freedom_2 - deg_freedom_1
    probability = 1.
    if delta_chisq > 0.:
        f_value = delta_chisq / delta_deg_freedom / (chisq_2 / deg_freedom_2)
        probability -= fdtr(delta_deg_freedom, deg_freedom_2, f_value)
    return probability


def compute_test_f_batched(statistics_1, statistics_2) -> np.ndarray:
    """
    compute_test_f over two equally long sequences of LSFitter.fit_statistics dictionaries (or
    fit_stream / BurstFitter outputs, whose "fit_statistics" member is used): entry k compares
    model #1 and model #2 of burst k. The reference passes fit_statistics["num_observations"]
    (already net of the fit parameters, analysis/fitter.py:461) as num_observations, and so does
    this function. Bursts without a successful fit in either list give NaN.
    """
    def fields(item):
        stats = item.get("fit_statistics", item) if item is not None else None
        if not stats or stats.get("chisq_final") is None:
            return np.nan, np.nan, np.nan
        return float(stats["chisq_final"]), float(stats["num_fit_parameters"]), float(stats["num_observations"])

    if len(statistics_1) != len(statistics_2):
        raise ValueError("the two result lists must have the same length")
    a = np.array([fields(item) for item in statistics_1], dtype=np.float64).reshape(-1, 3)
    b = np.array([fields(item) for item in statistics_2], dtype=np.float64).reshape(-1, 3)
    delta_chisq = b[:, 0] - a[:, 0]
    dof_1, dof_2 = a[:, 2] - a[:, 1], b[:, 2] - b[:, 1]
    delta_dof = dof_2 - dof_1
    probability = np.ones(len(a), dtype=np.float64)
    with np.errstate(all="ignore"):
        f_value = delta_chisq / delta_dof / (b[:, 0] / dof_2)
        positive = delta_chisq > 0.
        probability[positive] -= fdtr(delta_dof[positive], dof_2[positive], f_value[positive])
    probability[np.isnan(delta_chisq)] = np.nan
    return probability
