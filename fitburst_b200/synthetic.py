"""
Seeded synthetic CHIME-shaped inputs for the five BASELINE.json configs.

Everything here is axis/parameter bookkeeping plus a seeded noise draw; the noiseless burst
itself is produced by whichever model evaluator the caller passes in (the CUDA SpectrumModeler
in bench.py and the GPU tests, the reference / the oracle when golden fixtures are frozen), so
this module has no dependency on either.
"""
import copy

import numpy as np

CHIME_BAND_LO_MHZ = 400.1953125   # telescopes.yaml:25 (pivot of the lowest channel)
CHIME_BANDWIDTH_MHZ = 400.0
CHIME_RES_TIME_S = 0.00098304     # tutorial_1 notebook, cell 13


def chime_axes(num_freq: int = 16384, num_time: int = 1024, res_time: float = CHIME_RES_TIME_S):
    """
    CHIME/FRB-like axes: the generic reader builds channel centres as
    arange * res + freq_bin0 + res / 2 (reference backend/generic.py:90-96).
    """
    res_freq = CHIME_BANDWIDTH_MHZ / num_freq
    freqs = CHIME_BAND_LO_MHZ + (np.arange(num_freq, dtype=float) + 1.5) * res_freq
    times = np.arange(num_time, dtype=float) * res_time
    return freqs, times


def config2_truth(num_time: int = 1024, res_time: float = CHIME_RES_TIME_S) -> dict:
    """
    Config 2 burst: 3 scattered components with a running power law.
    Arrival times are quoted for the 1024-sample window and scale with the window length.
    """
    span = num_time * res_time / (1024 * CHIME_RES_TIME_S)
    return {
        "amplitude": [0.0, -0.3, -0.15],
        "arrival_time": [0.40 * span, 0.50 * span, 0.62 * span],
        "burst_width": [0.8e-3, 1.5e-3, 0.5e-3],
        "dm": [0.0, 0.0, 0.0],
        "dm_index": [-2.0, -2.0, -2.0],
        "ref_freq": [600.0, 600.0, 600.0],
        "scattering_timescale": [2.0e-3, 2.0e-3, 2.0e-3],
        "scattering_index": [-4.0, -4.0, -4.0],
        "spectral_index": [1.0, -1.0, 0.0],
        "spectral_running": [-3.0, 2.0, -8.0],
    }


def config2_start(truth: dict) -> dict:
    """Start point of the config-2 fit: truth perturbed by fixed offsets."""
    start = copy.deepcopy(truth)
    num_components = len(truth["amplitude"])
    shifts = [2e-4, -2e-4, 2e-4]
    scales = [1.2, 0.9, 1.1]
    for idx in range(num_components):
        start["arrival_time"][idx] = truth["arrival_time"][idx] + shifts[idx % 3]
        start["burst_width"][idx] = truth["burst_width"][idx] * scales[idx % 3]
    start["scattering_timescale"] = [x * 1.1 for x in truth["scattering_timescale"]]
    start["dm"] = [x + 0.01 for x in truth["dm"]]
    return start


def truncate_components(parameters: dict, num_components: int) -> dict:
    """First `num_components` components of a parameter dictionary."""
    return {key: list(value[:num_components]) for key, value in parameters.items()}


def random_burst(burst_index: int, num_time: int = 1024, res_time: float = CHIME_RES_TIME_S):
    """
    Config 3 burst `burst_index` (seeded by its index): 1-3 components, random DM offset,
    scattering time, widths and amplitudes. Returns (truth, start).
    """
    rng = np.random.default_rng(burst_index)
    num_components = 1 + (burst_index % 3)
    span = num_time * res_time
    widths = np.exp(rng.uniform(np.log(0.3e-3), np.log(3e-3), num_components))
    # arrivals spread at least 5 widths apart inside the middle half of the window.
    arrivals = np.sort(rng.uniform(0.30 * span, 0.70 * span, num_components))
    for idx in range(1, num_components):
        gap = 5 * max(widths[idx - 1], widths[idx])
        if arrivals[idx] - arrivals[idx - 1] < gap:
            arrivals[idx] = arrivals[idx - 1] + gap
    dm_offset = float(rng.uniform(-0.05, 0.05))
    sc_time = float(np.exp(rng.uniform(np.log(0.3e-3), np.log(10e-3))))
    truth = {
        "amplitude": rng.uniform(-0.5, 0.2, num_components).tolist(),
        "arrival_time": arrivals.tolist(),
        "burst_width": widths.tolist(),
        "dm": [dm_offset] * num_components,
        "dm_index": [-2.0] * num_components,
        "ref_freq": [600.0] * num_components,
        "scattering_timescale": [sc_time] * num_components,
        "scattering_index": [-4.0] * num_components,
        "spectral_index": rng.uniform(-2.0, 2.0, num_components).tolist(),
        "spectral_running": rng.uniform(-8.0, 2.0, num_components).tolist(),
    }
    return truth, config2_start(truth)


def make_good_freq(num_freq: int, bad_fraction: float, seed: int) -> np.ndarray:
    """Random RFI mask: True = usable channel."""
    rng = np.random.default_rng(seed + 7919)
    return rng.uniform(size=num_freq) >= bad_fraction


def add_noise(model: np.ndarray, good_freq: np.ndarray, sigma: float, seed: int) -> np.ndarray:
    """data = model + N(0, sigma); flagged channels are zeroed like the reference's readers do."""
    rng = np.random.default_rng(seed)
    data = model + rng.normal(0.0, sigma, size=model.shape)
    data[~good_freq, :] = 0.0
    return data


# ---- config 1: the tutorial workload (BASELINE.json configs[0]) ------------------------------------
# The .npz of the tutorial is not shipped with the reference (.MISSING_LARGE_BLOBS:1); SURVEY.md
# section 8d defines this stand-in from what the notebooks print: 16384 channels x 162 samples,
# one Gaussian (unscattered) component with the best-fit values of tutorial_3_model.ipynb cell 13,
# 5672 flagged channels (tutorial_2 cell 12), S/N ~ 40, seed 20220502, no up-sampling, and the
# pipeline's default fit (dm_index, scattering_index, scattering_timescale fixed: n = 6,
# pipelines/fitburst_pipeline.py:331).
CONFIG1_NUM_FREQ, CONFIG1_NUM_TIME, CONFIG1_SEED, CONFIG1_NUM_BAD = 16384, 162, 20220502, 5672
CONFIG1_FIXED = ["dm_index", "scattering_index", "scattering_timescale"]


def config1_truth(num_time: int = CONFIG1_NUM_TIME, res_time: float = CHIME_RES_TIME_S) -> dict:
    times = np.arange(num_time, dtype=float) * res_time
    return {
        "amplitude": [-0.4812764314320905],
        "arrival_time": [float(np.mean(times))],  # tutorial_3 cell 17
        "burst_width": [2.7258462551453457e-4],
        "dm": [0.0],
        "dm_index": [-2.0],
        "ref_freq": [400.1953125],
        "scattering_timescale": [0.0],
        "scattering_index": [-4.0],
        "spectral_index": [2.8705490561459213],
        "spectral_running": [-15.86167822219646],
    }


def config1_start(truth: dict) -> dict:
    """dm + 0.02, width x 1.5, amplitude - 0.1, arrival time + 2e-4."""
    start = copy.deepcopy(truth)
    start["dm"] = [truth["dm"][0] + 0.02]
    start["burst_width"] = [truth["burst_width"][0] * 1.5]
    start["amplitude"] = [truth["amplitude"][0] - 0.1]
    start["arrival_time"] = [truth["arrival_time"][0] + 2e-4]
    return start


def config1_good_freq(num_freq: int = CONFIG1_NUM_FREQ, seed: int = CONFIG1_SEED) -> np.ndarray:
    """Exactly round(5672 / 16384 * num_freq) randomly chosen flagged channels."""
    rng = np.random.default_rng(seed)
    good = np.ones(num_freq, dtype=bool)
    good[rng.choice(num_freq, size=int(round(CONFIG1_NUM_BAD * num_freq / CONFIG1_NUM_FREQ)), replace=False)] = False
    return good


def config1_data(clean: np.ndarray, good_freq: np.ndarray, snr: float = 40.0, seed: int = CONFIG1_SEED):
    """Noise level such that the matched-filter S/N over the usable channels is `snr`."""
    sigma = float(np.sqrt(np.sum(clean[good_freq] ** 2)) / snr)
    return add_noise(clean, good_freq, sigma, seed), sigma


def channel_subset(num_freq: int, stride: int, offset: int = 0) -> np.ndarray:
    """Indices of every `stride`-th channel, keeping the first pair adjacent: the models take the
    channel width from freqs[1] - freqs[0] (analysis/model.py:87), so a decimated axis must start
    with two neighbouring channels to keep the ORIGINAL width."""
    idx = np.arange(offset % stride, num_freq - 1, stride)
    return np.unique(np.concatenate([idx[:1], idx[:1] + 1, idx[1:]]))
