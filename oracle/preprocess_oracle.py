"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the pre-fit data preparation step
(ref: utilities/bases.py:47-336, routines/manipulate.py:10-72), written as pure functions on
numpy arrays. Pinned against the reference's own ReaderBaseClass by
tests/test_preprocess.py::test_oracle_matches_reference.
"""
import numpy as np

DISPERSION_CONSTANT = 4149.377593360996


def block_mean_2d(spectrum, factor_freq, factor_time):
    """ref: routines/manipulate.py:43-72 (mean over time blocks, then over frequency blocks)."""
    nf, nt = spectrum.shape
    return spectrum.reshape(nf // factor_freq, factor_freq, nt // factor_time, factor_time).mean(-1).mean(1)


def block_mean_1d(array, factor, boolean=False):
    """ref: routines/manipulate.py:10-41."""
    out = np.sum(np.reshape(array, (len(array) // factor, factor)), axis=1) / factor
    return out.astype(bool) if boolean else out


def preprocess(data, weights, apply_cut_variance=False, apply_cut_skewness=False, normalize_variance=True,
               remove_baseline=False, skewness_range=(-3.0, 3.0), variance_range=(0.2, 0.8), variance_weight=1.0):
    """ref: utilities/bases.py:201-298. Returns (cleaned data, good_freq)."""
    data = data.copy()
    mask_freq = np.sum(weights, -1)
    good = mask_freq != 0
    flat = data.min(axis=1) == data.max(axis=1)
    good = good & ~flat
    mean_spectrum = np.sum(data * weights, -1)
    mean_spectrum[good] /= mask_freq[good]
    if remove_baseline:
        data[good] /= mean_spectrum[good][:, None]
        data[good] -= 1
        data[np.logical_not(weights)] = 0
    variance = np.sum(data ** 2, -1)
    variance[good] /= mask_freq[good]
    skewness = np.sum(data ** 3, -1)
    skewness[good] /= mask_freq[good]
    sk_mean, sk_std = np.mean(skewness[good]), np.std(skewness[good])
    if normalize_variance:
        variance[good] /= np.max(variance[good])
    if apply_cut_variance:
        good = good & (variance / variance_weight < variance_range[1]) & (variance / variance_weight > variance_range[0])
    if apply_cut_skewness:
        good = good & (skewness < sk_mean + skewness_range[1] * sk_std) & (skewness > sk_mean + skewness_range[0] * sk_std)
    data[~good] = 0.0
    return data, good


def dedispersion_index(freqs, times, res_time, dm_value, arrival_time, ref_freq, dm_idx=-2.0,
                       is_dedispersed=False, dm_offset=None):
    """ref: utilities/bases.py:47-127."""
    def delay_of(dm):
        return dm * DISPERSION_CONSTANT * (freqs ** dm_idx - ref_freq ** dm_idx)
    idx = np.zeros(len(freqs), dtype=int)
    if not is_dedispersed:
        idx[:] = np.ceil((arrival_time + delay_of(dm_value) - times[0]) / res_time)
    elif dm_offset is not None:
        i0 = np.fabs(times - arrival_time).argmin()
        idx[:] = i0 + (np.around((delay_of(dm_value + dm_offset) - times[0]) / res_time) -
                       np.around((delay_of(dm_value) - times[0]) / res_time))
    return idx


def window(data, times, res_time, dedispersion_idx, arrival_time, window=0.08):
    """ref: utilities/bases.py:300-336."""
    i0 = np.fabs(times - arrival_time).argmin()
    nb = int(np.around(window / res_time))
    out = np.zeros((data.shape[0], 2 * nb))
    for i in range(data.shape[0]):
        out[i, :] = data[i, dedispersion_idx[i] - nb: dedispersion_idx[i] + nb]
    return out, times[i0 - nb: i0 + nb]
