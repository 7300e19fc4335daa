"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path (fitburst_b200/).

numpy restatement of the reference's small public routines next to the hot path, each citing the
reference lines it follows. Pinned on the unmodified reference by
tests/test_routines_and_stats.py::test_oracle_routines_match_reference (runs where /root/reference
is mounted).
"""
import numpy as np
from scipy.special import fdtr


def compute_spectrum_rpl(freqs, freq_ref, sp_idx, sp_run):  # routines/spectrum.py:39-41
    log_freq = np.log(freqs / freq_ref)
    return np.exp(sp_idx * log_freq + sp_run * log_freq ** 2)


def compute_time_dm_delay(dm_value, dm_const, dm_idx, freq1, freq2=np.inf):  # routines/ism.py:78
    return dm_value * dm_const * (freq1 ** dm_idx - freq2 ** dm_idx)


def compute_time_dm_smear(dm_value, dm_const, dm_idx, freq, width_chan):  # routines/ism.py:110
    return 2 * dm_value * dm_const * width_chan * freq ** (dm_idx - 1)


def compute_time_scattering(freq, ref_freq, sc_time_ref, sc_idx):  # routines/ism.py:139
    return sc_time_ref * (freq / ref_freq) ** sc_idx


def compute_amplitude_per_channel(data, model):  # routines/ism.py:30-46
    gram = model.T @ model
    return np.linalg.solve(gram, model.T @ data)


def downsample_1d(array, factor, boolean=False):  # routines/manipulate.py:33-40
    out = np.sum(np.reshape(array, (len(array) // factor, factor)), axis=1) / factor
    return out.astype(bool) if boolean else out


def downsample_2d(spectrum, factor_freq, factor_time):  # routines/manipulate.py:64-70
    num_freq, num_time = spectrum.shape
    return spectrum.reshape((num_freq // factor_freq, factor_freq, num_time // factor_time, factor_time)).mean(-1).mean(1)


def upsample_1d(array, diff_orig, factor):  # routines/manipulate.py:99-105
    diff_new = diff_orig / factor
    return np.linspace(array[0] - diff_orig / 2 + diff_new / 2, array[-1] + diff_orig / 2 - diff_new / 2,
                       num=len(array) * factor)


def compute_test_f(chisq_1, chisq_2, npar_1, npar_2, nobs_1, nobs_2):  # analysis/stats.py:44-59
    delta_chisq = chisq_2 - chisq_1
    dof_1, dof_2 = nobs_1 - npar_1, nobs_2 - npar_2
    probability = 1.
    if delta_chisq > 0.:
        probability -= fdtr(dof_2 - dof_1, dof_2, delta_chisq / (dof_2 - dof_1) / (chisq_2 / dof_2))
    return probability


def compute_downsampled_data(times, freqs, data, good_freq, model, factor_freq, factor_time):
    """utilities/plotting.py:57-88 (the matplotlib-free part of the module)."""
    good_ds = downsample_1d(good_freq, factor_freq, boolean=True)
    data_ds = downsample_2d(data * good_freq[:, None], factor_freq, factor_time)
    model_ds = downsample_2d(model, factor_freq, factor_time) * good_ds[:, None]
    freqs_ds, times_ds = downsample_1d(freqs, factor_freq), downsample_1d(times, factor_time)
    return {"data": data_ds, "freqs": freqs_ds, "good_freq": good_ds, "model": model_ds, "num_freq": len(freqs_ds),
            "num_time": len(times_ds), "res_freq": freqs_ds[1] - freqs_ds[0], "res_time": times_ds[1] - times_ds[0],
            "residuals": data_ds - model_ds, "times": times_ds}
