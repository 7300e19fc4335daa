"""Shared helpers for the parity tests: seeded cases and oracle/CUDA model pairs."""
import numpy as np

from fitburst_b200 import synthetic as syn


def make_case(num_freq=48, num_time=96, num_comp=3, kf=8, kt=4, folded=False, tau=2e-3,
              dm_inc=0.0, dm=0.0, scint=False):
    freqs, times = syn.chime_axes(num_freq, num_time)
    truth = syn.truncate_components(syn.config2_truth(num_time), num_comp)
    truth["scattering_timescale"] = [tau] * num_comp
    truth["dm"] = [dm] * num_comp
    kwargs = dict(dm_incoherent=dm_inc, factor_freq_upsample=kf, factor_time_upsample=kt,
                  num_components=num_comp, is_folded=folded, scintillation=scint)
    return freqs, times, truth, kwargs


def model_pair(freqs, times, kwargs, parameters):
    from fitburst_b200.analysis.model import SpectrumModeler
    from oracle.fitburst_oracle import OracleModel

    cuda = SpectrumModeler(freqs, times, **kwargs)
    oracle = OracleModel(freqs, times, **kwargs)
    cuda.update_parameters(parameters)
    oracle.update_parameters(parameters)
    return cuda, oracle


def rel_to_max(got, want):
    scale = np.max(np.abs(want))
    return np.max(np.abs(got - want)) / (scale if scale > 0 else 1.0)


def assert_close(got, want, rtol=1e-10, what=""):
    """Model tolerance of BASELINE.json: 1e-10 relative, element-wise, with an absolute floor of
    1e-10 * 1e-12 * max|want| so that denormal-range tail values do not dominate."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = np.max(np.abs(want))
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + rtol * 1e-12 * scale + 1e-300
    worst = np.max(err / tol)
    assert worst <= 1.0, f"{what}: worst error {worst:.3g} x tolerance (rel-to-max {rel_to_max(got, want):.3e})"
