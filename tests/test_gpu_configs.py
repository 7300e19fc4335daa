"""
Parity and timing of BASELINE.json's configs at (near) their stated sizes:

 * config 1 -- the tutorial workload stand-in of SURVEY.md section 8d: 16384 channels x 162
   samples, one Gaussian component, no up-sampling, pipeline-default fit (n = 6): the fit of a
   decimated channel axis against the oracle (same trajectory, best fit at 1e-6), the full-size
   model against the oracle on every 512th channel, the full-size fit against the injected burst;
 * config 2 -- 1024 samples, 3 scattered components, 8 x 4 up-sampling, n = 17: the fit of every
   64th channel at the FULL time resolution against the oracle (same status / nfev / njev, best
   fit at 1e-6). This is the subset bench.py's CPU leg times.
The frozen reference outputs for config 1 are tests/golden/config1_tutorial.npz (tests/test_golden.py).
"""
import contextlib
import io
import time

import numpy as np
import pytest

from fitburst_b200 import synthetic as syn
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731


def _fit_pair(freqs, times, kwargs, start, good, data, fixed):
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    from oracle.fitburst_oracle import OracleFitter, OracleModel

    cuda, oracle = SpectrumModeler(freqs, times, **kwargs), OracleModel(freqs, times, **kwargs)
    cuda.update_parameters(start)
    oracle.update_parameters(start)
    with quiet():
        fc, fo = LSFitter(data, cuda, good), OracleFitter(data, oracle, good)
        fc.fix_parameter(list(fixed))
        fo.fix_parameter(list(fixed))
        t0 = time.perf_counter()
        fc.fit()
        t_cuda = time.perf_counter() - t0
        t0 = time.perf_counter()
        fo.fit(with_statistics=False)
        t_cpu = time.perf_counter() - t0
    return fc, fo, t_cuda, t_cpu


def test_config1_decimated_fit_matches_oracle(native_lib):
    from oracle.fitburst_oracle import OracleModel
    freqs_full, times = syn.chime_axes(syn.CONFIG1_NUM_FREQ, syn.CONFIG1_NUM_TIME)
    idx = syn.channel_subset(syn.CONFIG1_NUM_FREQ, 64)
    freqs = freqs_full[idx]
    kwargs = dict(factor_freq_upsample=1, factor_time_upsample=1, num_components=1)
    truth = syn.config1_truth()
    clean_model = OracleModel(freqs, times, **kwargs)
    clean_model.update_parameters(truth)
    good = syn.config1_good_freq()[idx]
    data, _ = syn.config1_data(clean_model.compute_model(), good)
    fc, fo, _, _ = _fit_pair(freqs, times, kwargs, syn.config1_start(truth), good, data, syn.CONFIG1_FIXED)
    rc, ro = fc.results, fo.results
    assert (rc.status, rc.nfev, rc.njev) == (ro.status, ro.nfev, ro.njev)
    np.testing.assert_allclose(rc.x, ro.x, rtol=1e-6)  # BASELINE.json: best fit within 1e-6 relative
    assert abs(rc.cost - ro.cost) <= 1e-10 * ro.cost
    sigma = np.concatenate([np.atleast_1d(v) for v in fc.fit_statistics["bestfit_uncertainties"].values()])
    assert np.all(np.abs(rc.x - ro.x) < 1e-3 * sigma)  # well inside one sigma


def test_config1_full_size_model_and_fit(native_lib):
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    from oracle.fitburst_oracle import OracleModel
    freqs, times = syn.chime_axes(syn.CONFIG1_NUM_FREQ, syn.CONFIG1_NUM_TIME)
    kwargs = dict(factor_freq_upsample=1, factor_time_upsample=1, num_components=1)
    truth = syn.config1_truth()
    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(truth)
    full = model.compute_model()
    idx = syn.channel_subset(syn.CONFIG1_NUM_FREQ, 512)
    oracle = OracleModel(freqs[idx], times, **kwargs)
    oracle.update_parameters(truth)
    assert_close(full[idx], oracle.compute_model(), 1e-10, "config 1 full-size channel subset")
    good = syn.config1_good_freq()
    assert int(np.sum(~good)) == syn.CONFIG1_NUM_BAD
    data, sigma = syn.config1_data(full, good)
    model.update_parameters(syn.config1_start(truth))
    with quiet():
        fitter = LSFitter(data, model, good)
        fitter.fix_parameter(list(syn.CONFIG1_FIXED))
        assert len(fitter.get_fit_parameters_list()) == 6
        fitter.fit()
    res, stats = fitter.results, fitter.fit_statistics
    assert res.success and res.status in (1, 2, 3, 4)
    assert stats["num_freq_good"] == syn.CONFIG1_NUM_FREQ - syn.CONFIG1_NUM_BAD
    assert 35.0 < stats["snr"] < 45.0  # the stand-in is defined at S/N ~ 40
    assert 0.8 < stats["chisq_final_reduced"] < 1.2
    truth_packed = np.array([truth[name][0] for name in model.parameters if name in fitter.fit_parameters])
    unc = np.concatenate([np.atleast_1d(v) for v in stats["bestfit_uncertainties"].values()])
    assert np.all(np.abs(res.x - truth_packed) < 6.0 * unc), (res.x - truth_packed) / unc


def test_config2_full_time_resolution_subset_fit_matches_oracle(native_lib):
    """257 channels (every 64th of 16384, the original channel width) x 1024 samples x (8 x 4)
    sub-samples, 3 components, 17 free parameters: the CUDA fit follows scipy's trajectory on the
    oracle and ends at the same point."""
    from oracle.fitburst_oracle import OracleModel
    freqs_full, times = syn.chime_axes(16384, 1024)
    idx = syn.channel_subset(16384, 64)
    freqs = freqs_full[idx]
    kwargs = dict(factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
    truth = syn.config2_truth(1024)
    clean = OracleModel(freqs, times, **kwargs)
    clean.update_parameters(truth)
    good = syn.make_good_freq(len(freqs), 0.35, 2)
    data = syn.add_noise(clean.compute_model(), good, 0.1, 2)
    fc, fo, t_cuda, t_cpu = _fit_pair(freqs, times, kwargs, syn.config2_start(truth), good, data,
                                      ["dm_index", "scattering_index"])
    rc, ro = fc.results, fo.results
    assert len(rc.x) == 17
    assert (rc.status, rc.nfev, rc.njev) == (ro.status, ro.nfev, ro.njev)
    np.testing.assert_allclose(rc.x, ro.x, rtol=1e-6)
    assert abs(rc.cost - ro.cost) <= 1e-10 * ro.cost
    sigma = np.concatenate([np.atleast_1d(v) for v in fc.fit_statistics["bestfit_uncertainties"].values()])
    assert np.all(np.abs(rc.x - ro.x) < 1e-3 * sigma)
    print(f"config-2 subset fit: CUDA {t_cuda * 1e3:.1f} ms, oracle {t_cpu:.1f} s, nfev {rc.nfev}")
