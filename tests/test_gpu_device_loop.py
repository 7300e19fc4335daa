"""
GPU tests of the device-driven fit loop (fb_fit: k_trf_single after every evaluation, evaluations
enqueued ahead in chunks, early-out kernels) against the host-driven loop
(FITBURST_B200_DEVICE_LOOP=0: fb_trf.h on the host, one synchronisation per trial point), the
numpy oracle / scipy, and of the stepwise form fb_fit_device_* the channel-sharded fit uses.
"""
import contextlib
import ctypes
import io
import os

import numpy as np
import pytest

from fitburst_b200 import synthetic as syn
from tests.helpers import make_case, model_pair

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def host_loop():
    os.environ["FITBURST_B200_DEVICE_LOOP"] = "0"
    try:
        yield
    finally:
        del os.environ["FITBURST_B200_DEVICE_LOOP"]


def _problem(case, seed, sigma=0.1, bad=0.3, start_scale=None):
    freqs, times, truth, kwargs = make_case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    good = syn.make_good_freq(len(freqs), bad, seed)
    data = syn.add_noise(oracle.compute_model(), good, sigma, seed)
    start = syn.config2_start(truth)
    if start_scale:
        for key, factor in start_scale.items():
            start[key] = [v * factor for v in start[key]]
    return freqs, times, kwargs, truth, start, good, data


def _fit(freqs, times, kwargs, start, good, data, fixed):
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler

    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(start)
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = LSFitter(data, model, good)
        fitter.fix_parameter(list(fixed))
        fitter.fit()
    return fitter


CASES = {
    "fast_config2_shape": (dict(num_freq=64, num_time=256), ["dm_index", "scattering_index"], 2),
    "fast_two_components_all_global": (dict(num_freq=40, num_time=384, num_comp=2, kf=4, kt=2), [], 5),
    "fast_one_component": (dict(num_freq=48, num_time=192, num_comp=1, kf=2, kt=2), ["dm_index", "scattering_index"], 7),
    "general_gaussian_tau_fixed": (dict(num_freq=48, num_time=160, num_comp=2, kf=1, kt=1, tau=0.0),
                                   ["dm_index", "scattering_index", "scattering_timescale"], 3),
    "general_folded": (dict(num_freq=32, num_time=128, num_comp=2, kf=2, kt=2, folded=True), ["dm_index", "scattering_index"], 4),
    "general_eight_components": (dict(num_freq=32, num_time=640, num_comp=3, kf=1, kt=1), ["dm_index", "scattering_index", "spectral_running"], 6),
}


@pytest.mark.parametrize("name", list(CASES))
def test_device_loop_equals_host_loop(name, native_lib):
    case, fixed, seed = CASES[name]
    problem = _problem(case, seed)
    freqs, times, kwargs, truth, start, good, data = problem
    dev = _fit(freqs, times, kwargs, start, good, data, fixed)
    with host_loop():
        host = _fit(freqs, times, kwargs, start, good, data, fixed)
    rd, rh = dev.results, host.results
    assert (rd.status, rd.nfev, rd.njev) == (rh.status, rh.nfev, rh.njev)
    np.testing.assert_allclose(rd.x, rh.x, rtol=1e-9)
    assert abs(rd.cost - rh.cost) <= 1e-11 * abs(rh.cost)
    np.testing.assert_allclose(rd.grad, rh.grad, rtol=1e-6, atol=1e-9 * np.max(np.abs(rh.grad)) + 1e-12)
    np.testing.assert_allclose(dev.hessian, host.hessian, rtol=1e-8, atol=1e-10 * np.max(np.abs(host.hessian)))
    np.testing.assert_allclose(dev.hessian_approx, host.hessian_approx, rtol=1e-9,
                               atol=1e-12 * np.max(np.abs(host.hessian_approx)))
    for key, values in host.fit_statistics["bestfit_uncertainties"].items():
        np.testing.assert_allclose(dev.fit_statistics["bestfit_uncertainties"][key], values, rtol=1e-6)
    # both leave the model at the last evaluated point
    for key in ("arrival_time", "burst_width", "dm", "scattering_timescale"):
        np.testing.assert_allclose(getattr(dev.model, key), getattr(host.model, key), rtol=1e-9, atol=1e-15)


def test_device_loop_matches_scipy_trajectory(native_lib):
    """Same status / nfev / njev and best fit as scipy.optimize.least_squares on the oracle."""
    from oracle.fitburst_oracle import OracleFitter, OracleModel
    freqs, times, kwargs, truth, start, good, data = _problem(dict(num_freq=48, num_time=160, kf=4, kt=2), 9)
    fixed = ["dm_index", "scattering_index"]
    dev = _fit(freqs, times, kwargs, start, good, data, fixed)
    oracle = OracleModel(freqs, times, **kwargs)
    oracle.update_parameters(start)
    fo = OracleFitter(data, oracle, good)
    fo.fix_parameter(list(fixed))
    fo.fit(with_statistics=False)
    assert (dev.results.status, dev.results.nfev, dev.results.njev) == (fo.results.status, fo.results.nfev, fo.results.njev)
    np.testing.assert_allclose(dev.results.x, fo.results.x, rtol=1e-6)


def test_device_loop_hands_over_when_a_trial_leaves_the_fast_domain(native_lib):
    """Gaussian data fitted with a free, tiny scattering time: trial points with tau <= 0 are not
    covered by the fast kernels; the device loop must hand the state to the host loop, and the
    result must equal the all-host run."""
    freqs, times, kwargs, truth, start, good, data = _problem(
        dict(num_freq=48, num_time=160, num_comp=1, kf=2, kt=2, tau=2e-5), 12, sigma=0.3)
    start["scattering_timescale"] = [1e-5]
    fixed = ["dm_index", "scattering_index"]
    dev = _fit(freqs, times, kwargs, start, good, data, fixed)
    with host_loop():
        host = _fit(freqs, times, kwargs, start, good, data, fixed)
    assert (getattr(dev, "results", None) is None) == (getattr(host, "results", None) is None)
    if getattr(host, "results", None) is not None:
        assert (dev.results.status, dev.results.nfev, dev.results.njev) == \
            (host.results.status, host.results.nfev, host.results.njev)
        np.testing.assert_allclose(dev.results.x, host.results.x, rtol=1e-8)


def test_non_finite_start_reports_like_scipy(native_lib):
    freqs, times, kwargs, truth, start, good, data = _problem(dict(num_freq=24, num_time=96, kf=2, kt=2), 1)
    data = data.copy()
    data[np.flatnonzero(good)[0], 5] = np.nan
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(start)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        fitter = LSFitter(data, model, good, weighted_fit=False)
        fitter.fix_parameter(["dm_index", "scattering_index"])
        fitter.fit()
    assert not hasattr(fitter, "results") and "Residuals are not finite in the initial point" in out.getvalue()
    # the plan is usable afterwards (the early-out flag is disarmed)
    good_fit = _fit(freqs, times, kwargs, start, good, np.nan_to_num(data), ["dm_index", "scattering_index"])
    assert good_fit.results.status > 0


def test_stepwise_device_fit_single_rank(native_lib):
    """fb_fit_device_* (what ChannelShardedFitter drives, here without a process group) gives the
    fit of fb_fit; evaluations enqueued after the fit has stopped change nothing."""
    from fitburst_b200 import _lib
    from fitburst_b200.distributed import ChannelShardedFitter
    freqs, times, kwargs, truth, start, good, data = _problem(dict(num_freq=64, num_time=256), 21)
    fixed = ["dm_index", "scattering_index"]
    want = _fit(freqs, times, kwargs, start, good, data, fixed)
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(start)
    with contextlib.redirect_stdout(io.StringIO()):
        local = LSFitter(data, model, good)
        local.fix_parameter(fixed)
    sharded = ChannelShardedFitter(local, num_freq_total=len(freqs))
    res = sharded.fit()
    assert (res["status"], res["nfev"], res["njev"]) == (want.results.status, want.results.nfev, want.results.njev)
    np.testing.assert_allclose(res["x"], want.results.x, rtol=1e-12)
    model.update_parameters(start)
    res_host = sharded.fit(device_loop=False)
    assert (res_host["status"], res_host["nfev"]) == (res["status"], res["nfev"])
    np.testing.assert_allclose(res_host["x"], res["x"], rtol=1e-9)
    want_unc = np.concatenate([np.atleast_1d(v) for v in want.fit_statistics["bestfit_uncertainties"].values()])
    np.testing.assert_allclose(np.sort(res["uncertainties"]), np.sort(want_unc), rtol=1e-6)
    # protocol misuse is an error, not a crash
    plan, _ = local._prepare()
    stopped = ctypes.c_int32(0)
    assert native_lib.fb_fit_device_poll(plan, ctypes.byref(stopped), None, local._stream()) != 0


@pytest.mark.parametrize("name", ["fast_config2_shape", "fast_two_components_all_global", "general_eight_components"])
def test_trf_step_kernels_agree(name, native_lib):
    """The Cholesky step (default), the block-parallel eigen step (warm-started and cold Jacobi) and
    the one-warp eigen step take the same decisions and end at the same point."""
    case, fixed, seed = CASES[name]
    freqs, times, kwargs, truth, start, good, data = _problem(case, seed)
    runs = {}
    for label, env in (("chol", {}), ("block_warm", {"FITBURST_B200_TRF": "block"}),
                       ("block_cold", {"FITBURST_B200_TRF": "block", "FITBURST_B200_TRF_WARM": "0"}),
                       ("warp", {"FITBURST_B200_TRF": "warp"})):
        os.environ.update(env)
        try:
            runs[label] = _fit(freqs, times, kwargs, start, good, data, fixed).results
        finally:
            for key in env:
                del os.environ[key]
    ref = runs["warp"]
    for label in ("chol", "block_warm", "block_cold"):
        r = runs[label]
        assert (r.status, r.nfev, r.njev) == (ref.status, ref.nfev, ref.njev), label
        np.testing.assert_allclose(r.x, ref.x, rtol=1e-9, err_msg=label)


def test_fit_stream_float32_host_spectra(native_lib):
    """fit_stream on page-locked float32 spectra (uploaded as float32, widened by fb_widen_f32) gives
    the fits of the same values held as float64, with one and with several fits in flight."""
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.batch import fit_stream, pinned_like
    import torch
    case, fixed, _ = CASES["fast_config2_shape"]
    bursts32, bursts64 = [], []
    for seed in (31, 32, 33):
        freqs, times, kwargs, truth, start, good, data = _problem(case, seed)
        d32 = data.astype(np.float32)
        bursts32.append((torch.from_numpy(d32).pin_memory().numpy(), good, start))
        bursts64.append((pinned_like(d32.astype(np.float64)), good, start))
    model = SpectrumModeler(freqs, times, **kwargs)
    want = [o["results"].x for o in fit_stream(model, bursts64, fixed, workers=1)]
    for workers in (1, 3):
        got = [o["results"].x for o in fit_stream(model, bursts32, fixed, workers=workers)]
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
