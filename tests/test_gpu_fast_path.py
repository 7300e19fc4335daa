"""
GPU tests of the warp-granular fast kernels (fb_fast.cuh: k_vals3, k_gram3 with its far-tile
path, k_hess3 with the far-sample moments) at window lengths where their special cases actually
occur -- whole tiles in the exponential tail, far tiles, near ranges -- against the numpy oracle
and against the general kernels (FITBURST_B200_FAST=0) on identical inputs.
"""
import contextlib
import io
import os

import numpy as np
import pytest

from fitburst_b200 import synthetic as syn
from tests.helpers import model_pair, rel_to_max

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def general_kernels():
    """Routes the evaluation through the block-granular kernels of fb_kernels.cuh."""
    os.environ["FITBURST_B200_FAST"] = "0"
    try:
        yield
    finally:
        del os.environ["FITBURST_B200_FAST"]


def _case(num_freq=12, num_time=1024, num_comp=3, kf=8, kt=4, ref_freqs=None, dm=0.0, dm_inc=0.0, widths=None):
    freqs, times = syn.chime_axes(num_freq, num_time)
    truth = syn.truncate_components(syn.config2_truth(num_time), min(num_comp, 3))
    if num_comp == 4:  # a fourth component late in the window
        for key, value in (("amplitude", -0.2), ("arrival_time", 0.75 * times[-1]), ("burst_width", 1.0e-3),
                           ("dm", 0.0), ("dm_index", -2.0), ("ref_freq", 600.0), ("scattering_timescale", 2.0e-3),
                           ("scattering_index", -4.0), ("spectral_index", 0.5), ("spectral_running", -1.0)):
            truth[key] = list(truth[key]) + [value]
    truth["dm"] = [dm] * num_comp
    if ref_freqs is not None:
        truth["ref_freq"] = list(ref_freqs)
    if widths is not None:
        truth["burst_width"] = list(widths)
    kwargs = dict(dm_incoherent=dm_inc, factor_freq_upsample=kf, factor_time_upsample=kt, num_components=num_comp)
    return freqs, times, truth, kwargs


def _fitters(case, fixed, seed=3):
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter
    freqs, times, truth, kwargs = _case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    good = syn.make_good_freq(len(freqs), 0.25, seed)
    data = syn.add_noise(oracle.compute_model(), good, 0.1, seed)
    start = syn.config2_start(truth)
    cuda.update_parameters(start)
    oracle.update_parameters(start)
    with contextlib.redirect_stdout(io.StringIO()):
        fc = LSFitter(data, cuda, good)
        fo = OracleFitter(data, oracle, good)
        fc.fix_parameter(list(fixed))
        fo.fix_parameter(list(fixed))
    return fc, fo


def _gram(fitter, native_lib):
    from fitburst_b200 import _lib
    plan, n = fitter._prepare()
    fitter.model._push_parameters()
    gram = np.zeros((n + 1, n + 1))
    _lib.check(native_lib.fb_normal_equations(plan, fitter._data_on_device().data_ptr(), _lib.dptr(gram),
                                             fitter._stream()))
    return gram


GRAM_CASES = {
    "config2_columns": (dict(), ["dm_index", "scattering_index"]),
    "all_nine_free": (dict(), []),
    "reference_frequencies_differ": (dict(ref_freqs=[600.0, 500.0, 720.0]), ["dm_index", "scattering_index"]),
    "one_component": (dict(num_comp=1), ["dm_index", "scattering_index"]),
    "two_components_dm_offset": (dict(num_comp=2, dm=0.2), ["scattering_index"]),
    "four_components": (dict(num_comp=4), ["dm_index", "scattering_index"]),
    "generic_upsampling": (dict(kf=3, kt=2), ["dm_index", "scattering_index"]),
    "no_upsampling": (dict(kf=1, kt=1), ["dm_index"]),
    "ragged_window": (dict(num_time=1000), ["dm_index", "scattering_index"]),
    "incoherent_dm": (dict(dm_inc=30.0, dm=0.05, kf=4, kt=2), ["dm_index", "scattering_index"]),
    "only_shapes_free": (dict(), ["amplitude", "spectral_index", "spectral_running", "dm_index", "scattering_index"]),
    "wide_bursts_no_far_tiles": (dict(num_time=512, widths=[6e-3, 8e-3, 5e-3]), ["dm_index", "scattering_index"]),
    # short windows with more channels than blocks: the prefetch ring of k_gram3 runs ahead across
    # several channel boundaries (two / one tiles per warp and channel)
    "short_window_many_channels": (dict(num_freq=2048, num_time=256, kf=2, kt=2), ["dm_index", "scattering_index"]),
    "tutorial_window_many_channels": (dict(num_freq=1500, num_time=162, kf=1, kt=1), ["dm_index", "scattering_index"]),
}


@pytest.mark.parametrize("name", list(GRAM_CASES))
def test_normal_equations_fast_vs_oracle_and_general(name, native_lib):
    case, fixed = GRAM_CASES[name]
    fc, fo = _fitters(case, fixed)
    x0 = fo.get_fit_parameters_list()
    fast = _gram(fc, native_lib)  # k_vals3 + k_gram3
    with general_kernels():
        general = _gram(fc, native_lib)
    os.environ["FITBURST_B200_EVAL"] = "4"  # opt-in second generation: k_rise3 + k_gram4 (no HBM round trip)
    try:
        gen4 = _gram(fc, native_lib)
    finally:
        del os.environ["FITBURST_B200_EVAL"]
    # opt-in row fetch through the bulk-copy (TMA) unit: same arithmetic, bit-identical result
    # (read at the first fast launch of the process: set for the whole test run by
    # FITBURST_B200_BULK=1 python -m pytest ...; here only when the variable is already set)
    if os.environ.get("FITBURST_B200_BULK") == "1":
        assert np.array_equal(fast, _gram(fc, native_lib))
    resid = fo.compute_residuals(x0)  # also refreshes the caches the oracle Jacobian reads
    full = np.concatenate([fo.compute_jacobian(x0), resid[:, None]], axis=1)
    want = full.T @ full
    scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
    assert np.max(np.abs(fast - want) / scale) < 1e-9, "fast kernels vs oracle"
    assert np.max(np.abs(fast - general) / scale) < 1e-11, "fast vs general kernels"
    assert np.max(np.abs(gen4 - general) / scale) < 1e-11, "second-generation fast vs general kernels"
    assert np.array_equal(fast, fast.T)


@pytest.mark.parametrize("num_time,kf,kt", [(33, 2, 2), (64, 1, 1), (97, 2, 2), (130, 4, 2), (200, 1, 2), (320, 2, 1),
                                            (450, 2, 2)])
def test_window_lengths_with_several_channels_per_block(num_time, kf, kt, native_lib):
    """More good channels than blocks at window lengths around every tile-count boundary of the
    fast kernels (64-sample tiles in k_vals3, 32-sample tiles and a three-tile prefetch ring in
    k_gram3, eight-tile rounds in k_hess3): normal equations and Hessian equal the general
    kernels'."""
    fc, _ = _fitters(dict(num_freq=1300, num_time=num_time, kf=kf, kt=kt), ["dm_index", "scattering_index"])
    fast = _gram(fc, native_lib)
    with contextlib.redirect_stdout(io.StringIO()):
        h_fast, _ = fc.compute_hessian(fc.data, fc.fit_parameters)
        with general_kernels():
            general = _gram(fc, native_lib)
            h_general, _ = fc.compute_hessian(fc.data, fc.fit_parameters)
    assert np.all(np.isfinite(fast)) and np.all(np.isfinite(h_fast))
    scale = np.sqrt(np.outer(np.diag(general), np.diag(general)))
    assert np.max(np.abs(fast - general) / scale) < 1e-11
    hs = np.sqrt(np.outer(np.abs(np.diag(h_general)), np.abs(np.diag(h_general))))
    assert np.all(np.abs(h_fast - h_general) <= 1e-9 * hs + 1e-12 * np.abs(h_general))


@pytest.mark.parametrize("case,fixed", [
    (dict(num_freq=8), ["dm_index", "scattering_index"]),
    (dict(num_freq=6, num_comp=2, kf=2, kt=2), ["dm_index"]),
    (dict(num_freq=6, num_comp=2, kf=2, kt=2), []),  # dm_index free: no far-sample shortcut
    (dict(num_freq=6, ref_freqs=[600.0, 500.0, 720.0], kf=2, kt=2), ["dm_index", "scattering_index"]),
])
def test_hessian_fast_vs_oracle_and_general(case, fixed, native_lib):
    from oracle import hessian_oracle
    fc, fo = _fitters(case, fixed)
    with contextlib.redirect_stdout(io.StringIO()):
        fast, labels = fc.compute_hessian(fc.data, fc.fit_parameters)
        with general_kernels():
            general, _ = fc.compute_hessian(fc.data, fc.fit_parameters)
    want, want_labels = hessian_oracle.compute_hessian(fo, fo.data)
    assert labels == want_labels
    # entries are compared against the geometric mean of their diagonal entries; the reference's
    # quirky dm x dm_index term can exceed that scale by nine orders of magnitude, so its own
    # rounding floor (1e-12 relative) is allowed for as well
    scale = np.sqrt(np.outer(np.abs(np.diag(want)), np.abs(np.diag(want))))
    assert np.all(np.abs(fast - want) <= 1e-8 * scale + 1e-12 * np.abs(want)), "fast Hessian vs oracle"
    assert np.all(np.abs(fast - general) <= 1e-10 * scale + 1e-12 * np.abs(want)), "fast vs general Hessian"


def test_fit_trajectory_long_window(native_lib):
    """A fit at a window length with far tiles: same trajectory as scipy on the oracle."""
    fc, fo = _fitters(dict(num_freq=24, num_time=768), ["dm_index", "scattering_index"])
    with contextlib.redirect_stdout(io.StringIO()):
        fc.fit()
        fo.fit(with_statistics=False)
    assert fc.results.status == fo.results.status
    assert fc.results.nfev == fo.results.nfev and fc.results.njev == fo.results.njev
    np.testing.assert_allclose(fc.results.x, fo.results.x, rtol=1e-6)


def test_chisq_grid_fast_vs_general(native_lib):
    from fitburst_b200.batch import chisq_grid
    fc, _ = _fitters(dict(num_freq=16, num_time=512), ["dm_index", "scattering_index"])
    dm_values = np.linspace(-0.05, 0.05, 3)
    sc_values = np.logspace(-3.5, -2.0, 3)
    fast = chisq_grid(fc, dm_values, sc_values)
    with general_kernels():
        general = chisq_grid(fc, dm_values, sc_values)
    np.testing.assert_allclose(fast, general, rtol=1e-12)


@pytest.mark.parametrize("case", [
    dict(num_freq=16, num_time=512),
    dict(num_freq=24, num_time=1024, num_comp=4),
    dict(num_freq=10, num_time=301, num_comp=2, kf=4, kt=2),       # odd row length: no 16-byte pairs
    dict(num_freq=8, num_time=256, num_comp=1, kf=1, kt=1),        # in-place rise region (small sub-grid)
    dict(num_freq=8, num_time=640, widths=[4.0e-3, 6.0e-3, 5.0e-3]),  # long rise regions: queue overflows
], ids=["three_components", "four_components", "odd_length", "no_upsampling", "queue_overflow"])
def test_chisq_grid_fused_kernel(native_lib, case):
    """The one-kernel sweep (k_vals3<.., CHI2>) against the stored-values sweep of round 1, the
    general kernels and the reference's definition sum(compute_residuals(x)**2)
    (analysis/fitter.py:231-266) evaluated by the oracle."""
    from fitburst_b200.batch import chisq_grid
    fc, fo = _fitters(case, ["dm_index", "scattering_index"])
    dm_values = np.linspace(-0.04, 0.06, 3)
    sc_values = np.logspace(-3.5, -2.0, 4)
    fused = chisq_grid(fc, dm_values, sc_values)
    os.environ["FITBURST_B200_GRID_FUSED"] = "0"
    try:
        stored = chisq_grid(fc, dm_values, sc_values)
    finally:
        del os.environ["FITBURST_B200_GRID_FUSED"]
    with general_kernels():
        general = chisq_grid(fc, dm_values, sc_values)
    np.testing.assert_allclose(fused, stored, rtol=1e-13)
    np.testing.assert_allclose(fused, general, rtol=1e-12)
    fo.fit_parameters = ["dm", "scattering_timescale"]
    want = np.array([[np.sum(fo.compute_residuals([dm, sc]) ** 2) for sc in sc_values] for dm in dm_values])
    np.testing.assert_allclose(fused, want, rtol=1e-10)


def test_fit_stream_matches_individual_fits(native_lib):
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.batch import fit_stream, pinned_like
    freqs, times, truth, kwargs = _case(num_freq=16, num_time=256, kf=2, kt=2)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    clean = oracle.compute_model()
    start = syn.config2_start(truth)
    bursts = []
    for seed in range(5):
        good = syn.make_good_freq(len(freqs), 0.2, seed)
        bursts.append((pinned_like(syn.add_noise(clean, good, 0.1, seed)), good, start))
    fixed = ["dm_index", "scattering_index"]
    streamed = list(fit_stream(cuda, bursts, fixed))
    assert len(streamed) == len(bursts)
    for (data, good, begin), out in zip(bursts, streamed):
        cuda.update_parameters(begin)
        with contextlib.redirect_stdout(io.StringIO()):
            single = LSFitter(data, cuda, good)
            single.fix_parameter(fixed)
            single.fit()
        assert out["results"].status == single.results.status and out["results"].nfev == single.results.nfev
        np.testing.assert_array_equal(out["results"].x, single.results.x)
        assert out["fit_statistics"]["bestfit_uncertainties"] == single.fit_statistics["bestfit_uncertainties"]


def test_fit_stream_concurrent_workers_identical(native_lib):
    """Several fits in flight (threads, streams, model clones): same per-burst results, in order."""
    from fitburst_b200.batch import fit_stream
    freqs, times, truth, kwargs = _case(num_freq=32, num_time=256, kf=2, kt=2)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    clean = oracle.compute_model()
    bursts = []
    for seed in range(9):
        good = syn.make_good_freq(len(freqs), 0.2, seed)
        start = syn.config2_start(truth)
        start["dm"] = [0.01 + 0.002 * seed] * 3
        bursts.append((syn.add_noise(clean, good, 0.1, seed), good, start))
    fixed = ["dm_index", "scattering_index"]
    serial = list(fit_stream(cuda, bursts, fixed))
    threaded = list(fit_stream(cuda, bursts, fixed, workers=3))
    assert len(threaded) == len(serial) == len(bursts)
    for a, b in zip(serial, threaded):
        assert a["results"].status == b["results"].status and a["results"].nfev == b["results"].nfev
        np.testing.assert_array_equal(a["results"].x, b["results"].x)
        assert a["fit_statistics"]["bestfit_uncertainties"] == b["fit_statistics"]["bestfit_uncertainties"]
    # float32 host spectra are uploaded as float32 and widened on the device: same results as
    # fitting the widened array
    as_f32 = [(d.astype(np.float32), g, st) for d, g, st in bursts[:4]]
    widened = [(d.astype(np.float64), g, st) for d, g, st in as_f32]
    for workers in (1, 2):
        got = list(fit_stream(cuda, as_f32, fixed, workers=workers))
        want = list(fit_stream(cuda, widened, fixed, workers=workers))
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a["results"].x, b["results"].x)
    # page-locked spectra: plain copy, and the opt-in fb_upload_rows path (only the usable
    # channels cross the host link)
    from fitburst_b200.batch import pinned_like
    pinned = [(pinned_like(d), g, st) for d, g, st in bursts]
    for row_upload in ("0", "1"):
        os.environ["FITBURST_B200_ROW_UPLOAD"] = row_upload
        try:
            for workers in (1, 3):
                for a, b in zip(serial, fit_stream(cuda, pinned, fixed, workers=workers)):
                    assert a["results"].nfev == b["results"].nfev
                    np.testing.assert_array_equal(a["results"].x, b["results"].x)
                    assert a["fit_statistics"] == b["fit_statistics"]
        finally:
            del os.environ["FITBURST_B200_ROW_UPLOAD"]
    # device-resident inputs and an early stop of the consumer
    import torch
    resident = [(torch.from_numpy(d).cuda(), g, s) for d, g, s in bursts]
    stream = fit_stream(cuda, resident, fixed, workers=2)
    first = next(stream)
    stream.close()
    np.testing.assert_array_equal(first["results"].x, serial[0]["results"].x)
    # a slow consumer fills the look-ahead window (the workers' non-blocking take comes back
    # empty and they upload after publishing instead); host and device inputs mixed in one list,
    # every worker alternating between its two upload buffers many times
    import time
    mixed = [(resident[k % 9] if k % 4 == 3 else pinned[k % 9]) for k in range(22)]
    got = []
    for k, out in enumerate(fit_stream(cuda, mixed, fixed, workers=2)):
        if k < 4:
            time.sleep(0.05)
        got.append(out)
    assert len(got) == len(mixed)
    for k, out in enumerate(got):
        np.testing.assert_array_equal(out["results"].x, serial[k % 9]["results"].x)
        assert out["fit_statistics"] == serial[k % 9]["fit_statistics"]


def test_full_size_fast_vs_general(native_lib):
    """Config-2 size: the fast and the general kernels agree on [J r]^T [J r] and on the Hessian."""
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    freqs, times = syn.chime_axes(16384, 1024)
    truth = syn.config2_truth(1024)
    model = SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
    model.update_parameters(truth)
    good = syn.make_good_freq(16384, 0.35, 5)
    data = syn.add_noise(model.compute_model(), good, 0.1, 5)
    model.update_parameters(syn.config2_start(truth))
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = LSFitter(data, model, good)
        fitter.fix_parameter(["dm_index", "scattering_index"])
        fast = _gram(fitter, native_lib)
        h_fast, _ = fitter.compute_hessian(fitter.data, fitter.fit_parameters)
        with general_kernels():
            general = _gram(fitter, native_lib)
            h_general, _ = fitter.compute_hessian(fitter.data, fitter.fit_parameters)
    scale = np.sqrt(np.outer(np.diag(general), np.diag(general)))
    assert np.max(np.abs(fast - general) / scale) < 1e-11
    hs = np.sqrt(np.outer(np.abs(np.diag(h_general)), np.abs(np.diag(h_general))))
    assert np.max(np.abs(h_fast - h_general) / hs) < 1e-9


MODEL_REGIMES = {
    # (dm offset, scattering time at ref_freq, widths): the sub-frequency expansion of the rise
    # region is taken or refused by its truncation bound; either way the model must hold 1e-10
    "expansion_config2": (0.0, 2e-3, None),
    "expansion_small_dm_offset": (0.02, 2e-3, None),
    # beyond the moment bound: the per-sub-frequency series (series_eval) takes over
    "series_large_dm_offset": (0.5, 2e-3, None),
    "series_weak_scattering": (0.0, 2e-4, None),
    "series_very_weak_scattering": (0.0, 2e-5, None),
    "series_weak_scattering_and_dm_offset": (-0.3, 1e-4, None),
    # beyond the series bound as well: exact sub-grid
    "exact_huge_dm_offset_narrow_components": (4.0, 5e-4, [2e-4, 3e-4, 2e-4]),
    "expansion_strong_scattering": (0.01, 2e-2, None),
    "mixed_wide_and_narrow": (0.005, 1e-3, [3e-3, 4e-4, 1e-3]),
    # a component wider than the window: every sample is a rise-region sample, the per-channel
    # queue of k_vals3 overflows and the refused tiles are evaluated in place
    "queue_overflow_window_wide_component": (0.03, 2.3e-3, [3.6e-4, 47.0, 2.5e-3]),
}


@pytest.mark.parametrize("name", list(MODEL_REGIMES))
def test_model_rise_region_regimes(name, native_lib):
    from tests.helpers import assert_close
    dm, tau, widths = MODEL_REGIMES[name]
    num_time = 1024 if name.startswith("queue_overflow") else 512
    freqs, times, truth, kwargs = _case(num_freq=40, num_time=num_time, dm=dm, widths=widths)
    truth["scattering_timescale"] = [tau] * 3
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    got, want = cuda.compute_model(), oracle.compute_model()
    assert_close(got, want, 1e-10, name)
    assert rel_to_max(got, want) < 2e-12


@pytest.mark.parametrize("seed", range(16))
def test_model_random_regimes(seed, native_lib):
    """Random scattering times, widths and DM offsets across the moment / series / exact domains of
    the rise-region evaluation (make_tay_tab, series_eval): model within 1e-10 of the oracle and
    the fused chi^2 kernel equal to the stored-values sweep."""
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.batch import chisq_grid
    from tests.helpers import assert_close
    rng = np.random.default_rng(1000 + seed)
    tau = float(10 ** rng.uniform(np.log10(2e-5), np.log10(3e-2)))
    dm = float(rng.uniform(-0.6, 0.6))
    widths = [float(10 ** rng.uniform(np.log10(2e-4), np.log10(5e-3))) for _ in range(3)]
    freqs, times, truth, kwargs = _case(num_freq=24, num_time=512, dm=dm, widths=widths)
    truth["scattering_timescale"] = [tau] * 3
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    got, want = cuda.compute_model(), oracle.compute_model()
    what = f"tau={tau:.3g} dm={dm:.3g} widths={widths}"
    assert_close(got, want, 1e-10, what)
    assert rel_to_max(got, want) < 2e-12, what
    good = syn.make_good_freq(len(freqs), 0.2, seed)
    data = syn.add_noise(want, good, 0.1, seed)
    with contextlib.redirect_stdout(io.StringIO()):
        fitter = LSFitter(data, cuda, good)
    dm_values, sc_values = np.array([dm, 0.5 * dm]), np.array([tau, 2.0 * tau])
    fused = chisq_grid(fitter, dm_values, sc_values)
    os.environ["FITBURST_B200_GRID_FUSED"] = "0"
    try:
        stored = chisq_grid(fitter, dm_values, sc_values)
    finally:
        del os.environ["FITBURST_B200_GRID_FUSED"]
    np.testing.assert_allclose(fused, stored, rtol=1e-13, err_msg=what)
    resid = (data - want) * fitter.weights[:, None]
    np.testing.assert_allclose(fused[0, 0], np.sum(resid[good] ** 2), rtol=1e-9, err_msg=what)


def test_normal_equations_queue_overflow(native_lib):
    """Trial points far from the burst (a component wider than the window, arrival outside it --
    seen on weak components of config-3 bursts): fast and general kernels agree, nothing faults."""
    fc, fo = _fitters(dict(num_freq=16, num_time=1024), ["dm_index", "scattering_index"])
    far_out = fc.model.get_parameters_dict()
    far_out["burst_width"] = [3.6e-4, 47.0, 2.5e-3]
    far_out["arrival_time"] = [0.397, -30.5, 0.9]
    far_out["amplitude"] = [-0.44, -29.6, -2.9]
    fc.model.update_parameters(far_out)
    fast = _gram(fc, native_lib)
    with general_kernels():
        general = _gram(fc, native_lib)
    scale = np.sqrt(np.outer(np.diag(general), np.diag(general)))
    scale[scale == 0.0] = 1.0
    assert np.all(np.isfinite(fast))
    err = np.abs(fast - general) / scale
    # the first derivatives of the 47-second-wide, 1e-30-amplitude component are differences of
    # nearly equal terms (condition number ~1e12): its columns only agree to the rounding of that
    # cancellation, every other column to the usual bound
    wild = np.zeros(fast.shape[0], dtype=bool)
    wild[[1, 4, 7, 12, 15]] = True
    tame = ~wild[:, None] & ~wild[None, :]
    assert np.max(err[tame]) < 1e-10
    assert np.max(err) < 1e-2


@pytest.mark.parametrize("dtype,num_time", [(np.float64, 256), (np.float64, 131), (np.float32, 256), (np.float32, 130)])
def test_upload_rows_copies_exactly_the_listed_channels(dtype, num_time, native_lib):
    import torch
    from fitburst_b200 import _lib
    rng = np.random.default_rng(11)
    num_freq = 300
    host = torch.from_numpy(rng.normal(size=(num_freq, num_time)).astype(dtype)).pin_memory()
    good = rng.uniform(size=num_freq) > 0.4
    rows = torch.from_numpy(np.flatnonzero(good).astype(np.int32)).cuda()
    dst = torch.full((num_freq, num_time), -7.0, dtype=torch.float64, device="cuda")
    _lib.check(native_lib.fb_upload_rows(host.data_ptr(), host.element_size(), rows.data_ptr(), int(good.sum()),
                                         num_freq, num_time, dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
    got = dst.cpu().numpy()
    np.testing.assert_array_equal(got[good], host.numpy()[good].astype(np.float64))
    assert np.all(got[~good] == -7.0)
    # pageable memory is refused, not copied through a fallback
    pageable = np.zeros((num_freq, num_time), dtype=dtype)
    rc = native_lib.fb_upload_rows(pageable.ctypes.data, pageable.itemsize, rows.data_ptr(), int(good.sum()), num_freq,
                                   num_time, dst.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc != 0
