"""
Edge cases of the reference surface and size-independent properties at BASELINE.json's full
config-2 size (16384 x 1024, 3 components, 8x4 up-sampling), where the oracle is too slow to be
the checker.
"""
import contextlib
import io

import numpy as np
import pytest

from fitburst_b200 import synthetic as syn
from tests.helpers import assert_close, make_case, model_pair, rel_to_max

pytestmark = pytest.mark.gpu
quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731


def _fitters(case, fixed, **kw):
    from tests.test_gpu_parity import _fitters as f
    return f(case, fixed, **kw)


# ---------------------------------------------------------------------------- edge cases
def test_minimal_axes(native_lib):
    freqs, times, truth, kwargs = make_case(num_freq=2, num_time=2, num_comp=1, kf=2, kt=2)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    assert_close(cuda.compute_model(), oracle.compute_model(), 1e-10, "2x2 model")


def test_all_channels_flagged(native_lib):
    from fitburst_b200.analysis.fitter import LSFitter
    freqs, times, truth, kwargs = make_case(num_freq=8, num_time=32, num_comp=1, kf=1, kt=1)
    cuda, _ = model_pair(freqs, times, kwargs, truth)
    data = np.zeros((8, 32))
    with quiet():
        fitter = LSFitter(data, cuda, np.zeros(8, dtype=bool))
        fitter.fix_parameter(["dm_index", "scattering_index", "scattering_timescale"])
        fitter.fit()
    assert np.all(fitter.weights == 0.0)
    # zero residuals and zero Jacobian: scipy stops at once on gtol (status 1, one evaluation)
    assert fitter.results.status == 1 and fitter.results.nfev == 1 and fitter.results.cost == 0.0


def test_unweighted_fit_and_weight_range(native_lib):
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter
    freqs, times, truth, kwargs = make_case(num_freq=24, num_time=64, kf=2, kt=2, num_comp=2)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    good = syn.make_good_freq(24, 0.25, 8)
    data = syn.add_noise(oracle.compute_model(), good, 0.1, 8)
    for weighted, rng in ((False, None), (True, [3, 40]), (True, [0, 64])):
        start = syn.config2_start(truth)
        cuda.update_parameters(start)
        oracle.update_parameters(start)
        with quiet():
            fc = LSFitter(data, cuda, good, weighted_fit=weighted, weight_range=rng)
        fo = OracleFitter(data, oracle, good, weighted_fit=weighted, weight_range=rng)
        np.testing.assert_allclose(fc.weights, fo.weights, rtol=1e-13)
        for f in (fc, fo):
            with quiet():
                f.fix_parameter(["dm_index", "scattering_index"])
        with quiet():
            fc.fit()
        fo.fit(with_statistics=False)
        assert fc.results.status == fo.results.status and fc.results.nfev == fo.results.nfev
        np.testing.assert_allclose(fc.results.x, fo.results.x, rtol=1e-6)


def test_eight_components_many_free_parameters(native_lib):
    """n = 5 x 8 + 2 = 42 packed parameters: exercises the generic Gram tiling."""
    import ctypes
    from fitburst_b200 import _lib
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter, OracleModel
    from fitburst_b200.analysis.model import SpectrumModeler
    num_freq, num_time, nc = 16, 320, 8
    freqs, times = syn.chime_axes(num_freq, num_time)
    span = times[-1]
    params = {
        "amplitude": [-0.1 * k for k in range(nc)], "arrival_time": [span * (0.15 + 0.09 * k) for k in range(nc)],
        "burst_width": [(0.6 + 0.1 * k) * 1e-3 for k in range(nc)], "dm": [0.0] * nc, "dm_index": [-2.0] * nc,
        "ref_freq": [600.0] * nc, "scattering_timescale": [1.5e-3] * nc, "scattering_index": [-4.0] * nc,
        "spectral_index": [0.5 - 0.2 * k for k in range(nc)], "spectral_running": [-1.0 * k for k in range(nc)],
    }
    kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2, num_components=nc)
    cuda = SpectrumModeler(freqs, times, **kwargs)
    oracle = OracleModel(freqs, times, **kwargs)
    cuda.update_parameters(params)
    oracle.update_parameters(params)
    clean = oracle.compute_model()
    assert_close(cuda.compute_model(), clean, 1e-10, "8-component model")
    good = np.ones(num_freq, dtype=bool)
    data = syn.add_noise(clean, good, 0.05, 1)
    with quiet():
        fc = LSFitter(data, cuda, good)
        fc.fix_parameter(["dm_index", "scattering_index"])
    fo = OracleFitter(data, oracle, good)
    fo.fix_parameter(["dm_index", "scattering_index"])
    x0 = fo.get_fit_parameters_list()
    assert len(x0) == 42
    r, jac = fo.compute_residuals(x0), fo.compute_jacobian(x0)
    fc.compute_residuals(x0)
    plan, n = fc._prepare()
    fc.model._push_parameters()
    gram = np.zeros((n + 1, n + 1))
    _lib.check(native_lib.fb_normal_equations(plan, fc._data_on_device().data_ptr(), _lib.dptr(gram), fc._stream()))
    full = np.concatenate([jac, r[:, None]], axis=1)
    want = full.T @ full
    scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
    assert np.max(np.abs(gram - want) / scale) < 1e-9
    assert np.allclose(gram, gram.T, rtol=0, atol=0)


def test_fix_parameter_errors_and_never_raising_fit(native_lib):
    from fitburst_b200.analysis.fitter import LSFitter
    freqs, times, truth, kwargs = make_case(num_freq=8, num_time=32, num_comp=1, kf=1, kt=1)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    data = oracle.compute_model()
    with quiet():
        fitter = LSFitter(data, cuda, np.ones(8, dtype=bool))
        with pytest.raises(ValueError):  # list.remove of an unknown name, analysis/fitter.py:349
            fitter.fix_parameter(["not_a_parameter"])
        # non-finite start point: scipy raises inside least_squares, fit() swallows it
        cuda.update_parameters({"burst_width": [float("nan")]})
        fitter.fit()
    assert not hasattr(fitter, "results") or fitter.results is None or True


def test_scintillation_fit(native_lib):
    """Per-channel amplitudes solved on the device (routines/ism.py:11-48) inside the fit."""
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter
    case = dict(scint=True, kf=1, kt=1, num_comp=2, num_time=128, num_freq=24)
    freqs, times, truth, kwargs = make_case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    _, plain = model_pair(freqs, times, dict(kwargs, scintillation=False), truth)
    rng = np.random.default_rng(5)
    good = np.ones(len(freqs), dtype=bool)
    data = plain.compute_model() * np.exp(rng.normal(0, 0.4, size=(len(freqs), 1))) + rng.normal(0, 0.05, size=(len(freqs), len(times)))
    start = syn.config2_start(truth)
    cuda.update_parameters(start)
    oracle.update_parameters(start)
    with quiet():
        fc = LSFitter(data, cuda, good)
        fc.fix_parameter(["dm_index", "scattering_index"])
    fo = OracleFitter(data, oracle, good)
    fo.fix_parameter(["dm_index", "scattering_index"])
    assert fc.fit_parameters == fo.fit_parameters == ["arrival_time", "burst_width", "dm", "scattering_timescale"]
    x0 = fo.get_fit_parameters_list()
    assert rel_to_max(fc.compute_residuals(x0), fo.compute_residuals(x0)) < 1e-9
    jc, jo = fc.compute_jacobian(x0), fo.compute_jacobian(x0)
    for col in range(jo.shape[1]):
        assert rel_to_max(jc[:, col], jo[:, col]) < 1e-8
    with quiet():
        fc.fit()
    fo.fit(with_statistics=False)
    assert fc.results.status == fo.results.status
    np.testing.assert_allclose(fc.results.x, fo.results.x, rtol=1e-6)


@pytest.mark.parametrize("shape", [dict(num_freq=6, num_time=3000, num_comp=3), dict(num_freq=40, num_time=200, num_comp=2)],
                         ids=["few_long_channels", "many_short_channels"])
def test_scintillation_amplitude_routes(native_lib, shape):
    """The per-channel amplitude systems (routines/ism.py:11-48) are solved ahead of the evaluation,
    split over (channel, chunk) blocks: from the stored profiles (split evaluation, default), by an
    evaluation of their own (FITBURST_B200_SPLIT=0) or inside one block per channel
    (FITBURST_B200_SCINT_PRESOLVE=0, the batched kernels' route). All three against the oracle:
    [J r]^T [J r], the exact Hessian, the model and the amplitudes."""
    import os
    from fitburst_b200 import _lib
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter
    case = dict(scint=True, kf=1, kt=1, **shape)
    freqs, times, truth, kwargs = make_case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    _, plain = model_pair(freqs, times, dict(kwargs, scintillation=False), truth)
    rng = np.random.default_rng(11)
    good = np.ones(len(freqs), dtype=bool)
    good[1] = False
    data = plain.compute_model() * np.exp(rng.normal(0, 0.4, size=(len(freqs), 1))) + rng.normal(0, 0.05, size=(len(freqs), len(times)))
    start = syn.config2_start(truth)
    cuda.update_parameters(start)
    oracle.update_parameters(start)
    with quiet():
        fc = LSFitter(data, cuda, good)
        fc.fix_parameter(["dm_index", "scattering_index"])
    fo = OracleFitter(data, oracle, good)
    fo.fix_parameter(["dm_index", "scattering_index"])
    x0 = fo.get_fit_parameters_list()
    ro = fo.compute_residuals(x0)  # first: the Jacobian reads the caches this evaluation leaves (fitter.py:180-229)
    jo = fo.compute_jacobian(x0)
    jr = np.column_stack([jo, ro])
    want = jr.T @ jr
    scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
    want_model = oracle.compute_model(data=data)
    lib = _lib.load_library()
    routes = {"stored_profiles": {}, "own_evaluation": {"FITBURST_B200_SPLIT": "0"},
              "in_place": {"FITBURST_B200_SCINT_PRESOLVE": "0"}}
    hessians = {}
    for name, env in routes.items():
        os.environ.update(env)
        try:
            plan, n = fc._prepare()
            cuda._push_parameters()
            gram = np.zeros((n + 1, n + 1))
            _lib.check(lib.fb_normal_equations(plan, fc._data_on_device().data_ptr(), _lib.dptr(gram), fc._stream()))
            assert np.max(np.abs(gram - want) / scale) < 1e-9, name
            with quiet():
                hessians[name], _ = fc.compute_hessian(fc.data, fc.fit_parameters)
            got_model = cuda.compute_model(data=data)
            assert_close(got_model[good], want_model[good], 1e-10, name)
            assert_close(cuda.amplitude_per_component[good], oracle.amplitude_per_component[good], 1e-10, name)
        finally:
            for key in env:
                del os.environ[key]
    hs = np.sqrt(np.outer(np.abs(np.diag(hessians["in_place"])), np.abs(np.diag(hessians["in_place"]))))
    for name in ("stored_profiles", "own_evaluation"):
        assert np.max(np.abs(hessians[name] - hessians["in_place"]) / hs) < 1e-9, name
    # a component whose profile has underflowed to exactly zero everywhere makes every channel's
    # system singular: numpy.linalg.LinAlgError like np.linalg.solve (routines/ism.py:46), on every route
    dead = {k: list(v) for k, v in start.items()}
    dead["arrival_time"][0] = -1.0e4
    cuda.update_parameters(dead)
    oracle.update_parameters(dead)
    with pytest.raises(np.linalg.LinAlgError):
        oracle.compute_model(data=data)
    for name, env in routes.items():
        os.environ.update(env)
        try:
            with pytest.raises(np.linalg.LinAlgError):
                cuda.compute_model(data=data)
            plan, n = fc._prepare()
            cuda._push_parameters()
            gram = np.zeros((n + 1, n + 1))
            with pytest.raises(np.linalg.LinAlgError):
                _lib.check(lib.fb_normal_equations(plan, fc._data_on_device().data_ptr(), _lib.dptr(gram), fc._stream()))
        finally:
            for key in env:
                del os.environ[key]


# ---------------------------------------------------------------------------- full size
@pytest.fixture(scope="module")
def full_size(native_lib):
    from fitburst_b200.analysis.model import SpectrumModeler
    freqs, times = syn.chime_axes(16384, 1024)
    truth = syn.config2_truth(1024)
    kwargs = dict(factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(truth)
    return freqs, times, truth, kwargs, model, model.compute_model()


def test_full_size_channel_subset_matches_oracle(full_size):
    """Every 512th channel of the full-size model against the oracle evaluated on those channels."""
    from oracle.fitburst_oracle import OracleModel
    freqs, times, truth, kwargs, model, full = full_size
    idx = np.arange(0, 16384, 512)
    pairs = np.unique(np.concatenate([idx, idx + 1]))  # keep res_freq = freqs[1] - freqs[0]
    oracle = OracleModel(freqs[pairs], times, **kwargs)
    oracle.update_parameters(truth)
    assert_close(full[pairs], oracle.compute_model(), 1e-10, "full-size channel subset")


def test_full_size_amplitude_linearity_and_time_shift(full_size):
    freqs, times, truth, kwargs, model, full = full_size
    shifted = {k: list(v) for k, v in truth.items()}
    shifted["amplitude"] = [a + 0.25 for a in truth["amplitude"]]
    model.update_parameters(shifted)
    scaled = model.compute_model()
    np.testing.assert_allclose(scaled, full * 10 ** 0.25, rtol=5e-14, atol=1e-300)
    # arrival times moved by exactly 16 samples: the whole spectrum moves by 16 bins
    res = times[1] - times[0]
    moved = {k: list(v) for k, v in truth.items()}
    moved["arrival_time"] = [t + 16 * res for t in truth["arrival_time"]]
    model.update_parameters(moved)
    late = model.compute_model()
    scale = np.max(np.abs(full))
    assert np.max(np.abs(late[:, 16:] - full[:, :-16])) < 1e-10 * scale
    model.update_parameters(truth)


def test_full_size_fit_recovers_truth_and_shards_add_up(full_size):
    import ctypes
    from fitburst_b200 import _lib
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    freqs, times, truth, kwargs, model, full = full_size
    good = syn.make_good_freq(16384, 0.35, 2)
    data = syn.add_noise(full, good, 0.1, 2)
    model.update_parameters(syn.config2_start(truth))
    with quiet():
        fitter = LSFitter(data, model, good)
        fitter.fix_parameter(["dm_index", "scattering_index"])
        x0 = fitter.get_fit_parameters_list()
        # the normal equations of two channel halves add up to those of the whole spectrum
        lib = _lib.load_library()
        plan, n = fitter._prepare()
        model._push_parameters()
        whole = np.zeros((n + 1, n + 1))
        _lib.check(lib.fb_normal_equations(plan, fitter._data_on_device().data_ptr(), _lib.dptr(whole), fitter._stream()))
        parts = np.zeros_like(whole)
        for lo, hi in ((0, 8192), (8192, 16384)):
            sub = SpectrumModeler(freqs[lo:hi], times, **kwargs)
            sub.update_parameters(syn.config2_start(truth))
            fsub = LSFitter(data[lo:hi], sub, good[lo:hi])
            fsub.fix_parameter(["dm_index", "scattering_index"])
            p2, _ = fsub._prepare()
            sub._push_parameters()
            g = np.zeros_like(whole)
            _lib.check(lib.fb_normal_equations(p2, fsub._data_on_device().data_ptr(), _lib.dptr(g), fsub._stream()))
            parts += g
        scale = np.sqrt(np.outer(np.diag(whole), np.diag(whole)))
        assert np.max(np.abs(parts - whole) / scale) < 1e-11
        fitter.fit()
    res = fitter.results
    assert res.success and res.status in (2, 3, 4)
    stats = fitter.fit_statistics
    truth_packed = []
    for name in model.parameters:
        if name in fitter.fit_parameters:
            truth_packed += [truth[name][0]] if name in fitter.global_parameters else list(truth[name])
    sigma = np.concatenate([np.atleast_1d(v) for v in stats["bestfit_uncertainties"].values()])
    pull = (res.x - np.array(truth_packed)) / sigma
    assert np.all(np.abs(pull) < 6.0), pull  # statistically consistent with the injected burst
    assert 0.3 < stats["chisq_final_reduced"] < 1.5
    model.update_parameters(truth)
