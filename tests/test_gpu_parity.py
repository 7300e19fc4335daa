"""
Parity of the CUDA path (through the SpectrumModeler / LSFitter mirror, i.e. through the C ABI)
against the CPU oracle on seeded inputs. Tolerances are BASELINE.json's: model 1e-10 relative,
best-fit parameters 1e-6 relative.
"""
import numpy as np
import pytest

from fitburst_b200 import synthetic as syn
from tests.helpers import assert_close, make_case, model_pair, rel_to_max

pytestmark = pytest.mark.gpu

MODEL_CASES = {
    "config2_shape": dict(),
    "no_upsampling": dict(kf=1, kt=1),
    "odd_factors": dict(kf=3, kt=5, num_comp=2),
    "gaussian": dict(tau=0.0, kf=4, kt=2),
    "folded": dict(folded=True, kf=2, kt=3, num_comp=2),
    "folded_gaussian": dict(folded=True, tau=0.0, kf=2, kt=2, num_comp=1),
    "dedispersed_offset": dict(dm=0.3, kf=4, kt=2),
    "incoherent_dm": dict(dm_inc=50.0, dm=0.05, kf=4, kt=2, num_comp=1),
    "ragged_time": dict(num_time=301, num_freq=19, kf=2, kt=2),
    "eight_components": dict(num_comp=3, kf=1, kt=1, num_time=640),
}


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_model_and_caches(name, native_lib):
    freqs, times, truth, kwargs = make_case(**MODEL_CASES[name])
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    got = cuda.compute_model()
    want = oracle.compute_model()
    assert_close(got, want, 1e-10, f"{name} model")
    assert_close(cuda.spectrum_per_component, oracle.spectrum_per_component, 1e-10, f"{name} spectrum")
    assert_close(cuda.timeprof_per_component, oracle.timeprof_per_component, 1e-10, f"{name} timeprof")
    assert_close(cuda.amplitude_per_component, oracle.amplitude_per_component, 1e-10, f"{name} amplitude")
    # timediff is a difference of O(second) terms: absolute 1e-15 s is its rounding floor
    assert np.max(np.abs(cuda.timediff_per_component - oracle.timediff_per_component)) < 1e-13


def test_normalize_pbf_flag(native_lib):
    from fitburst_b200.backend import general
    freqs, times, truth, kwargs = make_case(kf=2, kt=2)
    general["flags"]["normalize_pbf"] = True
    try:
        cuda, oracle = model_pair(freqs, times, kwargs, truth)
        oracle.normalize_pbf = True
        assert_close(cuda.compute_model(), oracle.compute_model(), 1e-10, "normalize_pbf")
    finally:
        general["flags"]["normalize_pbf"] = False


def _fitters(case, fixed, bad_fraction=0.3, sigma=0.1, seed=2, weighted=True, start=None):
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter

    freqs, times, truth, kwargs = make_case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    clean = oracle.compute_model() if not kwargs["scintillation"] else None
    if clean is None:
        kw2 = dict(kwargs, scintillation=False)
        _, plain = model_pair(freqs, times, kw2, truth)
        clean = plain.compute_model()
    good = syn.make_good_freq(len(freqs), bad_fraction, seed)
    data = syn.add_noise(clean, good, sigma, seed)
    begin = start if start is not None else syn.config2_start(truth)
    cuda.update_parameters(begin)
    oracle.update_parameters(begin)
    fc = LSFitter(data, cuda, good, weighted_fit=weighted)
    fo = OracleFitter(data, oracle, good, weighted_fit=weighted)
    fc.fix_parameter(list(fixed))
    fo.fix_parameter(list(fixed))
    return fc, fo


def test_weights(native_lib):
    fc, fo = _fitters(dict(kf=2, kt=2), ["dm_index", "scattering_index"])
    np.testing.assert_allclose(fc.weights, fo.weights, rtol=1e-13, atol=0)
    assert np.all(fc.weights[~fo.good_freq] == 0.0)


@pytest.mark.parametrize("case,fixed", [
    (dict(), ["dm_index", "scattering_index"]),
    (dict(kf=1, kt=1), []),
    (dict(tau=0.0, kf=2, kt=2), ["dm_index", "scattering_index", "scattering_timescale"]),
    (dict(dm=0.2, kf=2, kt=2, num_comp=2), ["dm_index", "scattering_index"]),
])
def test_residuals_jacobian_and_normal_equations(case, fixed, native_lib):
    import ctypes, torch
    from fitburst_b200 import _lib
    fc, fo = _fitters(case, fixed)
    x0 = fo.get_fit_parameters_list()
    assert fc.get_fit_parameters_list() == x0
    r_c, r_o = fc.compute_residuals(x0), fo.compute_residuals(x0)
    assert rel_to_max(r_c, r_o) < 1e-12  # data - model cancels: element-wise relative error is meaningless
    j_c, j_o = fc.compute_jacobian(x0), fo.compute_jacobian(x0)
    assert j_c.shape == j_o.shape
    for col in range(j_o.shape[1]):
        assert rel_to_max(j_c[:, col], j_o[:, col]) < 1e-9, f"jacobian column {col}"
    # fused normal equations against the oracle's dense J and r
    plan, n = fc._prepare()
    fc.model._push_parameters()
    gram = np.zeros((n + 1, n + 1))
    _lib.check(native_lib.fb_normal_equations(plan, fc._data_on_device().data_ptr(), _lib.dptr(gram), fc._stream()))
    full = np.concatenate([j_o, r_o[:, None]], axis=1)
    want = full.T @ full
    scale = np.sqrt(np.outer(np.diag(want), np.diag(want)))
    assert np.max(np.abs(gram - want) / scale) < 1e-9


@pytest.mark.parametrize("case,fixed", [
    (dict(num_freq=64, num_time=128), ["dm_index", "scattering_index"]),
    (dict(num_freq=64, num_time=128, kf=1, kt=1, num_comp=1), ["dm_index", "scattering_index", "scattering_timescale"]),
    (dict(num_freq=64, num_time=128, tau=0.0, kf=2, kt=2, num_comp=2), ["dm_index", "scattering_index", "scattering_timescale"]),
])
def test_fit_matches_scipy_trajectory(case, fixed, native_lib):
    fc, fo = _fitters(case, fixed)
    fc.fit()
    fo.fit(with_statistics=False)
    rc, ro = fc.results, fo.results
    assert rc.status == ro.status
    assert rc.nfev == ro.nfev and rc.njev == ro.njev
    np.testing.assert_allclose(rc.x, ro.x, rtol=1e-6)
    assert abs(rc.cost - ro.cost) <= 1e-9 * abs(ro.cost)
    stats = fc.fit_statistics
    assert stats["num_fit_parameters"] == len(ro.x)
    assert abs(stats["chisq_final"] - float(np.sum(ro.fun ** 2))) <= 1e-9 * stats["chisq_final"]


def test_scintillation_model(native_lib):
    case = dict(scint=True, kf=1, kt=1, num_comp=2, num_time=160)
    freqs, times, truth, kwargs = make_case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    _, plain = model_pair(freqs, times, dict(kwargs, scintillation=False), truth)
    rng = np.random.default_rng(5)
    data = plain.compute_model() * np.exp(rng.normal(0, 0.5, size=(len(freqs), 1))) + rng.normal(0, 0.05, size=(len(freqs), len(times)))
    got = cuda.compute_model(data=data)
    want = oracle.compute_model(data=data)
    assert_close(got, want, 1e-10, "scintillation model")
    assert_close(cuda.amplitude_per_component, oracle.amplitude_per_component, 1e-10, "scintillation amplitudes")


@pytest.mark.parametrize("case", [dict(kf=2, kt=2), dict(kf=2, kt=2, tau=0.0, num_comp=2), dict(kf=1, kt=1, dm=0.2, num_comp=2)])
@pytest.mark.parametrize("normalize", [False, True])
def test_first_and_second_derivative_maps(case, normalize, native_lib):
    """All 9 first- and 45 second-derivative maps through fitburst_b200.routines.derivative."""
    import fitburst_b200.routines.derivative as deriv
    from fitburst_b200.backend import general
    from oracle.hessian_oracle import REFERENCE_PAIRS, OracleSecondDerivatives

    if normalize and case.get("tau", 1.0) == 0.0:
        pytest.skip("normalised PBF with zero scattering time divides by zero in the reference")
    general["flags"]["normalize_pbf"] = normalize
    try:
        freqs, times, truth, kwargs = make_case(num_freq=16, num_time=48, **case)
        cuda, oracle = model_pair(freqs, times, kwargs, truth)
        oracle.normalize_pbf = normalize
        cuda.compute_model()
        oracle.compute_model()
        second = OracleSecondDerivatives(oracle)
        P = oracle.get_parameters_dict()

        def close(got, want, what):
            want = np.broadcast_to(np.asarray(want, dtype=float), got.shape)
            finite = np.isfinite(want)
            if not finite.any():
                return
            scale = np.max(np.abs(want[finite]))
            err = np.max(np.abs(got[finite] - want[finite]))
            assert err <= 1e-9 * max(scale, 1e-300), f"{what}: {err / max(scale, 1e-300):.2e}"

        with np.errstate(all="ignore"):
            for name in oracle.parameters:
                for comp in range(oracle.num_components):
                    try:
                        want = second.first.first(name, P, comp)
                    except ZeroDivisionError:
                        continue
                    close(getattr(deriv, f"deriv_model_wrt_{name}")(P, cuda, comp), want, f"d/d{name}[{comp}]")
            for n1, seconds in REFERENCE_PAIRS.items():
                for n2 in seconds:
                    for comp in range(oracle.num_components):
                        try:
                            want = second.second(n1, n2, P, comp)
                        except ZeroDivisionError:
                            continue
                        got = getattr(deriv, f"deriv2_model_wrt_{n1}_{n2}")(P, cuda, comp)
                        close(got, want, f"d2/d{n1}d{n2}[{comp}]")
    finally:
        general["flags"]["normalize_pbf"] = False


@pytest.mark.parametrize("case,fixed", [
    (dict(num_freq=48, num_time=96), ["dm_index", "scattering_index"]),
    (dict(num_freq=32, num_time=64, kf=2, kt=2, num_comp=2), ["dm_index"]),
    (dict(num_freq=32, num_time=64, tau=0.0, kf=2, kt=2, num_comp=2), ["dm_index", "scattering_index", "scattering_timescale"]),
])
def test_hessian_and_fit_statistics(case, fixed, native_lib):
    from oracle import hessian_oracle
    fc, fo = _fitters(case, fixed)
    fo.hessian_fn = hessian_oracle.compute_hessian
    fc.fit()
    fo.fit()
    assert fc.results.status == fo.results.status
    np.testing.assert_allclose(fc.results.x, fo.results.x, rtol=1e-6)
    scale = np.sqrt(np.outer(np.abs(np.diag(fo.hessian)), np.abs(np.diag(fo.hessian))))
    assert np.max(np.abs(fc.hessian - fo.hessian) / scale) < 1e-6
    assert np.max(np.abs(fc.hessian_approx - fo.hessian_approx) / scale) < 1e-6
    assert fc.covariance_labels == fo.covariance_labels
    sc, so = fc.fit_statistics, fo.fit_statistics
    assert set(sc) == set(so)
    for key in ("num_freq", "num_freq_good", "num_fit_parameters", "num_observations", "num_time"):
        assert sc[key] == so[key]
    for key in ("chisq_initial", "chisq_final", "chisq_final_reduced", "snr"):
        assert abs(sc[key] - so[key]) <= 1e-8 * abs(so[key])
    for name, values in so["bestfit_uncertainties"].items():
        np.testing.assert_allclose(sc["bestfit_uncertainties"][name], values, rtol=1e-5)
        # best fit agrees far inside one sigma
        diff = np.abs(np.asarray(sc["bestfit_parameters"][name]) - np.asarray(so["bestfit_parameters"][name]))
        assert np.all(diff < 1e-3 * np.asarray(values))
    import json
    json.dumps(sc)  # the pipeline dumps fit_statistics as JSON (fitburst_pipeline.py:560-573)


@pytest.mark.parametrize("tau,folded", [(2e-3, False), (0.0, False), (2e-3, True), (0.0, True)])
def test_compute_profile_and_routines_profile(tau, folded, native_lib):
    """SpectrumModeler.compute_profile and fitburst_b200.routines.profile against the oracle's
    restatement of analysis/model.py:281-352 / routines/profile.py."""
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.routines import profile as gprof
    from oracle.fitburst_oracle import OracleModel, profile_gaussian, profile_pbf
    freqs, times = syn.chime_axes(8, 64)
    cuda = SpectrumModeler(freqs, times, num_components=1, is_folded=folded)
    oracle = OracleModel(freqs, times, num_components=1, is_folded=folded)
    sub = np.linspace(590.0, 610.0, 4)[:, None]
    tarr = (np.linspace(-0.02, 0.04, 96)[None, :] - np.linspace(0, 1e-3, 4)[:, None])
    got = cuda.compute_profile(tarr, 0.0, tau, -4.0, 1.2e-3, sub, 600.0, is_folded=folded)
    want = oracle._profile(tarr[None], tau, -4.0, 1.2e-3, sub[None], 600.0)[0]
    assert_close(got, want, 1e-10, "compute_profile")
    if not folded:
        g = gprof.compute_profile_gaussian(tarr, 1e-3, 1.2e-3)
        assert_close(g, profile_gaussian(tarr, 1e-3, 1.2e-3), 1e-12, "gaussian")
        if tau > 0:
            p = gprof.compute_profile_pbf(tarr, 0.0, 1.2e-3, sub, 600.0, tau, sc_index=-4.0)
            assert_close(p, profile_pbf(tarr, 1.2e-3, sub, 600.0, tau, -4.0, False), 1e-10, "pbf")
