"""
Known-answer fixtures frozen from the UNMODIFIED reference (oracle/make_golden.py, numpy 2.3.5 /
scipy 1.18.1). CPU tests pin the oracle to them; GPU tests pin the CUDA path (through the C ABI)
to them. These files are the pin that travels to the GPU box, where /root/reference is absent.
"""
import glob
import json
import os

import numpy as np
import pytest

from tests.helpers import assert_close, rel_to_max

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("pipeline_"))  # those belong to tests/test_pipeline.py
NAMES = [os.path.splitext(os.path.basename(p))[0] for p in GOLDEN]


def load(name):
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"{name}.npz"))
    out = {k: z[k] for k in z.files}
    for key in ("kwargs", "truth", "start", "fixed", "labels", "fit_statistics", "versions"):
        if key in out:
            out[key] = json.loads(str(out[key]))
    return out


def test_fixture_inventory():
    assert len(NAMES) >= 8
    for name in NAMES:
        g = load(name)
        assert g["model"].shape == (len(g["freqs"]), len(g["times"]))
        assert g["versions"]["scipy"]


# ------------------------------------------------------------------ CPU: oracle vs golden
@pytest.mark.parametrize("name", NAMES)
def test_oracle_model_matches_golden(name):
    from oracle.fitburst_oracle import OracleModel
    g = load(name)
    model = OracleModel(g["freqs"], g["times"], normalize_pbf=bool(g["normalize_pbf"]), **g["kwargs"])
    model.update_parameters(g["truth"])
    got = model.compute_model(data=g.get("scint_data"))
    rtol = 1e-11 if g["kwargs"]["dm_incoherent"] else 1e-13  # numpy SIMD-vs-scalar pow ulp, see test_oracle_vs_reference
    np.testing.assert_allclose(got, g["model"], rtol=rtol, atol=1e-300)
    for cache in ("amplitude", "spectrum", "timeprof"):
        np.testing.assert_allclose(getattr(model, f"{cache}_per_component"), g[f"cache_{cache}"], rtol=max(rtol, 1e-12), atol=1e-300)
    np.testing.assert_allclose(model.timediff_per_component, g["cache_timediff"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", [n for n in NAMES if "fit_x" in np.load(os.path.join(os.path.dirname(__file__), "golden", f"{n}.npz")).files])
def test_oracle_fit_matches_golden(name):
    from oracle import hessian_oracle
    from oracle.fitburst_oracle import OracleFitter, OracleModel
    g = load(name)
    model = OracleModel(g["freqs"], g["times"], **g["kwargs"])
    model.update_parameters(g["start"])
    fitter = OracleFitter(g["data"], model, g["good_freq"])
    fitter.fix_parameter(g["fixed"])
    fitter.hessian_fn = hessian_oracle.compute_hessian
    np.testing.assert_allclose(fitter.weights, g["weights"], rtol=1e-15)
    x0 = fitter.get_fit_parameters_list()
    np.testing.assert_allclose(x0, g["x0"], rtol=0)
    assert rel_to_max(fitter.compute_residuals(x0), g["resid0"]) < 1e-13
    jac = fitter.compute_jacobian(x0)
    for col in range(jac.shape[1]):
        assert rel_to_max(jac[:, col], g["jac0"][:, col]) < 1e-12
    fitter.fit()
    assert fitter.results.status == int(g["fit_status"]) and fitter.results.nfev == int(g["fit_nfev"])
    np.testing.assert_allclose(fitter.results.x, g["fit_x"], rtol=1e-9)
    scale = np.sqrt(np.outer(np.abs(np.diag(g["hessian"])), np.abs(np.diag(g["hessian"]))))
    assert np.max(np.abs(fitter.hessian - g["hessian"]) / scale) < 1e-9
    for key, values in g["fit_statistics"]["bestfit_uncertainties"].items():
        np.testing.assert_allclose(fitter.fit_statistics["bestfit_uncertainties"][key], values, rtol=1e-6)


# ------------------------------------------------------------------ GPU: CUDA path vs golden
@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_model_matches_golden(name, native_lib):
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.backend import general
    g = load(name)
    general["flags"]["normalize_pbf"] = bool(g["normalize_pbf"])
    try:
        model = SpectrumModeler(g["freqs"], g["times"], **g["kwargs"])
        model.update_parameters(g["truth"])
        got = model.compute_model(data=g.get("scint_data"))
        rtol = 1e-10  # BASELINE.json: model spectra within 1e-10 relative, scintillation mode included
        assert_close(got, g["model"], rtol, f"{name} model")
        for cache in ("amplitude", "spectrum", "timeprof"):
            assert_close(getattr(model, f"{cache}_per_component"), g[f"cache_{cache}"], rtol, f"{name} {cache}")
        assert np.max(np.abs(model.timediff_per_component - g["cache_timediff"])) < 1e-13
    finally:
        general["flags"]["normalize_pbf"] = False


@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in NAMES if "fit_x" in np.load(os.path.join(os.path.dirname(__file__), "golden", f"{n}.npz")).files])
def test_cuda_fit_matches_golden(name, native_lib):
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    g = load(name)
    model = SpectrumModeler(g["freqs"], g["times"], **g["kwargs"])
    model.update_parameters(g["start"])
    fitter = LSFitter(g["data"], model, g["good_freq"], weighted_fit=True)
    fitter.fix_parameter(g["fixed"])
    np.testing.assert_allclose(fitter.weights, g["weights"], rtol=1e-13)
    x0 = fitter.get_fit_parameters_list()
    assert rel_to_max(fitter.compute_residuals(x0), g["resid0"]) < 1e-12
    jac = fitter.compute_jacobian(x0)
    for col in range(jac.shape[1]):
        assert rel_to_max(jac[:, col], g["jac0"][:, col]) < 1e-9, f"column {col}"
    fitter.fit()
    res = fitter.results
    assert res.status == int(g["fit_status"])
    assert res.nfev == int(g["fit_nfev"]) and res.njev == int(g["fit_njev"])
    np.testing.assert_allclose(res.x, g["fit_x"], rtol=1e-6)  # BASELINE.json: best fit within 1e-6 relative
    assert abs(res.cost - float(g["fit_cost"])) <= 1e-9 * float(g["fit_cost"])
    scale = np.sqrt(np.outer(np.abs(np.diag(g["hessian"])), np.abs(np.diag(g["hessian"]))))
    assert np.max(np.abs(fitter.hessian - g["hessian"]) / scale) < 1e-6
    assert fitter.covariance_labels == g["labels"]
    want = g["fit_statistics"]
    for key in ("num_freq", "num_freq_good", "num_fit_parameters", "num_observations", "num_time"):
        assert fitter.fit_statistics[key] == want[key]
    for key in ("chisq_initial", "chisq_final", "chisq_final_reduced", "snr"):
        assert abs(fitter.fit_statistics[key] - want[key]) <= 1e-8 * abs(want[key])
    for key, values in want["bestfit_uncertainties"].items():
        np.testing.assert_allclose(fitter.fit_statistics["bestfit_uncertainties"][key], values, rtol=1e-5)
        diff = np.abs(np.asarray(fitter.fit_statistics["bestfit_parameters"][key]) - np.asarray(want["bestfit_parameters"][key]))
        assert np.all(diff < 1e-3 * np.asarray(values))  # well inside one sigma
    # dense members stay available lazily (results.jac feeds hessian_approx in the reference)
    assert res.fun.shape == g["resid0"].shape and res.jac.shape == g["jac0"].shape
