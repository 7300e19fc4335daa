"""
TEST INFRASTRUCTURE ONLY -- freezes known-answer fixtures by running the UNMODIFIED reference
(/root/reference, via oracle/ref_shim.py) on seeded synthetic inputs.

    python -m oracle.make_golden          # writes tests/golden/*.npz

The reference ships no golden vectors for this path; these files are the
pin that travels to the GPU box. Each file records the numpy/scipy versions it was made with.
"""
import json
import os
import sys

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fitburst_b200 import synthetic as syn  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (model-construction case, parameters fixed in the fit, extra options)
CASES = {
    "config2_small": (dict(num_freq=24, num_time=64, num_comp=3, kf=8, kt=4), ["dm_index", "scattering_index"], {}),
    "plain_one_component": (dict(num_freq=20, num_time=48, num_comp=1, kf=1, kt=1),
                            ["dm_index", "scattering_index", "scattering_timescale"], dict(tau=0.0)),
    "gaussian_two_components": (dict(num_freq=20, num_time=64, num_comp=2, kf=2, kt=2),
                                ["dm_index", "scattering_index", "scattering_timescale"], dict(tau=0.0)),
    "scattering_index_free": (dict(num_freq=24, num_time=64, num_comp=2, kf=2, kt=2), ["dm_index"], {}),
    "folded": (dict(num_freq=12, num_time=48, num_comp=2, kf=2, kt=3, folded=True), None, {}),
    "incoherent_dm": (dict(num_freq=12, num_time=48, num_comp=1, kf=4, kt=2, dm_inc=50.0, dm=0.05), None, {}),
    "normalized_pbf": (dict(num_freq=12, num_time=48, num_comp=2, kf=2, kt=2), None, dict(normalize=True)),
    "scintillation": (dict(num_freq=12, num_time=96, num_comp=2, kf=1, kt=1, scint=True), None, {}),
    # BASELINE.json configs[0] stand-in: every 341st of the 16384 channels x 162 samples, Gaussian
    # component, pipeline-default fit (n = 6)
    "config1_tutorial": (dict(config1=True, stride=341), list(syn.CONFIG1_FIXED), {}),
}


def case_inputs(case, tau=2e-3, dm=0.0):
    if case.get("config1"):
        # the tutorial stand-in on a decimated channel axis
        freqs_full, times = syn.chime_axes(syn.CONFIG1_NUM_FREQ, syn.CONFIG1_NUM_TIME)
        freqs = freqs_full[syn.channel_subset(syn.CONFIG1_NUM_FREQ, case["stride"])]
        kwargs = dict(dm_incoherent=0.0, factor_freq_upsample=1, factor_time_upsample=1, num_components=1,
                      is_folded=False, scintillation=False)
        return freqs, times, syn.config1_truth(), kwargs
    freqs, times = syn.chime_axes(case["num_freq"], case["num_time"])
    truth = syn.truncate_components(syn.config2_truth(case["num_time"]), case["num_comp"])
    truth["scattering_timescale"] = [case.get("tau", tau)] * case["num_comp"]
    truth["dm"] = [case.get("dm", dm)] * case["num_comp"]
    kwargs = dict(dm_incoherent=case.get("dm_inc", 0.0), factor_freq_upsample=case["kf"],
                  factor_time_upsample=case["kt"], num_components=case["num_comp"],
                  is_folded=case.get("folded", False), scintillation=case.get("scint", False))
    return freqs, times, truth, kwargs


def make_one(name):
    ref_shim.import_reference()
    from fitburst.analysis.fitter import LSFitter
    from fitburst.analysis.model import SpectrumModeler
    from fitburst.backend import general

    case, fixed, extra = CASES[name]
    case = dict(case, **{k: v for k, v in extra.items() if k in ("tau",)})
    general["flags"]["normalize_pbf"] = bool(extra.get("normalize", False))
    freqs, times, truth, kwargs = case_inputs(case)
    out = {"freqs": freqs, "times": times, "kwargs": json.dumps(kwargs), "truth": json.dumps(truth),
           "normalize_pbf": np.array(bool(extra.get("normalize", False))),
           "versions": json.dumps({"numpy": np.__version__, "scipy": scipy.__version__})}
    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(truth)
    data = None
    if kwargs["scintillation"]:
        plain = SpectrumModeler(freqs, times, **dict(kwargs, scintillation=False))
        plain.update_parameters(truth)
        with ref_shim.quiet():
            clean = plain.compute_model()
        rng = np.random.default_rng(11)
        data = clean * np.exp(rng.normal(0, 0.5, size=(len(freqs), 1))) + rng.normal(0, 0.05, size=clean.shape)
        out["scint_data"] = data
    with ref_shim.quiet():
        out["model"] = model.compute_model(data=data)
    for cache in ("amplitude", "spectrum", "timediff", "timeprof"):
        out[f"cache_{cache}"] = getattr(model, f"{cache}_per_component").copy()
    if fixed is not None:
        if case.get("config1"):
            good = syn.config1_good_freq(syn.CONFIG1_NUM_FREQ)[syn.channel_subset(syn.CONFIG1_NUM_FREQ, case["stride"])]
            data, _ = syn.config1_data(out["model"], good)
            start = syn.config1_start(truth)
        else:
            good = syn.make_good_freq(len(freqs), 0.3, 2)
            data = syn.add_noise(out["model"], good, 0.1, 2)
            start = syn.config2_start(truth)
        model.update_parameters(start)
        with ref_shim.quiet():
            fitter = LSFitter(data, model, good, weighted_fit=True)
            fitter.fix_parameter(list(fixed))
            x0 = fitter.get_fit_parameters_list()
            resid0 = fitter.compute_residuals(x0)
            jac0 = fitter.compute_jacobian(x0)
            fitter.fit()
        res = fitter.results
        out.update({
            "good_freq": good, "data": data, "start": json.dumps(start), "fixed": json.dumps(list(fixed)),
            "weights": fitter.weights, "x0": np.array(x0), "resid0": resid0, "jac0": jac0,
            "fit_x": res.x, "fit_cost": np.array(res.cost), "fit_grad": res.grad,
            "fit_status": np.array(res.status), "fit_nfev": np.array(res.nfev), "fit_njev": np.array(res.njev),
            "hessian": fitter.hessian, "hessian_approx": fitter.hessian_approx,
            "covariance": fitter.covariance, "labels": json.dumps(fitter.covariance_labels),
            "fit_statistics": json.dumps(fitter.fit_statistics),
        })
    general["flags"]["normalize_pbf"] = False
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = sys.argv[1:] or list(CASES)  # `python oracle/make_golden.py config1_tutorial` adds one case
    for name in names:
        out = make_one(name)
        size = os.path.getsize(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        print(f"{name}: {size / 1024:.0f} KiB, keys={len(out)}")


if __name__ == "__main__":
    main()
