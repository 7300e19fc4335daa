"""
TEST INFRASTRUCTURE ONLY -- freezes known-answer fixtures of the pipeline flow by running the
UNMODIFIED reference script fitburst/pipelines/fitburst_pipeline.py (runpy, with matplotlib /
pytz stubbed by oracle/ref_shim.py and the summary plot disabled) on synthetic generic-format
files.

    python -m oracle.make_pipeline_golden     # writes tests/golden/pipeline_*.npz / *.json

Each case stores the input .npz (the generic format of docs/usage/formatting_data_generic.md),
the command-line options and the reference's results_fitburst.json.
"""
import contextlib
import io
import json
import os
import runpy
import sys
import tempfile

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (burst description, reference command-line options)
CASES = {
    "pipeline_dedispersed_two_components": (
        dict(num_freq=48, num_time=192, is_dedispersed=True, seed=11, bad_chans=[3, 17, 30],
             params={"amplitude": [0.0, -0.2], "arrival_time": [0.075, 0.105], "burst_width": [0.0012, 0.002],
                     "dm": [349.5, 349.5], "dm_index": [-2.0, -2.0], "ref_freq": [600.0, 600.0],
                     "scattering_index": [-4.0, -4.0], "scattering_timescale": [0.0, 0.0],
                     "spectral_index": [0.5, -1.0], "spectral_running": [-2.0, 1.0]}),
        ["--iterations", "2"]),
    "pipeline_scattered_window_preprocess": (
        dict(num_freq=64, num_time=256, is_dedispersed=True, seed=12, bad_chans=[5, 6],
             params={"amplitude": [-0.1], "arrival_time": [0.12], "burst_width": [0.0015], "dm": [100.0],
                     "dm_index": [-2.0], "ref_freq": [600.0], "scattering_index": [-4.0],
                     "scattering_timescale": [0.004], "spectral_index": [1.0], "spectral_running": [-3.0]}),
        ["--fit", "scattering_timescale", "--window", "0.09", "--preprocess", "--downsample_freq", "2",
         "--upsample_freq", "2", "--upsample_time", "2", "--scattering_timescale", "0.003"]),
}


def make_input(desc, path):
    """A noisy model spectrum in the generic format (cf. pipelines/simulate_burst.py:28-87)."""
    ref_shim.import_reference()
    from fitburst.analysis.model import SpectrumModeler

    res_freq, res_time = 400.0 / desc["num_freq"], 0.00098304
    freqs_bin0 = 400.1953125
    freqs = np.arange(desc["num_freq"], dtype=np.float64) * res_freq + freqs_bin0 + res_freq / 2.0
    times = np.arange(desc["num_time"], dtype=np.float64) * res_time
    params = {key: list(value) for key, value in desc["params"].items()}
    model_params = dict(params)
    if desc["is_dedispersed"]:
        model_params["dm"] = [0.0] * len(params["dm"])
    model = SpectrumModeler(freqs, times, is_dedispersed=desc["is_dedispersed"], num_components=len(params["dm"]))
    with ref_shim.quiet():
        model.update_parameters(model_params)
        spectrum = model.compute_model()
    rng = np.random.default_rng(desc["seed"])
    spectrum = spectrum + rng.normal(0.0, 0.05, size=spectrum.shape)
    spectrum[desc["bad_chans"], :] = 0.0
    metadata = {"bad_chans": desc["bad_chans"], "freqs_bin0": freqs_bin0, "is_dedispersed": desc["is_dedispersed"],
                "num_freq": desc["num_freq"], "num_time": desc["num_time"], "times_bin0": 59000.0,
                "res_freq": res_freq, "res_time": res_time}
    np.savez(path, data_full=spectrum, metadata=metadata, burst_parameters=params)


def run_reference_pipeline(npz_path, cli_options, workdir):
    """Executes the reference script as __main__ in `workdir`; returns its results JSON."""
    fitburst = ref_shim.import_reference()
    fitburst.utilities.plotting.plot_summary_triptych = lambda *args, **kwargs: None
    argv, cwd = sys.argv, os.getcwd()
    script = os.path.join(ref_shim.REFERENCE_ROOT, "fitburst", "pipelines", "fitburst_pipeline.py")
    try:
        os.chdir(workdir)
        sys.argv = [script, npz_path] + list(cli_options)
        with contextlib.redirect_stdout(io.StringIO()):
            runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = argv
        os.chdir(cwd)
    with open(os.path.join(workdir, "results_fitburst.json")) as handle:
        return json.load(handle)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, (desc, cli_options) in CASES.items():
        npz_path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        make_input(desc, npz_path)
        with tempfile.TemporaryDirectory() as workdir:
            results = run_reference_pipeline(npz_path, cli_options, workdir)
        with open(os.path.join(GOLDEN_DIR, f"{name}.json"), "w") as handle:
            json.dump({"cli_options": cli_options, "numpy": np.__version__, "scipy": scipy.__version__,
                       "reference_results": results}, handle, indent=1)
        stats = results["fit_statistics"]
        print(f"{name}: {os.path.getsize(npz_path) // 1024} KB, chisq_final_reduced "
              f"{stats['chisq_final_reduced']:.4f}, parameters {stats['bestfit_parameters']}")


if __name__ == "__main__":
    main()
