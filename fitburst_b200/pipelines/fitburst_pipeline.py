"""
The flow of fitburst/pipelines/fitburst_pipeline.py (reference lines 298-573: read ->
preprocess -> downsample -> initial guess -> de-disperse / window -> optional FindPeak -> model ->
fit iterations -> results JSON) as a function over this package's device classes, and a
multi-file front end: `run_many()` reads and prepares file k+1 on a host thread while file k is
being fitted, so that .npz parsing, page-locked staging and the upload overlap the fits.

    python -m fitburst_b200.pipelines.fitburst_pipeline FILE [FILE ...] [the reference's options]

Same option names, defaults and output (`results_fitburst*.json` with the keys of
fitburst_pipeline.py:560-573) as the reference script; the summary plot is not produced
(plotting is out of scope).
"""
import argparse
import concurrent.futures
import contextlib
import io
import json
import os
from copy import deepcopy

import numpy as np

from ..analysis.fitter import LSFitter
from ..analysis.model import SpectrumModeler
from ..analysis.peak_finder import FindPeak
from ..backend.generic import DataReader
from .._say import quiet as _quiet, say

DEFAULTS = dict(
    amplitude=None, arrival_time=None, dm=None, dm_index=None, factor_freq_downsample=1,
    factor_time_downsample=1, parameters_to_fit=[], parameters_to_fix=[], is_folded=False, num_iterations=1,
    use_outfile_substring=False, peakfind_dist=5, peakfind_rms=None, preprocess_data=False, ref_freq=None,
    remove_dispersion_smearing=False, scattering_timescale=None, scintillation=False, spectral_index=None,
    spectral_running=None, solution_file=None, factor_freq_upsample=1, factor_time_upsample=1,
    variance_range=[0.2, 0.8], verbose=False, weight_range=None, width=None, window=None,
)


def build_parser():
    """The reference's command line (fitburst_pipeline.py:21-296), `file` accepting several paths."""
    parser = argparse.ArgumentParser(description=(
        "Reads, preprocesses and fits 'generic'-formatted data against a model of the dynamic spectrum on "
        "the GPU. NOTE: by default, the scattering timescale is not a fit parameter but can be turned into "
        "one through use of the --fit option."))
    parser.add_argument("file", nargs="+", type=str)
    floats = dict(action="store", default=None, nargs="+", type=float)
    parser.add_argument("--amplitude", dest="amplitude", **floats)
    parser.add_argument("--arrival_time", dest="arrival_time", **floats)
    parser.add_argument("--dm", dest="dm", **floats)
    parser.add_argument("--dm_index", dest="dm_index", **floats)
    parser.add_argument("--downsample_freq", dest="factor_freq_downsample", default=1, type=int)
    parser.add_argument("--downsample_time", dest="factor_time_downsample", default=1, type=int)
    parser.add_argument("--fit", dest="parameters_to_fit", default=[], nargs="+", type=str)
    parser.add_argument("--fix", dest="parameters_to_fix", default=[], nargs="+", type=str)
    parser.add_argument("--folded", action="store_true", dest="is_folded", default=False)
    parser.add_argument("--iterations", dest="num_iterations", default=1, type=int)
    parser.add_argument("--outfile", action="store_true", dest="use_outfile_substring", default=False)
    parser.add_argument("--peakfind_dist", dest="peakfind_dist", default=5, type=int)
    parser.add_argument("--peakfind_rms", dest="peakfind_rms", default=None, type=float)
    parser.add_argument("--preprocess", action="store_true", dest="preprocess_data")
    parser.add_argument("--ref_freq", dest="ref_freq", default=None, type=float)
    parser.add_argument("--remove_smearing", action="store_true", dest="remove_dispersion_smearing")
    parser.add_argument("--scattering_timescale", dest="scattering_timescale", **floats)
    parser.add_argument("--scintillation", action="store_true", dest="scintillation")
    parser.add_argument("--spectral_index", dest="spectral_index", **floats)
    parser.add_argument("--spectral_running", dest="spectral_running", **floats)
    parser.add_argument("--solution", dest="solution_file", default=None, type=str)
    parser.add_argument("--upsample_freq", dest="factor_freq_upsample", default=1, type=int)
    parser.add_argument("--upsample_time", dest="factor_time_upsample", default=1, type=int)
    parser.add_argument("--variance_range", dest="variance_range", default=[0.2, 0.8], nargs=2, type=float)
    parser.add_argument("--verbose", action="store_true", dest="verbose", default=False)
    parser.add_argument("--weight_range", dest="weight_range", default=None, nargs=2, type=int)
    parser.add_argument("--width", dest="width", **floats)
    parser.add_argument("--window", dest="window", default=None, type=float)
    return parser


def _options(overrides):
    options = deepcopy(DEFAULTS)
    unknown = set(overrides) - set(options)
    if unknown:
        raise TypeError(f"unknown pipeline option(s): {sorted(unknown)}")
    options.update(deepcopy(overrides))
    return options


def _fixed_parameters(options):
    """fitburst_pipeline.py:331-335."""
    parameters_to_fix = list(options["parameters_to_fix"]) + ["dm_index", "scattering_index", "scattering_timescale"]
    for current_fit_parameter in options["parameters_to_fit"]:
        if current_fit_parameter in parameters_to_fix:
            parameters_to_fix.remove(current_fit_parameter)
    return parameters_to_fix


def prepare(input_file: str, device=None, **overrides) -> dict:
    """
    Everything of the reference script up to the model construction (fitburst_pipeline.py:337-490):
    host-side parsing plus the device-side preprocessing / down-sampling / windowing. Returns the
    state `fit_prepared` needs; safe to run on a worker thread while another burst is being fitted.
    """
    opt = _options(overrides)
    existing_results = None
    if opt["solution_file"] is not None and os.path.isfile(opt["solution_file"]):
        try:
            with open(opt["solution_file"], "r") as handle:
                existing_results = json.load(handle)
        except Exception:  # noqa: BLE001 - like the reference's bare except
            say("WARNING: input solution cannot be read; using basic initial parameters...")
    else:
        say("INFO: no solution file found or provided; proceeding with fit...")

    outfile_substring = ""
    if opt["use_outfile_substring"]:
        # the reference splits the path as given (fitburst_pipeline.py:353-355), which only works
        # for files in the working directory; the base name is used here
        elems = os.path.basename(input_file).split(".")
        outfile_substring = "_" + ".".join(elems[0:len(elems) - 1])

    data = DataReader(input_file, device=device)
    data.load_data()
    say(f"INFO: there are {data.num_freq} frequencies and {data.num_time} time samples.")

    # fitburst_pipeline.py:364-372 (the second assignment wins, like in the reference)
    data.good_freq = np.sum(data.data_weights, axis=1) != 0.
    data_full = data.data_full
    data.good_freq = np.sum(data_full, axis=1) != 0.
    row_min, row_max = data_full.min(axis=1), data_full.max(axis=1)
    for idx_freq in np.flatnonzero(data.good_freq & (row_min == row_max)):
        say(f"ERROR: bad data value of {row_min[idx_freq]} in channel {idx_freq}!")
        data.good_freq[idx_freq] = False

    if opt["preprocess_data"]:
        data.preprocess_data(normalize_variance=True, variance_range=opt["variance_range"])
    say(f"INFO: there are {data.good_freq.sum()} good frequencies...")
    data.downsample(opt["factor_freq_downsample"], opt["factor_time_downsample"])

    # initial guesses, fitburst_pipeline.py:382-455
    initial_parameters = data.burst_parameters
    num_components = len(initial_parameters["dm"])
    basic_parameters = {
        "amplitude": [-2.0], "arrival_time": [np.mean(data.times)], "burst_width": [0.05], "dm": [0.0],
        "dm_index": [-2.0], "ref_freq": [np.min(data.freqs)], "scattering_index": [-4.0],
        "spectral_index": [0.0], "spectral_running": [0.0],
    }
    for current_parameter in initial_parameters.keys():
        if len(initial_parameters[current_parameter]) == 0:
            say(f"WARNING: parameter '{current_parameter}' has no value in data file, overloading a basic guess...")
            initial_parameters[current_parameter] = basic_parameters[current_parameter] * num_components
    for current_parameter in basic_parameters.keys():
        if current_parameter not in initial_parameters:
            initial_parameters[current_parameter] = basic_parameters[current_parameter] * num_components
    current_parameters = deepcopy(initial_parameters)

    dm_incoherent = initial_parameters["dm"][0]
    if data.is_dedispersed:
        say("INFO: input data cube is already dedispersed!")
        say("INFO: setting 'dm' entry to 0, now considered a dm-offset parameter...")
        current_parameters["dm"] = [0.0] * len(initial_parameters["dm"])
    if not opt["remove_dispersion_smearing"]:
        dm_incoherent = 0.

    if existing_results is not None:
        current_parameters = existing_results["model_parameters"]
        num_components = len(current_parameters["dm"])

    for name, key in (("amplitude", "amplitude"), ("arrival_time", "arrival_time"), ("dm", "dm"),
                      ("scattering_timescale", "scattering_timescale"), ("spectral_index", "spectral_index"),
                      ("spectral_running", "spectral_running"), ("width", "burst_width")):
        if opt[name] is not None:
            current_parameters[key] = opt[name]
    if opt["dm"] is not None:  # fitburst_pipeline.py:436-437 (keyed on --dm, like the reference)
        current_parameters["dm_index"] = opt["dm_index"]
    if opt["ref_freq"] is not None:
        current_parameters["ref_freq"] = [opt["ref_freq"]] * num_components

    if opt["verbose"]:
        say(f"INFO: initial guess for {len(current_parameters['dm'])}-component model:")
        for label, values in current_parameters.items():
            say(f"    * {label}: {values}")

    say("INFO: computing dedispersion-index matrix")
    data.dedisperse(initial_parameters["dm"][0], current_parameters["arrival_time"][0],
                    ref_freq=initial_parameters["ref_freq"][0], dm_offset=0.0)

    times_windowed = data.times
    if opt["window"] is not None:
        data_windowed, times_windowed = data.window_data(current_parameters["arrival_time"][0],
                                                         window=opt["window"], as_tensor=True)
    else:
        data_windowed = data._data_full.get_dev(data._device)  # the spectrum stays on the device

    if opt["peakfind_rms"] is not None:
        say("INFO: running FindPeak to isolate burst components...")
        peaks = FindPeak(data_windowed, times_windowed, data.freqs, rms=opt["peakfind_rms"])
        peaks.find_peak(distance=opt["peakfind_dist"])
        current_parameters = peaks.get_parameters_dict(current_parameters)
        num_components = len(current_parameters["dm"])

    return dict(input_file=input_file, options=opt, data=data, data_windowed=data_windowed,
                times_windowed=times_windowed, initial_parameters=initial_parameters,
                current_parameters=current_parameters, num_components=num_components,
                dm_incoherent=dm_incoherent, outfile_substring=outfile_substring)


def fit_prepared(state: dict, output_dir: str = ".", write_json: bool = True) -> dict:
    """Model construction, fit iterations and the results file (fitburst_pipeline.py:492-573)."""
    opt, data = state["options"], state["data"]
    num_components = state["num_components"]
    parameters_to_fix = _fixed_parameters(opt)
    say("INFO: initializing model")
    model = SpectrumModeler(
        data.freqs, state["times_windowed"], dm_incoherent=state["dm_incoherent"],
        factor_freq_upsample=opt["factor_freq_upsample"], factor_time_upsample=opt["factor_time_upsample"],
        num_components=num_components, is_dedispersed=data.is_dedispersed, is_folded=opt["is_folded"],
        scintillation=opt["scintillation"], verbose=opt["verbose"], device=data._device)
    model.update_parameters(state["current_parameters"])

    output = None
    for current_iteration in range(opt["num_iterations"]):
        fitter = LSFitter(state["data_windowed"], model, data.good_freq, weighted_fit=True,
                          weight_range=opt["weight_range"])
        fitter.fix_parameter(list(parameters_to_fix))
        fitter.fit(exact_jacobian=True)
        results = getattr(fitter, "results", None)
        if results is None or not results.success:
            continue
        bestfit_params = fitter.fit_statistics["bestfit_parameters"]
        model.update_parameters(bestfit_params)
        current_params = model.get_parameters_dict()
        if "dm" not in parameters_to_fix:
            current_params["dm"] = list(bestfit_params["dm"] * num_components)
        if "scattering_timescale" not in parameters_to_fix:
            current_params["scattering_timescale"] = list(bestfit_params["scattering_timescale"] * num_components)
        if current_iteration == opt["num_iterations"] - 1:
            if opt["verbose"]:
                uncertainties = fitter.fit_statistics["bestfit_uncertainties"]
                say(f"INFO: best-fit estimate for {num_components}-component model:")
                for label, values in bestfit_params.items():
                    say(f"    * {label}: {values} +/- {uncertainties[label]}")
            output = {
                "initial_dm": state["initial_parameters"]["dm"][0],
                "initial_time": data.times_bin0,
                "model_parameters": {key: (list(value) if value is not None else None)
                                     for key, value in current_params.items()},
                "fit_statistics": fitter.fit_statistics,
                "fit_logistics": {"weight_range": opt["weight_range"]},
            }
            if write_json:
                path = os.path.join(output_dir, f"results_fitburst{state['outfile_substring']}.json")
                with open(path, "w") as out:
                    json.dump(output, out, indent=4)
    return {"input_file": state["input_file"], "results": output}


def run(input_file: str, output_dir: str = ".", write_json: bool = True, device=None, **overrides) -> dict:
    """One file through the whole flow; the options are the reference's `dest` names (DEFAULTS)."""
    return fit_prepared(prepare(input_file, device=device, **overrides), output_dir, write_json)


def run_many(input_files, output_dir: str = ".", write_json: bool = True, device=None, quiet: bool = False,
             **overrides):
    """
    Generator over many files: file k+1 is read, parsed, uploaded and preprocessed on a worker
    thread (its kernels go to that thread's own CUDA stream) while file k is fitted. Yields the
    dictionaries of fit_prepared() in input order; a file that fails yields {"error": ...}.
    """
    import torch

    files = list(input_files)
    if write_json and len(files) > 1 and not overrides.get("use_outfile_substring", False):
        overrides = dict(overrides, use_outfile_substring=True)  # one results file per input

    def prepare_on_side_stream(path):
        dev = torch.device(device if device is not None else "cuda")
        stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream), _quiet(quiet):  # thread-local: sys.stdout is never swapped
            state = prepare(path, device=dev, **overrides)
        stream.synchronize()
        return state

    with concurrent.futures.ThreadPoolExecutor(max_workers=1) as pool:
        pending = pool.submit(prepare_on_side_stream, files[0]) if files else None
        for idx, path in enumerate(files):
            try:
                state = pending.result()
            except Exception as exc:  # noqa: BLE001 - keep going with the other files
                state = exc
            pending = pool.submit(prepare_on_side_stream, files[idx + 1]) if idx + 1 < len(files) else None
            if isinstance(state, Exception):
                yield {"input_file": path, "results": None, "error": repr(state)}
                continue
            with _quiet(quiet):
                result = fit_prepared(state, output_dir, write_json)
            yield result  # outside the quiet block: the consumer's own prints are not swallowed


def main(argv=None):
    args = vars(build_parser().parse_args(argv))
    files = args.pop("file")
    for item in run_many(files, **args):
        status = "no successful fit" if item.get("results") is None else "done"
        print(f"INFO: {item['input_file']}: {item.get('error', status)}")


if __name__ == "__main__":
    main()
