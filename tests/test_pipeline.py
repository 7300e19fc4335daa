"""
The callers of the hot path: generic-format reader, FindPeak and the
fitburst_pipeline flow. CPU tests pin the host-side mirrors to the UNMODIFIED reference (when
/root/reference is present); GPU tests run the whole flow on the golden inputs and compare with
the results the reference script itself produced (oracle/make_pipeline_golden.py).
"""
import contextlib
import glob
import io
import json
import os

import numpy as np
import pytest

from oracle import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "pipeline_*.npz")))
needs_reference = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


def test_pipeline_fixtures_present():
    assert len(CASES) >= 2
    for name in CASES:
        golden = json.load(open(os.path.join(GOLDEN_DIR, f"{name}.json")))
        assert golden["reference_results"]["fit_statistics"]["bestfit_uncertainties"] is not None


@needs_reference
@pytest.mark.parametrize("name", CASES)
def test_reader_matches_reference(name):
    ref_shim.import_reference()
    from fitburst.backend.generic import DataReader as ReferenceReader
    from fitburst_b200.backend.generic import DataReader
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    ours, theirs = DataReader(path), ReferenceReader(path)
    ours.load_data()
    theirs.load_data()
    for attribute in ("num_freq", "num_time", "res_time", "res_freq", "times_bin0", "is_dedispersed"):
        assert getattr(ours, attribute) == getattr(theirs, attribute)
    assert ours.burst_parameters == theirs.burst_parameters
    for attribute in ("data_full", "data_weights", "times", "freqs"):
        np.testing.assert_array_equal(getattr(ours, attribute), getattr(theirs, attribute))
    # dedispersion index bookkeeping (utilities/bases.py:47-127) on the host
    for reader in (ours, theirs):
        reader.dedisperse(reader.burst_parameters["dm"][0], reader.burst_parameters["arrival_time"][0],
                          ref_freq=reader.burst_parameters["ref_freq"][0], dm_offset=0.0)
    np.testing.assert_array_equal(ours.dedispersion_idx, theirs.dedispersion_idx)


def test_reader_errors(tmp_path):
    from fitburst_b200.backend.generic import DataReader
    with pytest.raises(IOError):
        DataReader(str(tmp_path / "missing.npz"))
    path = str(tmp_path / "incomplete.npz")
    np.savez(path, data_full=np.zeros((2, 2)))
    with pytest.raises(AssertionError):
        DataReader(path).load_data()
    path = str(tmp_path / "mismatch.npz")
    metadata = {"bad_chans": [], "freqs_bin0": 400.0, "is_dedispersed": True, "num_freq": 3, "num_time": 2,
                "times_bin0": 0.0, "res_freq": 1.0, "res_time": 1e-3}
    np.savez(path, data_full=np.zeros((2, 2)), metadata=metadata, burst_parameters={"dm": 1.0})
    with pytest.raises(AssertionError):
        DataReader(path).load_data()


@needs_reference
def test_find_peak_matches_reference():
    ref_shim.import_reference()
    import sys
    sys.modules["matplotlib.pyplot"].plot = lambda *args, **kwargs: None  # find_peak plots as it goes
    from fitburst.analysis.peak_finder import FindPeak as ReferenceFindPeak
    from fitburst_b200.analysis.peak_finder import FindPeak
    rng = np.random.default_rng(4)
    times = np.arange(400) * 1e-3
    series = 0.02 * rng.normal(size=400)
    for centre, height in ((0.1, 1.0), (0.18, 0.6), (0.3, 0.25)):
        series += height * np.exp(-0.5 * ((times - centre) / 4e-3) ** 2)
    data = np.outer(np.linspace(0.5, 1.5, 24), series) + 0.01 * rng.normal(size=(24, 400))
    original = {"arrival_time": [0.1], "burst_width": [1e-3], "dm": [5.0], "amplitude": [-1.0]}
    for rms in (None, 0.2):
        ours, theirs = FindPeak(data, times, None, rms=rms), ReferenceFindPeak(data, times, None, rms=rms)
        with contextlib.redirect_stdout(io.StringIO()):
            ours.find_peak(distance=7)
            theirs.find_peak(distance=7)
        np.testing.assert_array_equal(ours.time_of_arrivals, theirs.time_of_arrivals)
        assert ours.burst_widths == theirs.burst_widths
        assert ours.get_parameters_dict(original) == theirs.get_parameters_dict(original)
        assert ours.get_parameters_dict(original, update_width=False) == theirs.get_parameters_dict(original)


def test_command_line_matches_reference_names():
    from fitburst_b200.pipelines.fitburst_pipeline import DEFAULTS, build_parser
    args = vars(build_parser().parse_args(["one.npz", "two.npz"]))
    assert args.pop("file") == ["one.npz", "two.npz"]
    assert args == DEFAULTS  # the reference's `dest` names and defaults (fitburst_pipeline.py:21-296)


# ---------------------------------------------------------------------------------- GPU
def _compare_with_reference(results, reference):
    ours, theirs = results["fit_statistics"], reference["fit_statistics"]
    assert set(ours) == set(theirs)
    for key in ("num_freq", "num_freq_good", "num_fit_parameters", "num_observations", "num_time"):
        assert ours[key] == theirs[key], key
    for key in ("chisq_initial", "chisq_final", "chisq_final_reduced", "snr"):
        assert abs(ours[key] - theirs[key]) <= 1e-8 * abs(theirs[key]), key
    for name, values in theirs["bestfit_parameters"].items():
        sigma = np.asarray(theirs["bestfit_uncertainties"][name])
        np.testing.assert_allclose(ours["bestfit_uncertainties"][name], sigma, rtol=1e-5)
        diff = np.abs(np.asarray(ours["bestfit_parameters"][name]) - np.asarray(values))
        # BASELINE.json: 1e-6 relative and well inside one sigma (parameters near zero: sigma only)
        assert np.all(diff <= np.maximum(1e-6 * np.abs(values), 1e-4 * sigma)), name
    assert results["initial_dm"] == reference["initial_dm"]
    assert results["initial_time"] == reference["initial_time"]
    assert set(results["model_parameters"]) == set(reference["model_parameters"])
    for name, values in reference["model_parameters"].items():
        np.testing.assert_allclose(results["model_parameters"][name], values, rtol=1e-6, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_pipeline_matches_reference_script(name, native_lib, tmp_path):
    from fitburst_b200.pipelines.fitburst_pipeline import build_parser, run
    golden = json.load(open(os.path.join(GOLDEN_DIR, f"{name}.json")))
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    args = vars(build_parser().parse_args([path] + golden["cli_options"]))
    args.pop("file")
    with contextlib.redirect_stdout(io.StringIO()):
        out = run(path, output_dir=str(tmp_path), **args)
    written = json.load(open(tmp_path / "results_fitburst.json"))  # JSON-serialisable like the reference's
    _compare_with_reference(written, golden["reference_results"])
    assert out["results"]["fit_statistics"]["chisq_final"] == written["fit_statistics"]["chisq_final"]


@pytest.mark.gpu
def test_run_many_overlaps_and_matches_single_runs(native_lib, tmp_path):
    from fitburst_b200.pipelines.fitburst_pipeline import build_parser, run_many
    # the same option set for every file: the first golden case's
    golden = json.load(open(os.path.join(GOLDEN_DIR, f"{CASES[0]}.json")))
    args = vars(build_parser().parse_args(["x"] + golden["cli_options"]))
    args.pop("file")
    good = os.path.join(GOLDEN_DIR, f"{CASES[0]}.npz")
    files = [good, str(tmp_path / "missing.npz"), good]
    items = list(run_many(files, output_dir=str(tmp_path), quiet=True, **args))
    assert [item["input_file"] for item in items] == files
    assert "error" in items[1] and items[1]["results"] is None
    for item in (items[0], items[2]):
        _compare_with_reference(item["results"], golden["reference_results"])
    assert items[0]["results"]["fit_statistics"] == items[2]["results"]["fit_statistics"]


@pytest.mark.gpu
def test_find_peak_on_device_tensor(native_lib):
    import torch
    from fitburst_b200.analysis.peak_finder import FindPeak
    rng = np.random.default_rng(9)
    data = rng.normal(size=(150, 333))
    times = np.arange(333) * 1e-3
    with contextlib.redirect_stdout(io.StringIO()):
        host = FindPeak(data, times, None)
        host.find_peak()
        dev = FindPeak(torch.from_numpy(data).cuda(), times, None)
        dev.find_peak()
    np.testing.assert_allclose(dev.mean_intensity_freq, host.mean_intensity_freq, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(dev.time_of_arrivals, host.time_of_arrivals)


@pytest.mark.gpu
def test_float32_file_through_pinned_loader_and_quiet_run_many(native_lib, tmp_path, capsys):
    """A float32 .npz (what telescopes write) goes through the page-locked loader as float32, is
    widened on the device and gives the fit of the same values stored as float64; run_many(quiet=True)
    silences the workers without touching sys.stdout (the consumer's own prints survive)."""
    import sys
    from fitburst_b200.pipelines.fitburst_pipeline import build_parser, run, run_many
    name = CASES[0]
    golden = json.load(open(os.path.join(GOLDEN_DIR, f"{name}.json")))
    src = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=True)
    data32 = src["data_full"].astype(np.float32)
    paths = {}
    for label, data in (("f32", data32), ("f64", data32.astype(np.float64))):
        paths[label] = str(tmp_path / f"burst_{label}.npz")
        np.savez(paths[label], data_full=data, metadata=src["metadata"], burst_parameters=src["burst_parameters"])
    args = vars(build_parser().parse_args([paths["f32"]] + golden["cli_options"]))
    args.pop("file")
    results = {}
    for label in ("f32", "f64"):
        out_dir = tmp_path / label
        out_dir.mkdir()
        with contextlib.redirect_stdout(io.StringIO()):
            results[label] = run(paths[label], output_dir=str(out_dir), **args)["results"]
    assert results["f32"] is not None
    for key, values in results["f64"]["model_parameters"].items():
        np.testing.assert_array_equal(np.asarray(results["f32"]["model_parameters"][key], dtype=float),
                                      np.asarray(values, dtype=float))
    assert results["f32"]["fit_statistics"]["chisq_final"] == results["f64"]["fit_statistics"]["chisq_final"]
    stdout_before = sys.stdout
    seen = []
    for item in run_many([paths["f32"], paths["f64"]], output_dir=str(tmp_path), quiet=True, **args):
        print("consumer line")
        seen.append(item["results"]["fit_statistics"]["chisq_final"])
    assert sys.stdout is stdout_before
    captured = capsys.readouterr().out
    assert captured.count("consumer line") == 2 and "INFO" not in captured
    assert seen[0] == seen[1] == results["f64"]["fit_statistics"]["chisq_final"]
