"""
Small public routines next to the hot path (SURVEY.md section 8b "effectively public" functions and
8f.3 / 8f.4): routines.spectrum / ism / manipulate, analysis.stats.compute_test_f (+ batched form),
utilities.plotting.compute_downsampled_data, the page-locked .npz loader.
CPU: oracle vs the unmodified reference, host logic. GPU: CUDA mirrors vs the oracle.
"""
import os

import numpy as np
import pytest

from oracle import ref_shim
from oracle import routines_oracle as ro

RNG = np.random.default_rng(11)
FREQS = 400.0 + 400.0 * np.sort(RNG.uniform(size=257))


# ------------------------------------------------------------------ CPU: oracle pinned on the reference
@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_oracle_routines_match_reference():
    ref_shim.import_reference()
    from fitburst.analysis import stats as rstats
    from fitburst.routines import ism as rism, manipulate as rman, spectrum as rspec
    from fitburst.utilities import plotting as rplot

    np.testing.assert_array_equal(ro.compute_spectrum_rpl(FREQS, 600.0, 1.5, -3.0),
                                  rspec.compute_spectrum_rpl(FREQS, 600.0, 1.5, -3.0))
    np.testing.assert_array_equal(ro.compute_time_dm_delay(12.3, 4149.377593360996, -2.0, FREQS, 600.0),
                                  rism.compute_time_dm_delay(12.3, 4149.377593360996, -2.0, FREQS, 600.0))
    np.testing.assert_array_equal(ro.compute_time_dm_smear(12.3, 4149.377593360996, -2.0, FREQS, 0.0244),
                                  rism.compute_time_dm_smear(12.3, 4149.377593360996, -2.0, FREQS, 0.0244))
    np.testing.assert_array_equal(ro.compute_time_scattering(FREQS, 600.0, 2e-3, -4.0),
                                  rism.compute_time_scattering(FREQS, 600.0, 2e-3, -4.0))
    model = np.abs(RNG.normal(size=(96, 3)))
    data = model @ np.array([1.0, 0.4, 2.5]) + RNG.normal(0, 0.1, size=96)
    np.testing.assert_allclose(ro.compute_amplitude_per_channel(data, model),
                               rism.compute_amplitude_per_channel(data, model), rtol=1e-12)
    spec = RNG.normal(size=(32, 48))
    np.testing.assert_array_equal(ro.downsample_2d(spec, 4, 6), rman.downsample_2d(spec, 4, 6))
    np.testing.assert_array_equal(ro.downsample_1d(FREQS[:256], 8), rman.downsample_1d(FREQS[:256], 8))
    np.testing.assert_array_equal(ro.upsample_1d(FREQS[:16], 0.3, 8), rman.upsample_1d(FREQS[:16], 0.3, 8))
    for args in ((1000.0, 1100.0, 17, 12, 5000, 5000), (1000.0, 990.0, 17, 12, 5000, 5000), (820.0, 1300.0, 12, 6, 900, 900)):
        assert ro.compute_test_f(*args) == rstats.compute_test_f(*args)
    good = RNG.uniform(size=32) > 0.3
    times, freqs = np.arange(48) * 1e-3, 400.0 + np.arange(32) * 0.5
    got = ro.compute_downsampled_data(times, freqs, spec, good, 0.5 * spec, 4, 6)
    want = rplot.compute_downsampled_data(times, freqs, spec, good, 0.5 * spec, 4, 6)
    assert sorted(got) == sorted(want)
    for key in want:
        np.testing.assert_array_equal(got[key], want[key])


def test_compute_test_f_and_batched_form():
    from fitburst_b200.analysis import stats

    cases = [(1000.0, 1100.0, 17, 12, 5000, 5000), (1000.0, 990.0, 17, 12, 5000, 5000), (820.0, 1300.0, 12, 6, 900, 900)]
    for args in cases:
        assert stats.compute_test_f(*args) == ro.compute_test_f(*args)
    first = [{"fit_statistics": {"chisq_final": c[0], "num_fit_parameters": c[2], "num_observations": c[4]}} for c in cases]
    second = [{"chisq_final": c[1], "num_fit_parameters": c[3], "num_observations": c[5]} for c in cases]
    first.append(None)
    second.append({"chisq_final": 1.0, "num_fit_parameters": 3, "num_observations": 10})
    got = stats.compute_test_f_batched(first, second)
    np.testing.assert_allclose(got[:3], [ro.compute_test_f(*c) for c in cases], rtol=1e-14)
    assert np.isnan(got[3])
    with pytest.raises(ValueError):
        stats.compute_test_f_batched(first, second[:2])


def test_upsample_labels_are_host_arithmetic():
    from fitburst_b200.routines import manipulate

    np.testing.assert_array_equal(manipulate.upsample_1d(FREQS[:16], 0.3, 8), ro.upsample_1d(FREQS[:16], 0.3, 8))
    out = manipulate.upsample_orig(np.arange(5.0) * 2.0, 4)
    assert out.shape == (5, 4) and np.allclose(out.mean(axis=1), np.arange(5.0) * 2.0)


def _write_npz(path, data, bad):
    metadata = {"bad_chans": list(bad), "freqs_bin0": 400.1953125, "is_dedispersed": True, "num_freq": data.shape[0],
                "num_time": data.shape[1], "times_bin0": 59000.5, "res_freq": 0.5, "res_time": 1e-3}
    burst = {"amplitude": [-0.3], "arrival_time": [0.02], "burst_width": 1e-3, "dm": [0.0], "dm_index": [-2.0],
             "ref_freq": [600.0], "scattering_index": [-4.0], "scattering_timescale": [0.0], "spectral_index": [0.0],
             "spectral_running": [0.0]}
    np.savez(path, data_full=data, metadata=metadata, burst_parameters=burst)
    return metadata, burst


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16])
def test_generic_reader_loads_like_reference(tmp_path, dtype):
    """Host side of the loader (no GPU needed): same attribute values as the reference reader; the
    spectrum keeps its stored float dtype; the weights are expanded from the channel mask."""
    from fitburst_b200.backend.generic import DataReader

    data = (RNG.normal(size=(24, 40)) * 50).astype(dtype)
    path = os.path.join(tmp_path, "burst.npz")
    metadata, burst = _write_npz(path, data, bad=[2, 7, 23])
    reader = DataReader(path, device="cpu")
    reader.load_data()
    assert reader.data_pinned.dtype == (__import__("torch").float32 if dtype == np.float32 else __import__("torch").float64)
    np.testing.assert_array_equal(reader.data_full, data.astype(np.float64 if dtype != np.float32 else np.float32))
    weights = reader.data_weights
    assert weights.shape == data.shape and weights.dtype == np.float64
    assert np.all(weights[[2, 7, 23]] == 0.0) and weights.sum() == (24 - 3) * 40
    assert reader.burst_parameters["burst_width"] == [1e-3] and reader.burst_parameters["dm"] == [0.0]
    np.testing.assert_array_equal(reader.freqs, np.arange(24, dtype=np.float64) * 0.5 + 400.1953125 + 0.25)
    np.testing.assert_array_equal(reader.times, np.arange(40, dtype=np.float64) * 1e-3)
    assert (reader.num_freq, reader.num_time, reader.res_freq, reader.res_time) == (24, 40, 0.5, 1e-3)
    assert reader.times_bin0 == 59000.5 and reader.is_dedispersed is True
    if ref_shim.reference_available():
        ref_shim.import_reference()
        from fitburst.backend.generic import DataReader as RefReader
        ref = RefReader(path)
        ref.load_data()
        for name in ("freqs", "times", "data_weights"):
            np.testing.assert_array_equal(getattr(reader, name), getattr(ref, name))
        np.testing.assert_array_equal(np.asarray(reader.data_full, dtype=np.float64), np.asarray(ref.data_full, dtype=np.float64))
        assert reader.burst_parameters == ref.burst_parameters


def test_generic_reader_rejects_bad_files(tmp_path):
    from fitburst_b200.backend.generic import DataReader

    with pytest.raises(IOError):
        DataReader(os.path.join(tmp_path, "missing.npz"))
    path = os.path.join(tmp_path, "incomplete.npz")
    np.savez(path, data_full=np.zeros((4, 4)), metadata={})
    with pytest.raises(AssertionError):
        DataReader(path, device="cpu").load_data()
    path = os.path.join(tmp_path, "mismatch.npz")
    metadata, _ = _write_npz(path, np.zeros((4, 6)), bad=[])
    metadata["num_freq"] = 5
    np.savez(path, data_full=np.zeros((4, 6)), metadata=metadata, burst_parameters={})
    with pytest.raises(AssertionError):
        DataReader(path, device="cpu").load_data()


def test_two_fitters_share_a_model_host_logic():
    """the weights live on the model's plan; a fitter must notice that another fitter
    loaded its own since. Host-side part: the ownership token (the GPU part is in test_gpu_parity)."""
    from fitburst_b200.analysis import fitter as fitter_module
    import inspect
    source = inspect.getsource(fitter_module.LSFitter._sync_weights)
    assert "_weights_owner" in source and "_weights_token" in source


# ------------------------------------------------------------------ GPU: CUDA mirrors vs oracle
@pytest.mark.gpu
def test_cuda_routines_match_oracle(native_lib):
    from fitburst_b200.routines import ism, manipulate, spectrum

    np.testing.assert_allclose(spectrum.compute_spectrum_rpl(FREQS, 600.0, 1.5, -3.0),
                               ro.compute_spectrum_rpl(FREQS, 600.0, 1.5, -3.0), rtol=1e-13)
    assert isinstance(spectrum.compute_spectrum_rpl(450.0, 600.0, 1.5, -3.0), float)
    np.testing.assert_allclose(ism.compute_time_dm_delay(12.3, 4149.377593360996, -2.0, FREQS, 600.0),
                               ro.compute_time_dm_delay(12.3, 4149.377593360996, -2.0, FREQS, 600.0), rtol=1e-12, atol=1e-16)
    np.testing.assert_allclose(ism.compute_time_dm_delay(12.3, 4149.377593360996, -2.0, FREQS),
                               ro.compute_time_dm_delay(12.3, 4149.377593360996, -2.0, FREQS), rtol=1e-13)
    np.testing.assert_allclose(ism.compute_time_dm_smear(12.3, 4149.377593360996, -2.0, FREQS, 0.0244),
                               ro.compute_time_dm_smear(12.3, 4149.377593360996, -2.0, FREQS, 0.0244), rtol=1e-13)
    np.testing.assert_allclose(ism.compute_time_scattering(FREQS, 600.0, 2e-3, -4.0),
                               ro.compute_time_scattering(FREQS, 600.0, 2e-3, -4.0), rtol=1e-13)
    for nt, nc in ((96, 3), (1000, 8), (17, 1)):
        model = np.abs(RNG.normal(size=(nt, nc)))
        data = model @ RNG.uniform(0.2, 3.0, size=nc) + RNG.normal(0, 0.1, size=nt)
        np.testing.assert_allclose(ism.compute_amplitude_per_channel(data, model),
                                   ro.compute_amplitude_per_channel(data, model), rtol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        ism.compute_amplitude_per_channel(np.ones(8), np.zeros((8, 2)))
    spec = RNG.normal(size=(32, 48))
    np.testing.assert_allclose(manipulate.downsample_2d(spec, 4, 6), ro.downsample_2d(spec, 4, 6), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(manipulate.downsample_1d(FREQS[:256], 8), ro.downsample_1d(FREQS[:256], 8), rtol=1e-14)
    mask = RNG.uniform(size=64) > 0.5
    np.testing.assert_array_equal(manipulate.downsample_1d(mask, 4, boolean=True), ro.downsample_1d(mask, 4, boolean=True))
    with pytest.raises(ValueError):
        manipulate.downsample_2d(spec, 5, 6)


@pytest.mark.gpu
def test_cuda_downsampled_data_matches_oracle(native_lib):
    import torch
    from fitburst_b200.utilities import plotting

    spec, model = RNG.normal(size=(64, 96)), RNG.normal(size=(64, 96))
    good = RNG.uniform(size=64) > 0.3
    times, freqs = np.arange(96) * 1e-3, 400.0 + np.arange(64) * 0.5
    want = ro.compute_downsampled_data(times, freqs, spec, good, model, 4, 6)
    for as_tensor in (False, True):
        data_in = torch.from_numpy(spec).cuda() if as_tensor else spec
        got = plotting.compute_downsampled_data(times, freqs, data_in, good, model, 4, 6)
        assert sorted(got) == sorted(want)
        for key in want:
            np.testing.assert_allclose(np.asarray(got[key], dtype=float), np.asarray(want[key], dtype=float),
                                       rtol=1e-13, atol=1e-15, err_msg=key)
    assert plotting.compute_downsampled_data(times, freqs, spec, good, None, 4, 6)["model"] is None


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cuda_widen_and_reader_upload(tmp_path, dtype, native_lib):
    """float32 spectra are uploaded as float32 and widened exactly on the device (fb_widen_f32)."""
    import torch
    from fitburst_b200.backend.generic import DataReader
    from fitburst_b200.utilities.bases import upload_spectrum

    data = (RNG.normal(size=(37, 101)) * 50).astype(dtype)  # odd sizes: the scalar tail of the kernel
    out = upload_spectrum(data, torch.device("cuda"))
    assert out.dtype == torch.float64
    np.testing.assert_array_equal(out.cpu().numpy(), data.astype(np.float64))
    path = os.path.join(tmp_path, "burst.npz")
    _write_npz(path, data, bad=[1, 5])
    reader = DataReader(path)
    reader.load_data()
    assert reader.data_pinned.is_pinned()
    dev = reader._data_full.get_dev(reader._device)
    np.testing.assert_array_equal(dev.cpu().numpy(), data.astype(np.float64))
    wdev = reader._data_weights.get_dev(reader._device)
    assert wdev.shape == (37, 101) and float(wdev.sum()) == 35 * 101


@pytest.mark.gpu
def test_two_live_fitters_on_one_model(native_lib):
    """f1 must be fitted with ITS weights and channel list although f2 was
    created on the same model afterwards (in the reference every fitter owns its weights)."""
    from fitburst_b200 import synthetic as syn
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler

    freqs, times = syn.chime_axes(96, 128)
    truth = syn.truncate_components(syn.config2_truth(128), 2)
    kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2, num_components=2)
    model = SpectrumModeler(freqs, times, **kwargs)
    model.update_parameters(truth)
    clean = model.compute_model()
    good1, good2 = syn.make_good_freq(96, 0.3, 1), syn.make_good_freq(96, 0.5, 2)
    d1, d2 = syn.add_noise(clean, good1, 0.1, 1), syn.add_noise(clean, good2, 0.3, 2)
    fixed = ["dm_index", "scattering_index"]

    def fit_alone(data, good):
        m = SpectrumModeler(freqs, times, **kwargs)
        m.update_parameters(syn.config2_start(truth))
        f = LSFitter(data, m, good)
        f.fix_parameter(fixed)
        f.fit()
        return f

    want1, want2 = fit_alone(d1, good1), fit_alone(d2, good2)
    model.update_parameters(syn.config2_start(truth))
    f1 = LSFitter(d1, model, good1)
    f2 = LSFitter(d2, model, good2)
    for f in (f1, f2):
        f.fix_parameter(fixed)
    f1.fit()
    np.testing.assert_array_equal(f1.results.x, want1.results.x)
    assert f1.fit_statistics["num_freq_good"] == want1.fit_statistics["num_freq_good"]
    np.testing.assert_allclose(f1.hessian, want1.hessian, rtol=1e-12)
    model.update_parameters(syn.config2_start(truth))
    f2.fit()
    np.testing.assert_array_equal(f2.results.x, want2.results.x)
    # user code replacing the weights: chisq_initial follows
    f2.weights = f2.weights * 2.0
    model.update_parameters(syn.config2_start(truth))
    f2.fit()
    assert abs(f2.fit_statistics["chisq_initial"] - 4.0 * want2.fit_statistics["chisq_initial"]) \
        <= 1e-12 * f2.fit_statistics["chisq_initial"]
