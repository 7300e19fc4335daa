"""Pre-fit data preparation: oracle vs the reference's ReaderBaseClass on
CPU, CUDA mirror (fitburst_b200.utilities.bases.ReaderBaseClass) vs oracle on the GPU."""
import contextlib
import io

import numpy as np
import pytest

from oracle import preprocess_oracle as po
from oracle import ref_shim

quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731


def make_raw(num_freq=64, num_time=512, seed=3, dispersed=True):
    rng = np.random.default_rng(seed)
    res_freq, res_time = 400.0 / num_freq, 0.00098304
    freqs = 400.1953125 + (np.arange(num_freq) + 0.5) * res_freq
    times = np.arange(num_time) * res_time
    data = 10.0 + rng.normal(0, 1.0, size=(num_freq, num_time)) * (1 + 0.3 * rng.uniform(size=(num_freq, 1)))
    arrival, dm = 0.2, 12.0
    delay = arrival + dm * po.DISPERSION_CONSTANT * (freqs ** -2.0 - 600.0 ** -2.0) if dispersed else np.full(num_freq, arrival)
    for i in range(num_freq):
        data[i] += 8.0 * np.exp(-0.5 * ((times - delay[i]) / 2e-3) ** 2)
    weights = np.ones((num_freq, num_time))
    bad = rng.choice(num_freq, size=num_freq // 8, replace=False)
    weights[bad] = 0.0
    data[bad[0]] = 5.0            # constant channel with zero weight
    data[(bad[1] + 1) % num_freq] = 7.0  # constant channel with non-zero weight -> "bad data value"
    data[3] *= 6.0                # high-variance channel
    weights[5, 100:140] = 0.0     # partially masked channel
    return freqs, times, res_freq, res_time, data, weights, arrival, dm


def fill(reader, raw):
    freqs, times, res_freq, res_time, data, weights, _, _ = raw
    reader.data_full = data.copy()
    reader.data_weights = weights.copy()
    reader.freqs, reader.times = freqs.copy(), times.copy()
    reader.num_freq, reader.num_time = data.shape
    reader.res_freq, reader.res_time = res_freq, res_time
    return reader


OPTIONS = [
    dict(),
    dict(remove_baseline=True),
    dict(remove_baseline=True, apply_cut_variance=True, variance_range=[0.05, 0.9]),
    dict(apply_cut_skewness=True, skewness_range=[-1.5, 1.5], normalize_variance=False),
]


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("options", OPTIONS)
def test_oracle_matches_reference(options):
    ref_shim.import_reference()
    from fitburst.utilities.bases import ReaderBaseClass as RefReader
    raw = make_raw()
    ref = fill(RefReader(), raw)
    with quiet():
        ref.preprocess_data(**options)
    data, good = po.preprocess(raw[4], raw[5], **{k: (tuple(v) if isinstance(v, list) else v) for k, v in options.items()})
    assert np.array_equal(good, ref.good_freq)
    np.testing.assert_allclose(data, ref.data_full, rtol=1e-15, atol=0)
    # downsample / dedisperse / window
    with quiet():
        ref.downsample(2, 4)
    np.testing.assert_allclose(po.block_mean_2d(data, 2, 4) * po.block_mean_1d(good, 2, True)[:, None], ref.data_full, rtol=1e-15)
    np.testing.assert_allclose(po.block_mean_1d(raw[0], 2), ref.freqs, rtol=0)
    ref.dedisperse(raw[7], raw[6], ref_freq=600.0)
    idx = po.dedispersion_index(ref.freqs, ref.times, ref.res_time, raw[7], raw[6], 600.0)
    assert np.array_equal(idx, ref.dedispersion_idx)
    win_ref, t_ref = ref.window_data(raw[6], window=0.04)
    win, t = po.window(ref.data_full, ref.times, ref.res_time, idx, raw[6], window=0.04)
    np.testing.assert_array_equal(win, win_ref)
    np.testing.assert_array_equal(t, t_ref)
    ref.is_dedispersed = True
    ref.dedisperse(raw[7], raw[6], ref_freq=600.0, dm_offset=0.5)
    assert np.array_equal(po.dedispersion_index(ref.freqs, ref.times, ref.res_time, raw[7], raw[6], 600.0,
                                                is_dedispersed=True, dm_offset=0.5), ref.dedispersion_idx)


@pytest.mark.gpu
@pytest.mark.parametrize("options", OPTIONS)
def test_cuda_reader_matches_oracle(options, native_lib):
    from fitburst_b200.utilities.bases import ReaderBaseClass
    raw = make_raw(num_freq=96, num_time=1000)
    reader = fill(ReaderBaseClass(), raw)
    with quiet():
        reader.preprocess_data(**options)
    data, good = po.preprocess(raw[4], raw[5], **{k: (tuple(v) if isinstance(v, list) else v) for k, v in options.items()})
    assert np.array_equal(reader.good_freq, good)
    np.testing.assert_allclose(reader.data_full, data, rtol=1e-13, atol=1e-13)
    with quiet():
        reader.downsample(2, 4)
    want = po.block_mean_2d(data, 2, 4) * po.block_mean_1d(good, 2, True)[:, None]
    np.testing.assert_allclose(reader.data_full, want, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(reader.data_weights, po.block_mean_2d(raw[5], 2, 4), rtol=1e-15)
    np.testing.assert_array_equal(reader.freqs, po.block_mean_1d(raw[0], 2))
    np.testing.assert_array_equal(reader.times, po.block_mean_1d(raw[1], 4))
    assert reader.num_freq == 48 and reader.num_time == 250 and reader.res_time == raw[3] * 4
    reader.dedisperse(raw[7], raw[6], ref_freq=600.0)
    idx = po.dedispersion_index(reader.freqs, reader.times, reader.res_time, raw[7], raw[6], 600.0)
    assert np.array_equal(reader.dedispersion_idx, idx)
    win, t = reader.window_data(raw[6], window=0.04)
    win_o, t_o = po.window(want, reader.times, reader.res_time, idx, raw[6], window=0.04)
    np.testing.assert_allclose(win, win_o, rtol=1e-13, atol=1e-13)
    np.testing.assert_array_equal(t, t_o)
    # device hand-off: the windowed tensor feeds LSFitter without another host round trip
    tensor, _ = reader.window_data(raw[6], window=0.04, as_tensor=True)
    assert tensor.is_cuda and np.array_equal(tensor.cpu().numpy(), win)
    with pytest.raises(ValueError):
        reader.window_data(raw[6], window=10.0)


@pytest.mark.gpu
def test_reader_to_fitter_handoff(native_lib):
    """A windowed spectrum produced on the device goes straight into LSFitter."""
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.utilities.bases import ReaderBaseClass
    raw = make_raw(num_freq=64, num_time=1024, dispersed=False)
    reader = fill(ReaderBaseClass(), raw)
    with quiet():
        reader.preprocess_data(remove_baseline=True)
    reader.is_dedispersed = True
    reader.dedisperse(0.0, raw[6], ref_freq=600.0, dm_offset=0.0)
    tensor, t_win = reader.window_data(raw[6], window=0.06, as_tensor=True)
    model = SpectrumModeler(reader.freqs, t_win, num_components=1, is_dedispersed=True)
    model.update_parameters({"amplitude": [-0.2], "arrival_time": [raw[6] + 3e-4], "burst_width": [2.5e-3], "dm": [0.0],
                             "dm_index": [-2.0], "ref_freq": [600.0], "scattering_timescale": [0.0],
                             "scattering_index": [-4.0], "spectral_index": [0.0], "spectral_running": [0.0]})
    with quiet():
        fitter = LSFitter(tensor, model, reader.good_freq)
        fitter.fix_parameter(["dm_index", "scattering_index", "scattering_timescale"])
        fitter.fit()
    assert fitter.results.success
    best = fitter.fit_statistics["bestfit_parameters"]
    assert abs(best["arrival_time"][0] - raw[6]) < 2e-4 and abs(best["burst_width"][0] - 2e-3) < 2e-4
