"""GPU tests of the batched workloads: chi^2 grid (config 4), batched fits (config 3) and the
channel-sharded driver on one rank (config 5's code path; the all-reduce is a no-op at world 1)."""
import numpy as np
import pytest

from fitburst_b200 import synthetic as syn
from tests.helpers import make_case, model_pair

pytestmark = pytest.mark.gpu


def _setup(case, fixed=("dm_index", "scattering_index"), seed=2):
    from fitburst_b200.analysis.fitter import LSFitter
    from oracle.fitburst_oracle import OracleFitter
    freqs, times, truth, kwargs = make_case(**case)
    cuda, oracle = model_pair(freqs, times, kwargs, truth)
    clean = oracle.compute_model()
    good = syn.make_good_freq(len(freqs), 0.3, seed)
    data = syn.add_noise(clean, good, 0.1, seed)
    start = syn.config2_start(truth)
    cuda.update_parameters(start)
    oracle.update_parameters(start)
    fc = LSFitter(data, cuda, good)
    fo = OracleFitter(data, oracle, good)
    fc.fix_parameter(list(fixed))
    fo.fix_parameter(list(fixed))
    return fc, fo, truth, start, good, data


def test_chisq_grid_matches_reference_definition(native_lib):
    from fitburst_b200.batch import chisq_grid
    fc, fo, truth, start, good, data = _setup(dict(num_freq=32, num_time=64, kf=2, kt=2, num_comp=2))
    dm_values = np.linspace(-0.05, 0.05, 5)
    sc_values = np.logspace(-3.5, -2.0, 4)
    got = chisq_grid(fc, dm_values, sc_values)
    want = np.zeros_like(got)
    fo.fit_parameters = ["dm", "scattering_timescale"]
    for a, dm in enumerate(dm_values):
        for b, sc in enumerate(sc_values):
            want[a, b] = np.sum(fo.compute_residuals([dm, sc]) ** 2)  # the reference's definition
    np.testing.assert_allclose(got, want, rtol=1e-10)
    assert np.unravel_index(np.argmin(got), got.shape) == np.unravel_index(np.argmin(want), want.shape)


def test_fit_batched_matches_individual_fits(native_lib):
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.batch import fit_batched
    from oracle.fitburst_oracle import OracleFitter, OracleModel
    num_freq, num_time, count = 32, 64, 3
    freqs, times = syn.chime_axes(num_freq, num_time)
    kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2, num_components=2)
    datas, goods, starts, oracles = [], [], [], []
    truth = syn.truncate_components(syn.config2_truth(num_time), 2)
    for b in range(count):
        t = {k: list(v) for k, v in truth.items()}
        t["amplitude"] = [a + 0.05 * b for a in t["amplitude"]]
        om = OracleModel(freqs, times, **kwargs)
        om.update_parameters(t)
        good = syn.make_good_freq(num_freq, 0.2, 10 + b)
        data = syn.add_noise(om.compute_model(), good, 0.1, 10 + b)
        start = syn.config2_start(t)
        om.update_parameters(start)
        fo = OracleFitter(data, om, good)
        fo.fix_parameter(["dm_index", "scattering_index"])
        fo.fit(with_statistics=False)
        datas.append(data); goods.append(good); starts.append(start); oracles.append(fo)
    model = SpectrumModeler(freqs, times, **kwargs)
    out = fit_batched(model, np.stack(datas), np.stack(goods), starts, ["dm_index", "scattering_index"])
    for b in range(count):
        assert out[b]["status"] == oracles[b].results.status
        assert out[b]["nfev"] == oracles[b].results.nfev
        np.testing.assert_allclose(out[b]["x"], oracles[b].results.x, rtol=1e-6)
        assert np.all(np.isfinite(out[b]["uncertainties"])) and np.all(out[b]["uncertainties"] > 0)
        # uncertainties of the batched device Hessian (all bursts in one launch) = those of an
        # individual LSFitter fit of the same burst
        import contextlib, io
        from fitburst_b200.analysis.fitter import LSFitter
        single_model = SpectrumModeler(freqs, times, **kwargs)
        single_model.update_parameters(starts[b])
        with contextlib.redirect_stdout(io.StringIO()):
            single = LSFitter(datas[b], single_model, goods[b])
            single.fix_parameter(["dm_index", "scattering_index"])
            single.fit()
            # fb_fit_batched evaluates the Hessian at the accepted best fit (include/fitburst_b200.h),
            # LSFitter at the last evaluated point like the reference: move the model to the best fit
            single.model.update_parameters(single.load_fit_parameters_list(single.results.x.tolist()))
            hessian, _ = single.compute_hessian(single.data, single.fit_parameters)
        want = np.sqrt(np.diag(np.linalg.inv(0.5 * hessian)))
        np.testing.assert_allclose(out[b]["uncertainties"], want, rtol=1e-6)


def test_channel_sharded_driver_single_rank(native_lib):
    from fitburst_b200.distributed import ChannelShardedFitter
    fc, fo, truth, start, good, data = _setup(dict(num_freq=32, num_time=64, kf=2, kt=2, num_comp=2))
    sharded = ChannelShardedFitter(fc, num_freq_total=32)
    res = sharded.fit()
    fo.fit(with_statistics=False)
    assert res["status"] == fo.results.status and res["nfev"] == fo.results.nfev
    np.testing.assert_allclose(res["x"], fo.results.x, rtol=1e-6)
    assert np.all(res["uncertainties"] > 0)


def test_multi_gpu_nccl(native_lib):
    """Runs tests/multi_gpu_check.py over NCCL when the box has at least two GPUs."""
    import os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "multi_gpu_check.py")],
        capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    assert "MULTI_GPU_OK" in proc.stdout


def test_device_and_sequential_batched_paths_agree(native_lib, monkeypatch):
    """The on-device batched solver (one warp per burst) and the host-driven loop make the same
    decisions: same status / nfev and the same best fit."""
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.batch import fit_batched
    num_freq, num_time, count = 48, 96, 6
    freqs, times = syn.chime_axes(num_freq, num_time)
    kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2, num_components=2)
    model = SpectrumModeler(freqs, times, **kwargs)
    datas, goods, starts = [], [], []
    for b in range(count):
        truth, start = syn.random_burst(b, num_time)
        truth = syn.truncate_components({k: (v * 2)[:2] if len(v) < 2 else v for k, v in truth.items()}, 2)
        truth["arrival_time"] = [0.4 * times[-1], 0.6 * times[-1]]
        model.update_parameters(truth)
        good = syn.make_good_freq(num_freq, 0.2 if b else 0.0, 30 + b)
        datas.append(syn.add_noise(model.compute_model(), good, 0.05, 30 + b))
        goods.append(good)
        starts.append(syn.config2_start(truth))
    fixed = ["dm_index", "scattering_index"]
    out = {}
    for mode in ("device", "sequential"):
        monkeypatch.setenv("FITBURST_B200_BATCHED", mode)
        out[mode] = fit_batched(model, np.stack(datas), np.stack(goods), starts, fixed)
    for b in range(count):
        d, s = out["device"][b], out["sequential"][b]
        assert d["status"] == s["status"] and d["nfev"] == s["nfev"] and d["njev"] == s["njev"], (b, d["status"], s["status"])
        np.testing.assert_allclose(d["x"], s["x"], rtol=1e-7)
        np.testing.assert_allclose(d["uncertainties"], s["uncertainties"], rtol=1e-5)


def test_fit_bursts_groups_by_component_count(native_lib):
    """Config-3 driver: bursts with 1-3 components each, grouped by count; every burst ends where
    its own LSFitter fit ends."""
    import contextlib, io
    from fitburst_b200 import synthetic as syn
    from fitburst_b200.analysis.fitter import LSFitter
    from fitburst_b200.analysis.model import SpectrumModeler
    from fitburst_b200.batch import fit_bursts
    freqs, times = syn.chime_axes(24, 256)
    kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2)
    bursts = []
    for index in range(6):
        truth, start = syn.random_burst(index, 256)
        model = SpectrumModeler(freqs, times, num_components=len(truth["amplitude"]), **kwargs)
        model.update_parameters(truth)
        good = syn.make_good_freq(len(freqs), 0.2, index)
        bursts.append((syn.add_noise(model.compute_model(), good, 0.05, index), good, start))
    fixed = ["dm_index", "scattering_index"]
    out = fit_bursts(freqs, times, bursts, fixed, kwargs)
    assert [o["num_components"] for o in out] == [1, 2, 3, 1, 2, 3]
    for (data, good, start), got in zip(bursts, out):
        model = SpectrumModeler(freqs, times, num_components=len(start["amplitude"]), **kwargs)
        model.update_parameters(start)
        with contextlib.redirect_stdout(io.StringIO()):
            single = LSFitter(data, model, good)
            single.fix_parameter(fixed)
            single.fit()
        assert got["results"].status == single.results.status and got["results"].nfev == single.results.nfev
        np.testing.assert_array_equal(got["results"].x, single.results.x)
