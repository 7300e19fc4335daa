"""Timing of the BASELINE.json configs[0] stand-in: 16384 x 162, one Gaussian
component, no up-sampling, pipeline-default fit (n = 6), full size, resident and from host memory."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fitburst_b200 import _lib, synthetic as syn
from fitburst_b200.analysis.fitter import LSFitter
from fitburst_b200.analysis.model import SpectrumModeler

quiet = lambda: contextlib.redirect_stdout(io.StringIO())
freqs, times = syn.chime_axes(syn.CONFIG1_NUM_FREQ, syn.CONFIG1_NUM_TIME)
model = SpectrumModeler(freqs, times, factor_freq_upsample=1, factor_time_upsample=1, num_components=1)
truth = syn.config1_truth()
model.update_parameters(truth)
good = syn.config1_good_freq()
data, sigma = syn.config1_data(model.compute_model(), good)
start = syn.config1_start(truth)

def one(host):
    model.update_parameters(start)
    with quiet():
        f = LSFitter(data if host else data_dev, model, good)
        f.fix_parameter(list(syn.CONFIG1_FIXED))
        f.fit()
    return f

data_dev = torch.from_numpy(data).cuda()
for _ in range(3):
    f = one(False)
out = {}
for label, host in (("resident", False), ("from_host_array", True)):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        f = one(host)
    torch.cuda.synchronize(); out[label + "_ms_per_fit"] = (time.perf_counter() - t0) / 10 * 1e3
plan, n = f._prepare(); model._push_parameters()
gram = np.zeros((n + 1, n + 1))
lib = _lib.load_library()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    _lib.check(lib.fb_normal_equations(plan, f._data_on_device().data_ptr(), _lib.dptr(gram), f._stream()))
torch.cuda.synchronize(); out["ms_per_evaluation"] = (time.perf_counter() - t0) / 20 * 1e3
out.update({"config": 1, "shape": [syn.CONFIG1_NUM_FREQ, syn.CONFIG1_NUM_TIME], "free_parameters": n, "status": int(f.results.status),
            "nfev": int(f.results.nfev), "snr": f.fit_statistics["snr"], "chisq_reduced": f.fit_statistics["chisq_final_reduced"],
            "kernels": "general (Gaussian branch: scattering_timescale = 0 is outside the scattered fast path)"})
print(json.dumps(out))
