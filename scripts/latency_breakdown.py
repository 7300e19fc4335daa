"""Where one resident config-2 fit spends its wall time: C call (FITBURST_B200_TIMING=1 prints its own
breakdown), Python mirror around it. Usage: python scripts/latency_breakdown.py"""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FITBURST_B200_TIMING"] = "1"
import numpy as np, torch
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis import fitter as fm
from fitburst_b200.analysis.model import SpectrumModeler

freqs, times = syn.chime_axes(16384, 1024)
model = SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
truth = syn.config2_truth(1024)
model.update_parameters(truth)
good = syn.make_good_freq(16384, 0.35, 0)
data = syn.add_noise(model.compute_model(), good, 0.1, 0)
start = syn.config2_start(truth)
model.update_parameters(start)
with fm.quiet():
    fitter = fm.LSFitter(data, model, good)
    fitter.fix_parameter(["dm_index", "scattering_index"])
marks = {}
orig_native, orig_stats = fitter._fit_native, fitter._compute_fit_statistics
def timed(name, fn):
    def wrapper(*a, **k):
        t0 = time.perf_counter(); out = fn(*a, **k); marks[name] = marks.get(name, 0.0) + time.perf_counter() - t0; return out
    return wrapper
fitter._fit_native = timed("_fit_native", orig_native)
fitter._compute_fit_statistics = timed("_compute_fit_statistics", orig_stats)
for rep in range(6):
    model.update_parameters(start)
    marks.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with fm.quiet():
        fitter.fit()
    torch.cuda.synchronize(); total = time.perf_counter() - t0
    print(f"python: fit() {total*1e3:.3f} ms  _fit_native {marks['_fit_native']*1e3:.3f}  statistics {marks['_compute_fit_statistics']*1e3:.3f}  nfev {fitter.results.nfev}", file=sys.stderr)
