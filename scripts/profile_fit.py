"""Host-side profile of resident config-2 fits (cProfile): where the time outside the kernels goes."""
import contextlib, cProfile, io, pstats, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis.fitter import LSFitter
from fitburst_b200.analysis.model import SpectrumModeler

class A: nfreq = 16384; ntime = 1024; ncomp = 3; kf = 8; kt = 4
args = A()
freqs, times = syn.chime_axes(args.nfreq, args.ntime)
model = SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
def model_fn(p):
    model.update_parameters(p); return model.compute_model()
bursts = [bench.make_burst(args, k, model_fn) for k in range(2)]
fitters = []
with contextlib.redirect_stdout(io.StringIO()):
    for k in range(2):
        model.update_parameters(bursts[k][1])
        f = LSFitter(bursts[k][3], model, bursts[k][2]); f.fix_parameter(["dm_index", "scattering_index"]); fitters.append(f)
def one(k):
    model.update_parameters(bursts[k][1])
    with contextlib.redirect_stdout(io.StringIO()):
        fitters[k].fit()
for r in range(4): one(r % 2)
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 20
pr = cProfile.Profile(); pr.enable()
for r in range(N): one(r % 2)
torch.cuda.synchronize(); pr.disable()
print(f"{1e3 * (time.perf_counter() - t0) / N:.2f} ms per fit (under cProfile)")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(45); print(s.getvalue()[:9000])
