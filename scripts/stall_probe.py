"""Looks for intermittent stalls of fit_stream(workers=4) on device-resident spectra: repeats a
40-fit batch and prints the per-fit trace of every batch that took more than 1.5x the median."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fitburst_b200 import batch as fb_batch, synthetic as syn
from fitburst_b200.analysis.model import SpectrumModeler

class A: nfreq = 16384; ntime = 1024; ncomp = 3; kf = 8; kt = 4
args = A()
freqs, times = syn.chime_axes(args.nfreq, args.ntime)
model = SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
def model_fn(p):
    model.update_parameters(p); return model.compute_model()
bursts = [bench.make_burst(args, k, model_fn) for k in range(2)]
dev = [(torch.from_numpy(b[3]).cuda(), b[2], b[1]) for b in bursts]
fixed = ["dm_index", "scattering_index"]
workers = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sampler = None
if os.environ.get("WITH_SAMPLER"):
    sampler = bench.ClockSampler(0); sampler.start()
walls, traces = [], []
for rep in range(reps):
    fb_batch.TRACE = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    arrivals = []
    for out in fb_batch.fit_stream(model, (dev[s % 2] for s in range(40)), fixed, workers=workers):
        arrivals.append(time.perf_counter() - t0)
    torch.cuda.synchronize(); walls.append(time.perf_counter() - t0)
    traces.append((t0, list(fb_batch.TRACE), arrivals))
med = float(np.median(walls))
print("batch wall ms: median %.1f min %.1f max %.1f" % (1e3 * med, 1e3 * min(walls), 1e3 * max(walls)))
for rep, w in enumerate(walls):
    if w > 1.5 * med and rep > 0:
        t0, trace, arrivals = traces[rep]
        print(f"--- slow batch {rep}: {1e3 * w:.1f} ms")
        for idx, wk, a, b, c in sorted(trace):
            flag = " <<<" if (c - b) > 0.03 or (b - a) > 0.03 else ""
            print(f"  burst {idx:2d} worker {wk} taken {1e3*(a-t0):7.1f} prep {1e3*(b-a):6.1f} fit {1e3*(c-b):7.1f} ms{flag}")
