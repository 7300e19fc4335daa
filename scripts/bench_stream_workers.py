"""Throughput of fit_stream() at config-2 size against the number of fits in flight (workers),
with host-resident (page-locked) and device-resident spectra."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis.model import SpectrumModeler
from fitburst_b200.batch import fit_stream, pinned_like

class A: nfreq = 16384; ntime = 1024; ncomp = 3; kf = 8; kt = 4
args = A()
freqs, times = syn.chime_axes(args.nfreq, args.ntime)
model = SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
def model_fn(p):
    model.update_parameters(p); return model.compute_model()
bursts = [bench.make_burst(args, k % 2, model_fn) for k in range(4)]  # the two bursts of bench.py rank 0 (7 evaluations each)
host = [(pinned_like(b[3]), b[2], b[1]) for b in bursts]
dev = [(torch.from_numpy(b[3]).cuda(), b[2], b[1]) for b in bursts]
fixed = ["dm_index", "scattering_index"]
steps = int(os.environ.get("STEPS", "60"))
warm = int(os.environ.get("WARM", "20"))
for label, src in (("host", host), ("device", dev), ("host", host), ("device", dev)):
    for workers in [int(w) for w in (sys.argv[1:] or ["1", "2", "3", "4"])]:
        gen = fit_stream(model, (src[s % 4] for s in range(warm + steps)), fixed, workers=workers)
        for _ in range(warm):
            next(gen)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(steps):
            out = next(gen)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        for _ in gen: pass
        print(json.dumps({"spectra": label, "workers": workers, "fits_per_s": steps / dt, "ms_per_fit": 1e3 * dt / steps,
                          "status": int(out["results"].status), "nfev": int(out["results"].nfev)}), flush=True)
