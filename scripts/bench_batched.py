"""Throughput of batched SMALL-burst fits: on-device solver vs host-driven loop (not the headline
benchmark; evidence for DESIGN.md section 4, 'Solver')."""
import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis.model import SpectrumModeler
from fitburst_b200.batch import fit_batched

num_freq, num_time, count = 256, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 512
freqs, times = syn.chime_axes(num_freq, num_time)
kwargs = dict(factor_freq_upsample=4, factor_time_upsample=2, num_components=2)
model = SpectrumModeler(freqs, times, **kwargs)
truth = syn.truncate_components(syn.config2_truth(num_time), 2)
model.update_parameters(truth)
clean = model.compute_model()
rng = np.random.default_rng(0)
data = clean[None] + rng.normal(0, 0.1, size=(count, num_freq, num_time))
good = np.ones((count, num_freq), dtype=bool)
starts = [syn.config2_start(truth) for _ in range(count)]
data_dev = torch.from_numpy(data).cuda()
for mode in ("device", "sequential"):
    os.environ["FITBURST_B200_BATCHED"] = mode
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fit_batched(model, data_dev, good, starts, ["dm_index", "scattering_index"], compute_uncertainties=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{mode:10s}: {count} fits of {num_freq}x{num_time} in {dt*1e3:.1f} ms -> {count/dt:.0f} fits/s "
          f"(statuses {sorted(set(o['status'] for o in out))}, mean nfev {np.mean([o['nfev'] for o in out]):.1f})")
