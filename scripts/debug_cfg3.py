import contextlib, io, os, sys
os.environ.setdefault("CUDA_LAUNCH_BLOCKING", "1")
os.environ.setdefault("FITBURST_B200_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fitburst_b200 import synthetic as syn
from fitburst_b200.analysis.model import SpectrumModeler
from fitburst_b200.batch import fit_batched
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
freqs, times = syn.chime_axes(nf, 1024)
model = SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
count = 6
datas, goods, starts = [], [], []
for bidx in range(count):
    tr, st = syn.random_burst(bidx, 1024)
    nc = len(tr["amplitude"])
    tr3 = {k: (list(v) + [v[-1]] * 3)[:3] for k, v in tr.items()}
    tr3["amplitude"] = list(tr["amplitude"]) + [-3.0] * (3 - nc)
    tr3["arrival_time"] = (list(tr["arrival_time"]) + [0.8, 0.9])[:3]
    model.update_parameters(tr3)
    g = syn.make_good_freq(nf, 0.3, 100 + bidx)
    datas.append(syn.add_noise(model.compute_model(), g, 0.1, 100 + bidx)); goods.append(g)
    starts.append(syn.config2_start(tr3))
data_dev = torch.from_numpy(np.stack(datas)).cuda()
os.environ["FITBURST_B200_BATCHED"] = "sequential"
mode = sys.argv[2] if len(sys.argv) > 2 else "single"
if mode == "single":
    order = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else range(count)
    for b in order:
        try:
            out = fit_batched(model, data_dev[b:b + 1], np.stack(goods)[b:b + 1], starts[b:b + 1], ["dm_index", "scattering_index"])
            print("burst", b, "status", out[0]["status"], "nfev", out[0]["nfev"], flush=True)
        except Exception as exc:
            print("burst", b, "FAILED", str(exc)[:300], flush=True)
            break
else:
    out = fit_batched(model, data_dev, np.stack(goods), starts, ["dm_index", "scattering_index"])
    print([o["status"] for o in out], [o["nfev"] for o in out])
