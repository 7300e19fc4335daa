"""Timings of the other BASELINE.json configurations on one GPU (evidence for DESIGN.md; the
headline benchmark is bench.py): config 3 (batched full-size fits), config 4 (chi^2 grid points at
full size), config 5 (one rank's channel shard of the 4096 x 262144 scintillation fit)."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fitburst_b200 import _lib, synthetic as syn
from fitburst_b200.analysis.fitter import LSFitter
from fitburst_b200.analysis.model import SpectrumModeler
from fitburst_b200.batch import chisq_grid, fit_batched

quiet = lambda: contextlib.redirect_stdout(io.StringIO())
which = sys.argv[1:] or ["3", "4", "5"]

def timed(fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps, out

if "4" in which or "3" in which:
    freqs, times = syn.chime_axes(16384, 1024)
    kwargs = dict(factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
    model = SpectrumModeler(freqs, times, **kwargs)
    truth = syn.config2_truth(1024)
    model.update_parameters(truth)
    clean = model.compute_model()

if "4" in which:
    good = syn.make_good_freq(16384, 0.35, 2)
    data = syn.add_noise(clean, good, 0.1, 2)
    with quiet():
        fitter = LSFitter(data, model, good)
    dm = np.linspace(-0.5, 0.5, 8); sc = np.logspace(-4, -1.5, 8)
    chisq_grid(fitter, dm[:2], sc[:2])
    dt, surf = timed(lambda: chisq_grid(fitter, dm, sc))
    a, b = np.unravel_index(np.argmin(surf), surf.shape)
    print(json.dumps({"config": 4, "grid": "8x8 of the 1024x1024 sweep", "ms_per_grid_point": dt / 64 * 1e3,
                      "projected_full_grid_s_1gpu": dt / 64 * 1024 * 1024, "argmin": [float(dm[a]), float(sc[b])]}))

if "3" in which:
    # bursts with their own component count (1-3), grouped by count; host-resident spectra in
    # page-locked memory, uploads overlapped with fits (fit_bursts -> fit_stream)
    from fitburst_b200.batch import BurstFitter, pinned_like
    count = int(os.environ.get("CFG3_BURSTS", "24"))
    models = {k: SpectrumModeler(freqs, times, factor_freq_upsample=8, factor_time_upsample=4, num_components=k)
              for k in (1, 2, 3)}
    bursts = []
    for bidx in range(count):
        tr, st = syn.random_burst(bidx, 1024)
        m = models[len(tr["amplitude"])]
        m.update_parameters(tr)
        g = syn.make_good_freq(16384, 0.3, 100 + bidx)
        bursts.append((pinned_like(syn.add_noise(m.compute_model(), g, 0.1, 100 + bidx)), g, st))
    del models
    kw = dict(factor_freq_upsample=8, factor_time_upsample=4)
    workers = int(os.environ.get("CFG3_WORKERS", "4"))
    batch = BurstFitter(freqs, times, ["dm_index", "scattering_index"], kw, workers=workers)
    batch.fit(bursts)   # warm-up: plans, kernels, worker contexts
    dt, out = timed(lambda: batch.fit(bursts))
    print(json.dumps({"config": 3, "bursts": count, "shape": "16384x1024, 1-3 comps (own count per burst), host-resident spectra",
                      "fits_per_s_1gpu": count / dt, "fits_in_flight": workers,
                      "statuses": [int(o["results"].status) for o in out], "nfev": [int(o["results"].nfev) for o in out],
                      "chisq_reduced": [round(float(o["fit_statistics"]["chisq_final_reduced"]), 4) for o in out]}))

if "5" in which:
    nf, nt, nc = 512, 262144, 8          # one of 8 ranks' channel block of the 4096-channel spectrum
    freqs = 400.1953125 + (np.arange(4096)[:nf] + 0.5) * (400.0 / 4096)
    times = np.arange(nt) * 2.56e-6
    params = {"amplitude": [0.0] * nc, "arrival_time": [0.3 + 2e-3 * k for k in range(nc)],
              "burst_width": [(20 + 25 * k) * 1e-6 for k in range(nc)], "dm": [0.0] * nc, "dm_index": [-2.0] * nc,
              "ref_freq": [600.0] * nc, "scattering_timescale": [1e-4] * nc, "scattering_index": [-4.0] * nc,
              "spectral_index": [0.0] * nc, "spectral_running": [0.0] * nc}
    plain = SpectrumModeler(freqs, times, num_components=nc)
    plain.update_parameters(params)
    dev = plain.compute_model_device()
    rng = torch.Generator(device="cuda"); rng.manual_seed(5)
    data_dev = dev * torch.exp(0.5 * torch.randn((nf, 1), device="cuda", dtype=torch.float64, generator=rng)) + \
        torch.randn((nf, nt), device="cuda", dtype=torch.float64, generator=rng)
    model5 = SpectrumModeler(freqs, times, num_components=nc, scintillation=True)
    start = {k: list(v) for k, v in params.items()}
    start["arrival_time"] = [t + 4e-6 for t in params["arrival_time"]]
    start["scattering_timescale"] = [1.1e-4] * nc
    model5.update_parameters(start)
    with quiet():
        f5 = LSFitter(data_dev, model5, np.ones(nf, dtype=bool))
        f5.fix_parameter(["dm_index", "scattering_index"])
        plan, n = f5._prepare(); model5._push_parameters()
        gram = np.zeros((n + 1, n + 1))
        lib = _lib.load_library()
        dt_eval, _ = timed(lambda: _lib.check(lib.fb_normal_equations(plan, data_dev.data_ptr(), _lib.dptr(gram), f5._stream())), 2)
        dt_fit, _ = timed(lambda: f5.fit())
    print(json.dumps({"config": 5, "shard": f"{nf} of 4096 channels x {nt} samples, {nc} comps, scintillation, n={n}",
                      "ms_per_evaluation": dt_eval * 1e3, "fit_s": dt_fit, "status": int(f5.results.status),
                      "nfev": int(f5.results.nfev), "chisq_reduced": f5.fit_statistics.get("chisq_final_reduced")}))
