"""chi^2 grid sweep (BASELINE.json config 4) on one GPU at config-2 size: the fused kernel
(k_vals3<.., CHI2>) against the stored-values sweep (FITBURST_B200_GRID_FUSED=0).
    python scripts/time_grid.py [--dm 4] [--sc 64] > gpurun_out/grid_timing.json"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitburst_b200 import synthetic as syn  # noqa: E402
from fitburst_b200._say import quiet  # noqa: E402
from fitburst_b200.analysis.fitter import LSFitter  # noqa: E402
from fitburst_b200.analysis.model import SpectrumModeler  # noqa: E402
from fitburst_b200.batch import chisq_grid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dm", type=int, default=4)
    ap.add_argument("--sc", type=int, default=64)
    args = ap.parse_args()
    freqs, times = syn.chime_axes(16384, 1024)
    model = SpectrumModeler(freqs, times, device="cuda:0", factor_freq_upsample=8, factor_time_upsample=4,
                            num_components=3)
    model.update_parameters(syn.config2_truth(1024))
    good = syn.make_good_freq(16384, 0.35, 2)
    data = syn.add_noise(model.compute_model(), good, 0.1, 2)
    with quiet():
        fitter = LSFitter(data, model, good)
    # the same slice of the 1024 x 1024 grid of config 4 (rows around the injected dm)
    dm = np.linspace(-0.5, 0.5, 1024)[512 - args.dm // 2:512 - args.dm // 2 + args.dm]
    sc = np.logspace(-4, -1.5, args.sc)
    out = {"shape": [16384, 1024], "good_channels": int(good.sum()), "grid": [len(dm), len(sc)]}
    surfaces = {}
    for name, env in (("fused", None), ("stored_values", "0")):
        if env is not None:
            os.environ["FITBURST_B200_GRID_FUSED"] = env
        chisq_grid(fitter, dm[:1], sc[:4])
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            surfaces[name] = chisq_grid(fitter, dm, sc)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out[name + "_ms_per_point"] = 1e3 * best / (len(dm) * len(sc))
        os.environ.pop("FITBURST_B200_GRID_FUSED", None)
    # scattering-time dependence (narrow scattering: long rise region, series instead of moments)
    per_tau = {}
    for tau in (1e-4, 3e-4, 1e-3, 3e-3, 1e-2):
        row = np.full(8, tau)
        chisq_grid(fitter, dm[:1], row)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        chisq_grid(fitter, dm[:2], row)
        torch.cuda.synchronize()
        per_tau["%g" % tau] = round(1e3 * (time.perf_counter() - t0) / 16, 4)
    out["fused_ms_per_point_by_scattering_time"] = per_tau
    out["max_rel_difference"] = float(np.max(np.abs(surfaces["fused"] / surfaces["stored_values"] - 1.0)))
    out["projected_seconds_1024x1024_on_8_gpus"] = out["fused_ms_per_point"] * 1024 * 1024 / 8 / 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
