"""General (block-granular) evaluation kernels at a config-5-like shape without scintillation:
split (k_vals + k_gram2_mma) against fused (k_eval<GRAM>, FITBURST_B200_SPLIT=0).
    python scripts/time_general_eval.py [--nfreq 128] [--ntime 262144]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fitburst_b200 import _lib  # noqa: E402
from fitburst_b200._say import quiet  # noqa: E402
from fitburst_b200.analysis.fitter import LSFitter  # noqa: E402
from fitburst_b200.analysis.model import SpectrumModeler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nfreq", type=int, default=128)
    ap.add_argument("--ntime", type=int, default=262144)
    ap.add_argument("--scintillation", type=int, default=0)
    args = ap.parse_args()
    nc, nt = 8, args.ntime
    freqs = 400.1953125 + (np.arange(args.nfreq) + 0.5) * (400.0 / 4096)
    times = np.arange(nt) * 2.56e-6
    span = nt * 2.56e-6
    params = {"amplitude": [0.0] * nc, "arrival_time": [0.45 * span + 2e-3 * k for k in range(nc)],
              "burst_width": [(20 + 25 * k) * 1e-6 for k in range(nc)], "dm": [0.0] * nc, "dm_index": [-2.0] * nc,
              "ref_freq": [600.0] * nc, "scattering_timescale": [1e-4] * nc, "scattering_index": [-4.0] * nc,
              "spectral_index": [0.0] * nc, "spectral_running": [0.0] * nc}
    model = SpectrumModeler(freqs, times, num_components=nc, scintillation=bool(args.scintillation), device="cuda:0")
    model.update_parameters(params)
    gen = torch.Generator(device="cuda:0")
    gen.manual_seed(1)
    data = torch.randn((args.nfreq, nt), device="cuda:0", dtype=torch.float64, generator=gen)
    with quiet():
        fitter = LSFitter(data, model, np.ones(args.nfreq, dtype=bool))
        fitter.fix_parameter(["dm_index", "scattering_index"] + (
            [] if args.scintillation else ["amplitude", "spectral_index", "spectral_running"]))
    plan, n = fitter._prepare()
    model._push_parameters()
    lib = _lib.load_library()
    out = {"shape": [args.nfreq, nt], "components": nc, "free_parameters": n, "scintillation": bool(args.scintillation)}
    grams = {}
    for name, env in (("split", None), ("fused", "0")):
        if env is not None:
            os.environ["FITBURST_B200_SPLIT"] = env
        gram = np.zeros((n + 1, n + 1))
        for _ in range(2):
            _lib.check(lib.fb_normal_equations(plan, fitter._data_on_device().data_ptr(), _lib.dptr(gram), fitter._stream()))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            _lib.check(lib.fb_normal_equations(plan, fitter._data_on_device().data_ptr(), _lib.dptr(gram), fitter._stream()))
        torch.cuda.synchronize()
        out[name + "_ms"] = round(1e3 * (time.perf_counter() - t0) / 5, 3)
        grams[name] = gram.copy()
        os.environ.pop("FITBURST_B200_SPLIT", None)
    scale = np.sqrt(np.outer(np.diag(grams["fused"]), np.diag(grams["fused"])))
    scale[scale == 0] = 1.0
    out["max_scaled_difference"] = float(np.max(np.abs(grams["split"] - grams["fused"]) / scale))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
