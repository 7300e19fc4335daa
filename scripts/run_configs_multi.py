"""
BASELINE.json configs 3, 4 and 5 at their stated sizes, one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
      scripts/run_configs_multi.py {3|4|5} [options]          (also runs with plain `python` on one GPU)

 3  batched independent fits: `--bursts` random 1-3-component bursts (fitburst_b200.synthetic.random_burst,
    burst b seeded by b), dealt round-robin to the ranks, generated on the device in waves, fitted through
    BurstFitter / fit_stream(workers=4), results gathered; per-burst results are checked against
    individual LSFitter fits for a few bursts.
 4  chi^2 grid sweep: --grid x --grid points of (dm offset, scattering_timescale) on one config-2 spectrum
    replicated on every rank, DM rows dealt to the ranks, final argmin gather (distributed.grid_argmin).
 5  one baseband-like spectrum of 4096 channels x 262144 samples, 8 components with per-channel
    scintillation amplitudes, channel-sharded over the ranks with an all-reduce of the (n+1)^2
    normal-equation matrix per trust-region evaluation (distributed.ChannelShardedFitter).

Each prints one JSON line on rank 0 (and appends it to gpurun_out/r2_configs.jsonl when that directory exists).
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fitburst_b200 import distributed as fbd  # noqa: E402
from fitburst_b200 import synthetic as syn  # noqa: E402
from fitburst_b200.analysis.fitter import LSFitter  # noqa: E402
from fitburst_b200.analysis.model import SpectrumModeler  # noqa: E402
from fitburst_b200.batch import BurstFitter, chisq_grid  # noqa: E402

quiet = lambda: contextlib.redirect_stdout(io.StringIO())  # noqa: E731
FIXED = ["dm_index", "scattering_index"]


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    return rank, world, device


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def emit(rank, line):
    if rank != 0:
        return
    text = json.dumps(line)
    print(text, flush=True)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "r2_configs.jsonl"), "a") as fh:
            fh.write(text + "\n")


def max_over_ranks(value, world, device):
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


# ------------------------------------------------------------------------------------------ config 3
def config3(args, rank, world, device):
    freqs, times = syn.chime_axes(16384, 1024)
    kw = dict(factor_freq_upsample=8, factor_time_upsample=4)
    models = {k: SpectrumModeler(freqs, times, num_components=k, device=device, **kw) for k in (1, 2, 3)}
    mine = fbd.shard_indices(args.bursts, rank, world)
    batch = BurstFitter(freqs, times, FIXED, kw, device=device, workers=args.workers)
    generator = torch.Generator(device=device)

    def make(b):
        truth, start = syn.random_burst(b, 1024)
        model = models[len(truth["amplitude"])]
        model.update_parameters(truth)
        good = syn.make_good_freq(16384, 0.3, 100 + b)
        generator.manual_seed(1000 + b)
        data = model.compute_model_device().clone()
        data += 0.1 * torch.randn(data.shape, device=device, dtype=torch.float64, generator=generator)
        data[torch.from_numpy(~good).to(device)] = 0.0
        return data, good, start

    # warm-up: plans, kernels, worker contexts of the three component counts
    with quiet():
        batch.fit([make(b) for b in (0, 1, 2)])
    results, fit_seconds, gen_seconds = {}, 0.0, 0.0
    check = {}
    barrier(world)
    t_all = time.perf_counter()
    for lo in range(0, len(mine), args.wave):
        wave = mine[lo:lo + args.wave]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bursts = [make(b) for b in wave]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        with quiet():
            out = batch.fit(bursts)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        gen_seconds += t1 - t0
        fit_seconds += t2 - t1
        for b, o in zip(wave, out):
            r = o["results"]
            results[b] = None if r is None else {
                "status": int(r.status), "nfev": int(r.nfev), "x": [float(v) for v in r.x],
                "chisq_reduced": float(o["fit_statistics"].get("chisq_final_reduced", float("nan")))}
        if lo == 0 and rank == 0:  # a few bursts again, one at a time through LSFitter: identical results
            for b, (data, good, start) in list(zip(wave, bursts))[:3]:
                model = models[len(start["amplitude"])]
                model.update_parameters(start)
                with quiet():
                    single = LSFitter(data, model, good)
                    single.fix_parameter(FIXED)
                    single.fit()
                check[b] = bool(results[b] is not None and np.array_equal(single.results.x, np.array(results[b]["x"])))
        del bursts
    barrier(world)
    wall = time.perf_counter() - t_all
    fit_max = max_over_ranks(fit_seconds, world, device)
    merged = fbd.gather_results(results, args.bursts)
    if rank == 0:
        ok = [m for m in merged if m is not None]
        nfev = np.array([m["nfev"] for m in ok])
        hist = {str(int(k)): int(v) for k, v in zip(*np.unique(nfev, return_counts=True))}
        chi = np.array([m["chisq_reduced"] for m in ok])
        emit(rank, {
            "config": 3, "n_gpus": world, "bursts": args.bursts, "fitted": len(ok),
            "shape": "16384 x 1024, 1-3 components per burst (random_burst), 8x4 up-sampling, device-generated",
            "fits_per_s": args.bursts / fit_max, "fit_seconds_max_over_ranks": fit_max,
            "generation_seconds_this_rank": gen_seconds, "wall_seconds": wall, "fits_in_flight_per_gpu": args.workers,
            "status_counts": {str(int(k)): int(v) for k, v in zip(*np.unique([m["status"] for m in ok], return_counts=True))},
            "mean_nfev": float(nfev.mean()), "max_nfev": int(nfev.max()), "nfev_histogram": hist,
            "chisq_reduced_median": float(np.median(chi)), "chisq_reduced_max": float(np.max(chi)),
            "equal_to_individual_fits": check,
        })


# ------------------------------------------------------------------------------------------ config 4
def config4(args, rank, world, device):
    freqs, times = syn.chime_axes(16384, 1024)
    kwargs = dict(factor_freq_upsample=8, factor_time_upsample=4, num_components=3)
    model = SpectrumModeler(freqs, times, device=device, **kwargs)
    truth = syn.config2_truth(1024)
    model.update_parameters(truth)
    clean = model.compute_model()
    good = syn.make_good_freq(16384, 0.35, 2)
    data = syn.add_noise(clean, good, 0.1, 2)  # the same spectrum on every rank (replicated)
    with quiet():
        fitter = LSFitter(data, model, good)
    dm = np.linspace(-0.5, 0.5, args.grid)
    sc = np.logspace(-4, -1.5, args.grid)
    rows = fbd.shard_indices(len(dm), rank, world)
    chisq_grid(fitter, dm[:1], sc[:2])  # warm-up
    barrier(world)
    t0 = time.perf_counter()
    local = chisq_grid(fitter, dm[rows], sc)
    torch.cuda.synchronize()
    seconds = time.perf_counter() - t0
    chisq, flat = fbd.grid_argmin(local, rows, len(sc))
    barrier(world)
    wall = time.perf_counter() - t0
    seconds = max_over_ranks(seconds, world, device)
    # the surface on rank 0 (8 MB at 1024 x 1024) for the coarse record
    parts = [None] * world
    if world > 1:
        dist.all_gather_object(parts, (rows, local))
    else:
        parts = [(rows, local)]
    if rank == 0:
        surface = np.zeros((len(dm), len(sc)))
        for r, block in parts:
            surface[r] = block
        a, b = divmod(flat, len(sc))
        # consistency with the fit path: chi^2 at the argmin equals the corner of [J r]^T [J r] there
        point = {k: list(v) for k, v in truth.items()}
        point["dm"] = [float(dm[a])] * 3
        point["scattering_timescale"] = [float(sc[b])] * 3
        model.update_parameters(point)
        with quiet():
            fitter.fix_parameter(FIXED)
        from fitburst_b200 import _lib
        plan, n = fitter._prepare()
        model._push_parameters()
        gram = np.zeros((n + 1, n + 1))
        _lib.check(_lib.load_library().fb_normal_equations(plan, fitter._data_on_device().data_ptr(), _lib.dptr(gram),
                                                           fitter._stream()))
        step = max(1, len(dm) // 16)
        emit(rank, {
            "config": 4, "n_gpus": world, "grid": [len(dm), len(sc)], "points": len(dm) * len(sc),
            "seconds_max_over_ranks": seconds, "wall_seconds_with_argmin_gather": wall,
            "ms_per_grid_point_per_gpu": seconds / (len(rows) * len(sc)) * 1e3,
            "argmin": {"dm": float(dm[a]), "scattering_timescale": float(sc[b]), "chisq": chisq, "flat_index": flat},
            "argmin_equals_numpy_on_gathered_surface": bool(flat == int(np.argmin(surface))),
            "nearest_grid_point_to_truth": {"dm": float(dm[np.argmin(np.abs(dm - 0.0))]),
                                            "scattering_timescale": float(sc[np.argmin(np.abs(np.log(sc / 2e-3)))])},
            "chisq_at_argmin_vs_normal_equations_corner": [chisq, float(gram[n, n])],
            "coarse_surface_every": step, "coarse_surface": np.round(surface[::step, ::step], 3).tolist(),
        })


# ------------------------------------------------------------------------------------------ config 5
def config5(args, rank, world, device):
    nf_total, nt, nc = args.nfreq5, args.ntime5, 8
    freqs_all = 400.1953125 + (np.arange(nf_total) + 0.5) * (400.0 / nf_total)
    times = np.arange(nt) * 2.56e-6
    lo, hi = fbd.shard_channels(nf_total, rank, world)
    freqs = freqs_all[lo:hi]
    span = nt * 2.56e-6
    params = {"amplitude": [0.0] * nc, "arrival_time": [0.45 * span + 2e-3 * k for k in range(nc)],
              "burst_width": [(20 + 25 * k) * 1e-6 for k in range(nc)], "dm": [0.0] * nc, "dm_index": [-2.0] * nc,
              "ref_freq": [600.0] * nc, "scattering_timescale": [1e-4] * nc, "scattering_index": [-4.0] * nc,
              "spectral_index": [0.0] * nc, "spectral_running": [0.0] * nc}
    # data = sum_c a_ic * profile_c + N(0, 1), a_ic = exp(N(0, 0.5)) per (channel, component), seed 5
    rng = np.random.default_rng(5)
    amplitudes = np.exp(rng.normal(0.0, 0.5, size=(nf_total, nc)))[lo:hi]
    one = SpectrumModeler(freqs, times, num_components=1, device=device)
    data_dev = torch.zeros((hi - lo, nt), dtype=torch.float64, device=device)
    for c in range(nc):
        single = {k: [v[c]] for k, v in params.items()}
        one.update_parameters(single)
        data_dev += one.compute_model_device() * torch.from_numpy(amplitudes[:, c:c + 1].copy()).to(device)
    del one
    gen = torch.Generator(device=device)
    gen.manual_seed(5000 + rank)
    data_dev += torch.randn(data_dev.shape, device=device, dtype=torch.float64, generator=gen)
    model = SpectrumModeler(freqs, times, num_components=nc, scintillation=True, device=device)
    start = {k: list(v) for k, v in params.items()}
    start["arrival_time"] = [t + 4e-6 for t in params["arrival_time"]]
    start["burst_width"] = [w * 1.1 for w in params["burst_width"]]
    start["scattering_timescale"] = [1.1e-4] * nc
    start["dm"] = [2e-4] * nc
    model.update_parameters(start)
    with quiet():
        local = LSFitter(data_dev, model, np.ones(hi - lo, dtype=bool))
        local.fix_parameter(FIXED)
    n = len(local.get_fit_parameters_list())
    sharded = fbd.ChannelShardedFitter(local, num_freq_total=nf_total)
    barrier(world)
    t0 = time.perf_counter()
    res = sharded.fit()
    barrier(world)
    seconds = max_over_ranks(time.perf_counter() - t0, world, device)
    truth_x = np.array([params["arrival_time"][c] for c in range(nc)] + [params["burst_width"][c] for c in range(nc)]
                       + [params["dm"][0], params["scattering_timescale"][0]])
    xs = torch.from_numpy(res["x"]).to(device)
    x_min, x_max = xs.clone(), xs.clone()
    if world > 1:
        dist.all_reduce(x_min, op=dist.ReduceOp.MIN)
        dist.all_reduce(x_max, op=dist.ReduceOp.MAX)
    pull = (res["x"] - truth_x) / res["uncertainties"]
    emit(rank, {
        "config": 5, "n_gpus": world, "shape": [nf_total, nt], "components": nc, "scintillation": True,
        "free_parameters": n, "channels_per_rank": hi - lo, "fit_seconds": seconds, "status": res["status"],
        "nfev": res["nfev"], "njev": res["njev"], "seconds_per_evaluation": seconds / max(1, res["nfev"]),
        "chisq_reduced": 2.0 * res["cost"] / (nf_total * nt - n),
        "max_abs_pull_vs_truth": float(np.max(np.abs(pull))), "x_identical_on_all_ranks": bool(torch.equal(x_min, x_max)),
        "x": res["x"].tolist(), "uncertainties": res["uncertainties"].tolist(),
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["3", "4", "5"])
    ap.add_argument("--bursts", type=int, default=4096)
    ap.add_argument("--wave", type=int, default=48, help="config 3: bursts generated and fitted per wave on one GPU")
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--grid", type=int, default=1024)
    ap.add_argument("--nfreq5", type=int, default=4096)
    ap.add_argument("--ntime5", type=int, default=262144)
    args = ap.parse_args()
    rank, world, device = setup()
    {"3": config3, "4": config4, "5": config5}[args.config](args, rank, world, device)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
