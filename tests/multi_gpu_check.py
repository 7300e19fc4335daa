"""
Multi-GPU check, launched with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

1. Channel-sharded fit (BASELINE.json config 5 partitioning): every rank owns a contiguous
   channel block, the (n+1)^2 normal-equation matrix is all-reduced over NCCL at every
   trust-region evaluation; the result must equal the single-GPU fit of the whole spectrum.
2. chi^2 grid with DM rows dealt to ranks and the final argmin gather (config 4).
3. Batched fits dealt round-robin to ranks and gathered (config 3).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fitburst_b200 import distributed as fbd  # noqa: E402
from fitburst_b200 import synthetic as syn  # noqa: E402
from fitburst_b200.analysis.fitter import LSFitter  # noqa: E402
from fitburst_b200.analysis.model import SpectrumModeler  # noqa: E402
from fitburst_b200.batch import chisq_grid, fit_batched  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    num_freq, num_time = 2048, 256
    freqs, times = syn.chime_axes(num_freq, num_time)
    truth = syn.config2_truth(num_time)
    kwargs = dict(factor_freq_upsample=4, factor_time_upsample=2, num_components=3)
    fixed = ["dm_index", "scattering_index"]
    quiet = contextlib.redirect_stdout(io.StringIO())

    full_model = SpectrumModeler(freqs, times, device=device, **kwargs)
    full_model.update_parameters(truth)
    clean = full_model.compute_model()
    good = syn.make_good_freq(num_freq, 0.3, 4)
    data = syn.add_noise(clean, good, 0.1, 4)
    start = syn.config2_start(truth)

    # -- 1. channel-sharded fit vs single-GPU fit
    lo, hi = fbd.shard_channels(num_freq, rank, world)
    local_model = SpectrumModeler(freqs[lo:hi], times, device=device, **kwargs)
    local_model.update_parameters(start)
    with quiet:
        local_fitter = LSFitter(data[lo:hi], local_model, good[lo:hi])
        local_fitter.fix_parameter(fixed)
    sharded = fbd.ChannelShardedFitter(local_fitter, num_freq_total=num_freq)
    res = sharded.fit()  # trust-region loop on the device, all-reduce enqueued between evaluation and step
    local_model.update_parameters(start)
    res_host = sharded.fit(device_loop=False)  # host-driven loop: one matrix read-back per trial point
    assert (res["status"], res["nfev"], res["njev"]) == (res_host["status"], res_host["nfev"], res_host["njev"]), \
        (res["status"], res["nfev"], res_host["status"], res_host["nfev"])
    np.testing.assert_allclose(res["x"], res_host["x"], rtol=1e-10)
    full_model.update_parameters(start)
    log = io.StringIO()
    with contextlib.redirect_stdout(log):
        single = LSFitter(data, full_model, good)
        single.fix_parameter(fixed)
        single.fit()
    assert getattr(single, "results", None) is not None, "single-GPU fit failed:\n" + log.getvalue()
    assert res["status"] == single.results.status, (res["status"], single.results.status)
    assert res["nfev"] == single.results.nfev
    np.testing.assert_allclose(res["x"], single.results.x, rtol=1e-9)
    want_unc = np.concatenate([np.atleast_1d(v) for v in single.fit_statistics["bestfit_uncertainties"].values()])
    np.testing.assert_allclose(np.sort(res["uncertainties"]), np.sort(want_unc), rtol=1e-6)
    gathered = [None] * world
    dist.all_gather_object(gathered, res["x"].tolist())
    assert all(g == gathered[0] for g in gathered), "ranks diverged"

    # -- 2. chi^2 grid rows dealt to ranks + argmin gather
    dm_values = np.linspace(-0.05, 0.05, 6)
    sc_values = np.logspace(-3.3, -2.3, 5)
    full_model.update_parameters(truth)
    rows = fbd.shard_indices(len(dm_values), rank, world)
    local = chisq_grid(single, dm_values[rows], sc_values)
    chisq, flat = fbd.grid_argmin(local, rows, len(sc_values))
    if rank == 0:
        surface = chisq_grid(single, dm_values, sc_values)
        assert flat == int(np.argmin(surface)) and abs(chisq - surface.min()) <= 1e-12 * surface.min()

    # -- 3. batched fits dealt round-robin
    count = 4
    bursts = []
    small_f, small_t = syn.chime_axes(256, 128)
    small_kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2, num_components=2)
    small_model = SpectrumModeler(small_f, small_t, device=device, **small_kwargs)
    for b in range(count):
        tr = syn.truncate_components(syn.config2_truth(128), 2)
        tr["amplitude"] = [a + 0.03 * b for a in tr["amplitude"]]
        small_model.update_parameters(tr)
        g = syn.make_good_freq(256, 0.2, 50 + b)
        bursts.append((syn.add_noise(small_model.compute_model(), g, 0.1, 50 + b), g, syn.config2_start(tr)))
    mine = fbd.shard_indices(count, rank, world)
    out = fit_batched(small_model, np.stack([bursts[b][0] for b in mine]), np.stack([bursts[b][1] for b in mine]),
                      [bursts[b][2] for b in mine], fixed) if mine else []
    merged = fbd.gather_results({b: out[k]["x"].tolist() for k, b in enumerate(mine)}, count)
    assert all(m is not None for m in merged)
    if rank == 0:
        ref = fit_batched(small_model, np.stack([b[0] for b in bursts]), np.stack([b[1] for b in bursts]),
                          [b[2] for b in bursts], fixed)
        for b in range(count):
            np.testing.assert_allclose(merged[b], ref[b]["x"], rtol=1e-12)
        print(f"MULTI_GPU_OK world={world} sharded_fit nfev={res['nfev']} status={res['status']}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
