"""
World-size-2 gloo tests (CPU) of the multi-GPU host logic: channel-sharded trust-region fit with
an all-reduce of the normal-equation matrix, grid argmin gather, burst sharding. The per-rank
arithmetic is supplied by the numpy oracle here (no GPU in this container); on the B200 box the
same drivers are fed by the CUDA kernels (tests/test_gpu_batch.py).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fitburst_b200 import synthetic as syn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_problem(num_freq=24, num_time=64):
    from oracle.fitburst_oracle import OracleFitter, OracleModel
    freqs, times = syn.chime_axes(num_freq, num_time)
    truth = syn.truncate_components(syn.config2_truth(num_time), 2)
    kwargs = dict(factor_freq_upsample=2, factor_time_upsample=2, num_components=2)
    model = OracleModel(freqs, times, **kwargs)
    model.update_parameters(truth)
    clean = model.compute_model()
    good = syn.make_good_freq(num_freq, 0.25, 3)
    data = syn.add_noise(clean, good, 0.1, 3)
    return freqs, times, kwargs, truth, good, data


def _oracle_fitter(freqs, times, kwargs, start, good, data, weights=None):
    from oracle.fitburst_oracle import OracleFitter, OracleModel
    model = OracleModel(freqs, times, **kwargs)
    model.update_parameters(start)
    fitter = OracleFitter(data, model, good)
    fitter.fix_parameter(["dm_index", "scattering_index"])
    return fitter


def _worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fitburst_b200 import distributed as fbd
        freqs, times, kwargs, truth, good, data = _make_problem()
        start = syn.config2_start(truth)
        lo, hi = fbd.shard_channels(len(freqs), rank, world)
        # channel block of this rank; res_freq comes from the first two local labels, same spacing
        local = _oracle_fitter(freqs[lo:hi], times, kwargs, start, good[lo:hi], data[lo:hi])
        x0 = local.get_fit_parameters_list()

        def local_gram(x):
            r = local.compute_residuals(list(x))
            jac = local.compute_jacobian(list(x))
            full = np.concatenate([jac, r[:, None]], axis=1)
            return full.T @ full

        res, gram, evals = fbd.sharded_trust_region_fit(x0, len(freqs) * len(times), local_gram)
        n = len(x0)
        # grid argmin over DM rows dealt to ranks
        rows = fbd.shard_indices(6, rank, world)
        surface = np.add.outer(np.arange(6.0), np.arange(5.0))
        surface[4, 3] = -1.0
        chisq, flat = fbd.grid_argmin(surface[rows], rows, 5)
        gathered = fbd.gather_results({i: i * i for i in fbd.shard_indices(7, rank, world)}, 7)
        queue.put((rank, np.array(res.x[:n]), int(res.status), int(res.nfev), evals, chisq, flat, gathered))
    finally:
        dist.destroy_process_group()


def test_channel_sharded_fit_matches_single_process():
    from oracle.fitburst_oracle import OracleFitter
    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [queue.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort(key=lambda o: o[0])
    # every rank ends on the same point, having made the same decisions
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    assert outs[0][2:5] == outs[1][2:5]
    # ... which is scipy's own trajectory on the unsharded problem
    freqs, times, kwargs, truth, good, data = _make_problem()
    single = _oracle_fitter(freqs, times, kwargs, syn.config2_start(truth), good, data)
    single.fit(with_statistics=False)
    assert outs[0][2] == single.results.status and outs[0][3] == single.results.nfev
    np.testing.assert_allclose(outs[0][1], single.results.x, rtol=1e-6)
    for out in outs:
        assert out[5] == -1.0 and out[6] == 4 * 5 + 3
        assert out[7] == [i * i for i in range(7)]


def test_shard_helpers():
    from fitburst_b200 import distributed as fbd
    for world in (1, 2, 3, 8):
        blocks = [fbd.shard_channels(16384 + 5, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 16384 + 5
        assert all(blocks[k][1] == blocks[k + 1][0] for k in range(world - 1))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
        dealt = sorted(sum((fbd.shard_indices(4096, r, world) for r in range(world)), []))
        assert dealt == list(range(4096))


def test_trust_region_state_machine_matches_scipy_on_cpu():
    """fb_trf_* (host code of the C-ABI library, no GPU needed) fed with oracle normal equations
    makes the same accept/reject/terminate decisions as scipy's least_squares."""
    from fitburst_b200 import distributed as fbd
    freqs, times, kwargs, truth, good, data = _make_problem(num_freq=16, num_time=48)
    start = syn.config2_start(truth)
    fitter = _oracle_fitter(freqs, times, kwargs, start, good, data)
    x0 = fitter.get_fit_parameters_list()

    def gram(x):
        r = fitter.compute_residuals(list(x))
        jac = fitter.compute_jacobian(list(x))
        full = np.concatenate([jac, r[:, None]], axis=1)
        return full.T @ full

    res, _, evals = fbd.sharded_trust_region_fit(x0, len(freqs) * len(times), gram)
    ref = _oracle_fitter(freqs, times, kwargs, start, good, data)
    ref.fit(with_statistics=False)
    n = len(x0)
    assert int(res.status) == ref.results.status
    assert int(res.nfev) == ref.results.nfev and int(res.njev) == ref.results.njev
    np.testing.assert_allclose(np.array(res.x[:n]), ref.results.x, rtol=1e-7)
    assert abs(res.cost - ref.results.cost) <= 1e-10 * ref.results.cost
