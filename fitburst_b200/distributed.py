"""
Multi-GPU drivers of the hot path: one process per GPU, torch.distributed for the plumbing.

The reference has no parallelism of any kind; the three partitionings below are the natural ones
of its arithmetic:

 * batched independent fits (BASELINE.json config 3): bursts are dealt round-robin to ranks,
   no data-path collective, one gather of the per-burst results at the end;
 * DM x scattering chi^2 grid (config 4): rows of the DM axis are dealt to ranks, the spectrum is
   replicated, one gather of (chi^2_min, flat index) pairs for the final argmin;
 * one spectrum too large / too slow for one GPU (configs 2 and 5): channels are sharded -- the
   model has no cross-channel coupling (analysis/model.py:147 loop body) -- and every trust-region
   evaluation all-reduces the (n+1)^2 matrix [J r]^T [J r] (<= 37 KB); each rank then advances an
   identical copy of the solver state, so the replicas stay in lock-step without a broadcast.

The collective is a plain all-reduce of a few KB: latency-bound, NCCL over NVLink on GPUs, gloo
in the CPU tests. Everything numerical happens in the C-ABI library.
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(count: int, rank: int, world: int) -> list:
    """Round-robin deal of `count` independent work items (bursts, grid rows)."""
    return list(range(rank, count, world))


def shard_channels(num_freq: int, rank: int, world: int):
    """Contiguous channel block [lo, hi) of rank `rank`; blocks differ by at most one channel."""
    base, extra = divmod(num_freq, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class TrustRegionState:
    """Thin wrapper over fb_trf_* (the scipy TRF restatement of fitburst_b200/csrc/fb_trf.h)."""

    def __init__(self, x0, num_residuals, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=0):
        self._lib = _lib.load_library()
        self.n = len(x0)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        options = _lib.FitOptions(ftol, xtol, gtol, int(max_nfev), 0)
        self._handle = _lib._P()
        _lib.check(self._lib.fb_trf_create(self.n, int(num_residuals), _lib.dptr(x0), ctypes.byref(options),
                                           ctypes.byref(self._handle)))

    def start(self, gram):
        gram = np.ascontiguousarray(gram, dtype=np.float64)
        _lib.check(self._lib.fb_trf_start(self._handle, _lib.dptr(gram)))

    def propose(self):
        x = np.zeros(self.n, dtype=np.float64)
        code = self._lib.fb_trf_propose(self._handle, _lib.dptr(x))
        if code < 0:
            _lib.check(code)
        return x if code == 1 else None

    def update(self, gram):
        gram = np.ascontiguousarray(gram, dtype=np.float64)
        _lib.check(self._lib.fb_trf_update(self._handle, _lib.dptr(gram)))

    def result(self):
        res = _lib.FitResult()
        gram = np.zeros((self.n + 1, self.n + 1), dtype=np.float64)
        _lib.check(self._lib.fb_trf_result(self._handle, ctypes.byref(res), _lib.dptr(gram)))
        return res, gram

    def __del__(self):
        if getattr(self, "_handle", None) is not None:
            try:
                self._lib.fb_trf_destroy(self._handle)
            except Exception:  # noqa: BLE001
                pass
            self._handle = None


def sharded_trust_region_fit(x0, num_residuals_total, local_gram, group=None, device=None, **tolerances):
    """
    Channel-sharded least-squares fit. `local_gram(x)` returns this rank's (n+1, n+1) contribution
    to [J r]^T [J r] at packed parameters x (a torch tensor on `device`, or a numpy array); the
    contributions are summed over ranks with one all-reduce per evaluation and every rank runs the
    same trust-region state machine on the sum. Returns (FitResult, accepted gram, evaluations).
    """
    rank, world = _world(group)

    def reduced(x):
        g = local_gram(x)
        if not isinstance(g, torch.Tensor):
            g = torch.from_numpy(np.ascontiguousarray(g, dtype=np.float64))
            if device is not None:
                g = g.to(device)
        if world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
        return g.detach().cpu().numpy()

    state = TrustRegionState(x0, num_residuals_total, **tolerances)
    state.start(reduced(np.asarray(x0, dtype=np.float64)))
    evaluations = 1
    while True:
        x_trial = state.propose()
        if x_trial is None:
            break
        state.update(reduced(x_trial))
        evaluations += 1
    res, gram = state.result()
    return res, gram, evaluations


class ChannelShardedFitter:
    """
    LSFitter for one spectrum whose channels are split over the ranks of a process group: rank r
    holds data[lo:hi], freqs[lo:hi], good_freq[lo:hi] (shard_channels) and a SpectrumModeler built
    on its own channel block. fit() produces the same results on every rank.
    """

    def __init__(self, local_fitter, num_freq_total, group=None):
        self.fitter = local_fitter  # fitburst_b200.analysis.fitter.LSFitter on the local block
        self.group = group
        self.num_freq_total = int(num_freq_total)
        self.results = None

    def _local_gram(self, x):
        f = self.fitter
        f.model.update_parameters(f.load_fit_parameters_list(list(x)))
        plan, n = f._prepare()
        f.model._push_parameters()
        gram = torch.empty((n + 1, n + 1), dtype=torch.float64, device=f.model._device)
        with torch.cuda.device(f.model._device):
            _lib.check(_lib.load_library().fb_normal_equations_dev(
                plan, f._data_on_device().data_ptr(), gram.data_ptr(), f._stream()))
        return gram

    def _fit_device_loop(self, x0, m_total):
        """
        The trust-region loop on the device (fb_fit_device_*): per evaluation the local kernels, ONE
        all-reduce of the (n+1)^2 matrix on the same stream, then the one-warp solver step -- all
        enqueued ahead in chunks, no host synchronisation per trial point. Every rank sees the same
        reduced matrix, takes the same decisions and therefore stops in the same chunk.
        Returns (FitResult, accepted matrix, evaluations) or None when a trial point left the
        domain of the fast kernels (the caller then runs the host-driven loop).
        """
        f = self.fitter
        lib = _lib.load_library()
        world = _world(self.group)[1]
        plan, n = f._prepare()
        f.model._push_parameters()
        device = f.model._device
        gram = torch.zeros((n + 1, n + 1), dtype=torch.float64, device=device)
        data_ptr = f._data_on_device().data_ptr()
        options = _lib.FitOptions(1e-8, 1e-8, 1e-8, 0, 0)
        stopped, evaluations = ctypes.c_int32(0), ctypes.c_int32(0)
        with torch.cuda.device(device):
            stream = f._stream()
            _lib.check(lib.fb_fit_device_begin(plan, ctypes.byref(options), int(m_total), stream))
            chunk, enqueued, limit = 6, 0, 100 * n + 8
            while True:
                for _ in range(chunk):
                    _lib.check(lib.fb_fit_device_evaluate(plan, data_ptr, gram.data_ptr(), stream))
                    if world > 1:
                        dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=self.group)
                    _lib.check(lib.fb_fit_device_step(plan, gram.data_ptr(), stream))
                enqueued += chunk
                _lib.check(lib.fb_fit_device_poll(plan, ctypes.byref(stopped), ctypes.byref(evaluations), stream))
                if stopped.value or enqueued > limit:
                    break
                chunk = 4
            res = _lib.FitResult()
            accepted = np.zeros((n + 1, n + 1), dtype=np.float64)
            bailed = ctypes.c_int32(0)
            _lib.check(lib.fb_fit_device_end(plan, ctypes.byref(res), _lib.dptr(accepted), ctypes.byref(bailed), stream))
        if bailed.value or not stopped.value:
            return None
        return res, accepted, int(res.nfev)

    def fit(self, device_loop=True):
        """device_loop=False keeps the trust-region state machine on the host (one all-reduce AND
        one device->host read of the matrix per trial point): the cross-check of the device loop."""
        f = self.fitter
        x0 = f.get_fit_parameters_list()
        m_total = self.num_freq_total * f.model.num_time
        out = None
        if device_loop and not f.model.scintillation and f._data_on_device().is_cuda:
            start = {key: list(value) for key, value in f.load_fit_parameters_list(list(x0)).items()}
            out = self._fit_device_loop(x0, m_total)
            if out is None:  # rare: restart from x0 with the host-driven loop on the general kernels
                f.model.update_parameters(start)
        if out is None:
            out = sharded_trust_region_fit(x0, m_total, self._local_gram, group=self.group)
        res, gram, _ = out
        n = len(x0)
        x = np.array(res.x[:n])
        f.model.update_parameters(f.load_fit_parameters_list(x.tolist()))
        # exact Hessian: additive over channels as well
        hessian = np.zeros((n, n), dtype=np.float64)
        plan, _ = f._prepare()
        f.model._push_parameters()
        with torch.cuda.device(f.model._device):
            _lib.check(_lib.load_library().fb_hessian(plan, f._data_on_device().data_ptr(), _lib.dptr(hessian), f._stream()))
        h = torch.from_numpy(hessian).to(f.model._device)
        if _world(self.group)[1] > 1:
            dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
        hessian = h.cpu().numpy()
        covariance = np.linalg.inv(0.5 * hessian)
        self.results = {
            "x": x, "status": int(res.status), "nfev": int(res.nfev), "njev": int(res.njev),
            "cost": float(res.cost), "hessian": hessian, "hessian_approx": gram[:n, :n].copy(),
            "uncertainties": np.sqrt(np.diag(covariance)),
        }
        return self.results


def gather_results(local_items: dict, count: int, group=None) -> list:
    """Gathers {global index: result} dictionaries from all ranks into one ordered list."""
    rank, world = _world(group)
    if world == 1:
        merged = dict(local_items)
    else:
        parts = [None] * world
        dist.all_gather_object(parts, local_items, group=group)
        merged = {}
        for part in parts:
            merged.update(part)
    return [merged.get(idx) for idx in range(count)]


def grid_argmin(local_chisq: np.ndarray, local_rows: list, num_cols: int, group=None):
    """
    Final argmin of a chi^2 grid whose DM rows were dealt to ranks: every rank reduces its rows to
    (chi^2_min, global flat index) and the pairs are gathered. Ties resolve to the smallest flat
    index, like np.argmin on the full surface.
    """
    rank, world = _world(group)
    best = (np.inf, -1)
    if len(local_rows):
        flat = int(np.argmin(local_chisq))
        r, c = divmod(flat, num_cols)
        best = (float(local_chisq.reshape(len(local_rows), num_cols)[r, c]), int(local_rows[r]) * num_cols + c)
    if world > 1:
        pairs = [None] * world
        dist.all_gather_object(pairs, best, group=group)
    else:
        pairs = [best]
    pairs = [p for p in pairs if p[1] >= 0]
    chisq, flat = min(pairs, key=lambda p: (p[0], p[1]))
    return chisq, flat


def bind_to_gpu_numa(device_index: int) -> bool:
    """
    Binds the calling process to the CPUs NVML reports as closest to GPU `device_index` (same
    NUMA node / PCIe root), so that page-locked staging buffers allocated afterwards are
    node-local and host->device copies of different ranks do not share an inter-socket link.
    One process per GPU launched by torchrun is not bound by default. Returns False when NVML is
    unavailable (nothing is changed then).
    """
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:  # noqa: BLE001 - best effort, never fatal
        return False
