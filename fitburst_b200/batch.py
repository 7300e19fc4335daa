"""
Batched workloads of the hot path on one GPU (through the C ABI):

 * chisq_grid(): chi^2 surface over a (dm, scattering_timescale) grid -- BASELINE.json config 4;
   the reference equivalent is np.sum(LSFitter.compute_residuals(x) ** 2) per grid point.
 * fit_batched(): many independent bursts of identical shape fitted back to back with device-
   resident spectra -- config 3.
Multi-GPU sharding of both lives in fitburst_b200.distributed.
"""
import ctypes
import os
import time

import numpy as np
import torch

from . import _lib


def chisq_grid(fitter, dm_values, sc_values) -> np.ndarray:
    """
    chi^2[a, b] = sum(((data - model(dm_values[a], sc_values[b])) * w) ** 2) with every other
    parameter held at the model's current value. `fitter` is a fitburst_b200 LSFitter.
    """
    dm_values = np.ascontiguousarray(dm_values, dtype=np.float64)
    sc_values = np.ascontiguousarray(sc_values, dtype=np.float64)
    if len(dm_values) == 0 or len(sc_values) == 0:  # a rank that was dealt no rows of a sharded sweep
        return np.zeros((len(dm_values), len(sc_values)), dtype=np.float64)
    plan, _ = fitter._prepare()
    fitter.model._push_parameters()
    device = fitter.model._device
    out = torch.empty((len(dm_values), len(sc_values)), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.load_library().fb_chisq_grid(
            plan, fitter._data_on_device().data_ptr(), _lib.dptr(dm_values), len(dm_values),
            _lib.dptr(sc_values), len(sc_values), out.data_ptr(), fitter._stream()))
    return out.cpu().numpy()


def fit_batched(model, data, good_freq, start_parameters, fixed_parameters, weighted_fit=True,
                compute_uncertainties=True):
    """
    Fits `count` bursts of identical shape. data: (count, Nf, Nt) array or device tensor;
    good_freq: (count, Nf) bool; start_parameters: list of parameter dicts (one per burst).
    Returns a list of dicts {x, status, nfev, njev, cost, uncertainties, parameters}.
    """
    from .analysis.fitter import LSFitter

    lib = _lib.load_library()
    device = model._device
    count = len(start_parameters)
    data_dev = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64))
    data_dev = data_dev.to(device=device, dtype=torch.float64).contiguous()
    good_freq = np.asarray(good_freq, dtype=bool)
    num_freq, num_comp = model.num_freq, int(model.num_components)
    weights = torch.empty((count, num_freq), dtype=torch.float64, device=device)
    params = np.zeros((count, _lib.FB_NUM_PARAMETERS, num_comp), dtype=np.float64)
    fitter = None
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        for b in range(count):
            model.update_parameters(start_parameters[b])
            params[b] = model._packed_parameters()
            # weights through the same kernel LSFitter uses (analysis/fitter.py:506-542)
            fitter = LSFitter(data_dev[b], model, good_freq[b], weighted_fit=weighted_fit)
            weights[b] = torch.from_numpy(fitter.weights).to(device)
        fitter.fix_parameter(list(fixed_parameters))
    plan, num_free = fitter._prepare()
    params_dev = torch.from_numpy(params).to(device)
    results = (_lib.FitResult * count)()
    options = _lib.FitOptions(1e-8, 1e-8, 1e-8, 0, int(bool(compute_uncertainties)))
    with torch.cuda.device(device):
        _lib.check(lib.fb_fit_batched(plan, count, data_dev.data_ptr(), weights.data_ptr(), params_dev.data_ptr(),
                                      ctypes.byref(options), results, model._stream()))
    best = params_dev.cpu().numpy()
    out = []
    for b in range(count):
        r = results[b]
        item = {
            "x": np.array(r.x[:num_free]), "status": int(r.status), "nfev": int(r.nfev), "njev": int(r.njev),
            "cost": float(r.cost), "uncertainties": np.array(r.uncertainty[:num_free]) if compute_uncertainties else None,
            "parameters": {name: best[b, row].tolist() for name, row in _lib.PARAMETER_INDEX.items()},
        }
        out.append(item)
    return out


def pinned_like(array) -> np.ndarray:
    """Copy of `array` in page-locked host memory (what fit_stream wants its spectra in: the
    host->device copy of a pageable array is staged by the driver and cannot overlap a fit)."""
    tensor = torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64)).pin_memory()
    return tensor.numpy()


_HEAP_FROZEN = False


def _freeze_heap_once():
    """
    A full pass of Python's cyclic garbage collector over the ~1e6 objects that torch / scipy /
    numpy leave on the heap after import holds the GIL for 0.1-0.8 s and stalls every fit in
    flight (measured: one 24-fit batch in four took 0.2-0.5 s longer). Those objects are module
    level and immortal in practice: the first concurrent fit_stream call of a process collects
    once and moves them to the permanent generation (gc.freeze), after which full collections only
    look at what was allocated since and take microseconds. The collector stays enabled.
    FITBURST_B200_GC_FREEZE=0 opts out.
    """
    global _HEAP_FROZEN
    if _HEAP_FROZEN or os.environ.get("FITBURST_B200_GC_FREEZE", "1") == "0":
        return
    import gc
    gc.collect()
    gc.freeze()
    _HEAP_FROZEN = True


TRACE = None  # set to a list to collect (burst index, worker, t_taken, t_fit_start, t_fit_end) per concurrent fit


def _strip_lazy(results):
    if results is not None and hasattr(results, "_lazy"):
        object.__setattr__(results, "_lazy", {})
    return results


def _host_tensor(data):
    """Host spectrum as a torch tensor WITHOUT a host-side dtype conversion for float32 input
    (telescope data usually are float32: half the upload, widened exactly on the device)."""
    if isinstance(data, torch.Tensor):
        data = data if data.dtype in (torch.float32, torch.float64) else data.to(torch.float64)
        return data.contiguous()
    array = np.asarray(data)
    if array.dtype not in (np.float32, np.float64):
        array = array.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(array))


def _upload(buffer, source, staging, good_freq=None):
    """buffer (float64, device) <- source (host, float32 or float64) on the current stream.
    Returns the float32 staging buffer (created on first use) so that the caller can keep it.

    FITBURST_B200_ROW_UPLOAD=1 (opt-in): page-locked sources go through fb_upload_rows -- only the
    rows of the usable channels cross the host link (flagged channels have weight 0 and are never
    read by the fit; their rows of `buffer` keep whatever an earlier burst left there, the buffers
    start zeroed), read by a kernel straight from the pinned array, float32 widened on the way.
    Meant for hosts whose link is the bound (eight GPUs uploading at once); on one GPU the copy
    engine is faster than the kernel (287 vs 251 fits/s end to end with the uploads on the workers'
    copy streams; 297 vs 220 in round 1), hence off by default."""
    if good_freq is not None and os.environ.get("FITBURST_B200_ROW_UPLOAD", "0") == "1" and source.is_pinned():
        rows = np.flatnonzero(np.asarray(good_freq, dtype=bool)).astype(np.int32)
        rows_dev = torch.from_numpy(rows).to(buffer.device, non_blocking=True)
        with torch.cuda.device(buffer.device):
            _lib.check(_lib.load_library().fb_upload_rows(
                source.data_ptr(), source.element_size(), rows_dev.data_ptr(), len(rows), buffer.shape[0],
                buffer.shape[1], buffer.data_ptr(), torch.cuda.current_stream(buffer.device).cuda_stream))
        return staging
    if source.dtype == torch.float64:
        buffer.copy_(source, non_blocking=True)
        return staging
    if staging is None:
        staging = torch.empty(buffer.shape, dtype=torch.float32, device=buffer.device)
    staging.copy_(source, non_blocking=True)
    with torch.cuda.device(buffer.device):  # exact widening by the library's own kernel
        _lib.check(_lib.load_library().fb_widen_f32(staging.data_ptr(), staging.numel(), buffer.data_ptr(),
                                                    torch.cuda.current_stream(buffer.device).cuda_stream))
    return staging


def _clone_model(model):
    """A second SpectrumModeler with the configuration of `model` (own plan, own device tables)."""
    from .analysis.model import SpectrumModeler
    return SpectrumModeler(model.freqs, model.times, dm_incoherent=model.dm_incoherent,
                           factor_freq_upsample=model.factor_freq_upsample,
                           factor_time_upsample=model.factor_time_upsample,
                           num_components=int(model.num_components), is_dedispersed=model.is_dedispersed,
                           is_folded=model.is_folded, scintillation=model.scintillation, verbose=False,
                           device=model._device)


def fit_stream(model, bursts, fixed_parameters, weighted_fit=True, weight_range=None, exact_jacobian=True,
               workers=1):
    """
    Fits a sequence of bursts of the model's shape, overlapping the host->device copy of the next
    burst(s) with the fit of the current one -- the multi-file caller of LSFitter (the loop of
    pipelines/fitburst_pipeline.py:494-532 over many inputs) without the per-burst upload stall.

    bursts: iterable of (data, good_freq, start_parameters); data is a (Nf, Nt) float64 host array
    (page-locked for the overlap to happen, see pinned_like) or a device tensor.
    Yields one dict per burst, in order: {"results": OptimizeResult without the lazy `fun` / `jac`
    members, "fit_statistics": ..., "success": ...}. The device copy of a spectrum is recycled, so
    nothing that needs it is handed out.

    workers = 1: one fit at a time on the current stream, uploads on a second stream (double-
    buffered). workers > 1: that many fits in flight, each on its own host thread, CUDA stream,
    device buffer and clone of `model` (the C calls release the GIL). A single fit leaves the GPU
    idle about a quarter of the time -- trust-region steps on the host between evaluations, the
    statistics, the read-backs -- and a second / third fit fills those gaps; the per-burst
    results are identical to workers = 1 (every fit is still one sequential, deterministic
    computation). `model` itself is left untouched in that mode.
    """
    if workers > 1:
        _freeze_heap_once()
        yield from _fit_stream_concurrent(model, bursts, fixed_parameters, weighted_fit, weight_range,
                                          exact_jacobian, int(workers))
        return
    from .analysis import fitter as fitter_module

    device = model._device
    shape = (model.num_freq, model.num_time)
    main = torch.cuda.current_stream(device)
    copier = torch.cuda.Stream(device=device)
    buffers = [torch.zeros(shape, dtype=torch.float64, device=device) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]
    staging = [None]  # float32 staging buffer, only for float32 input (uploads are serial on `copier`)
    iterator = iter(bursts)

    def upload(slot, item):
        data = item[0]
        if isinstance(data, torch.Tensor) and data.is_cuda:
            return data.to(dtype=torch.float64).contiguous(), None
        source = _host_tensor(data)
        with torch.cuda.stream(copier):
            copier.wait_event(free[slot])  # the fit that used this buffer two bursts ago is done
            staging[0] = _upload(buffers[slot], source, staging[0], item[1])
            ready[slot].record(copier)
        return buffers[slot], ready[slot]

    for event in free:
        event.record(main)
    slot = 0
    item = next(iterator, None)
    staged = upload(slot, item) if item is not None else None
    while item is not None:
        following = next(iterator, None)
        staged_next = upload(slot ^ 1, following) if following is not None else None
        data_dev, event = staged
        if event is not None:
            main.wait_event(event)
        model.update_parameters(item[2])
        with fitter_module.quiet():
            fitter = fitter_module.LSFitter(data_dev, model, item[1], weighted_fit=weighted_fit,
                                            weight_range=weight_range)
            fitter.fix_parameter(list(fixed_parameters))
            fitter.fit(exact_jacobian=exact_jacobian)
        free[slot].record(main)
        yield {"results": _strip_lazy(getattr(fitter, "results", None)), "fit_statistics": fitter.fit_statistics,
               "success": fitter.success}
        item, staged, slot = following, staged_next, slot ^ 1
    copier.synchronize()


def _fit_stream_concurrent(model, bursts, fixed_parameters, weighted_fit, weight_range, exact_jacobian, workers):
    """fit_stream with `workers` fits in flight (see there). Bursts are taken from the iterator
    in order under a lock, at most 3 * workers ahead of the consumer; results come out in order.
    Every worker uploads its NEXT burst on a copy stream of its own into the second of two device
    buffers while its current fit runs (a worker that uploaded and fitted on one stream spent a
    third of its cycle without a fit in flight: 67 MB at the speed of the host link per burst)."""
    import threading

    from .analysis import fitter as fitter_module

    device = model._device
    shape = (model.num_freq, model.num_time)
    caller_stream = torch.cuda.current_stream(device)
    device_index = device.index if device.index is not None else torch.cuda.current_device()
    iterator = iter(bursts)
    cond = threading.Condition()
    state = {"taken": 0, "yielded": 0, "exhausted": False, "stop": False, "error": None, "alive": workers}
    done = {}

    def take(block=True):
        # block=False: the look-ahead of a worker that still holds an unfitted burst -- it must not
        # wait for the consumer, who may be waiting for exactly that burst
        with cond:
            while not state["stop"] and not state["exhausted"] and state["taken"] - state["yielded"] >= 3 * workers:
                if not block:
                    return None
                cond.wait()
            if state["stop"] or state["exhausted"]:
                return None
            try:
                item = next(iterator)
            except StopIteration:
                state["exhausted"] = True
                cond.notify_all()
                return None
            index = state["taken"]
            state["taken"] += 1
            return index, item

    def work(worker_index):
        try:
            torch.cuda.set_device(device_index)
            slot = slots[worker_index]
            clone, stream, copier = slot["clone"], slot["stream"], slot["copier"]
            buffers, stagings, ready = slot["buffers"], slot["stagings"], slot["ready"]

            def stage(item, which):
                """Starts the upload of a host burst into buffer `which` on the copy stream. The
                fit that last read that buffer has returned (fits are synchronous for their
                worker), so only the consumer side needs an event."""
                data = item[1][0]
                if isinstance(data, torch.Tensor) and data.is_cuda:
                    return
                source = _host_tensor(data)
                if buffers[which] is None:
                    buffers[which] = torch.zeros(shape, dtype=torch.float64, device=device)
                copier.wait_stream(stream)  # the buffer's zero fill; nothing else is pending on `stream` here
                with torch.cuda.stream(copier):
                    stagings[which] = _upload(buffers[which], source, stagings[which], item[1][1])
                    ready[which].record(copier)

            with torch.cuda.stream(stream), fitter_module.quiet():
                which = 0
                taken = take()
                if taken is not None:
                    stage(taken, which)
                while taken is not None:
                    upcoming = take(block=False)
                    if upcoming is not None:
                        stage(upcoming, which ^ 1)
                    index, (data, good_freq, start) = taken
                    t_taken = time.perf_counter()
                    if isinstance(data, torch.Tensor) and data.is_cuda:
                        stream.wait_stream(caller_stream)  # whatever produced the tensor
                        data_dev = data.to(dtype=torch.float64).contiguous()
                    else:
                        stream.wait_event(ready[which])
                        data_dev = buffers[which]
                    clone.update_parameters(start)
                    t_start = time.perf_counter()
                    fitter = fitter_module.LSFitter(data_dev, clone, good_freq, weighted_fit=weighted_fit,
                                                    weight_range=weight_range)
                    fitter.fix_parameter(list(fixed_parameters))
                    fitter.fit(exact_jacobian=exact_jacobian)
                    out = {"results": _strip_lazy(getattr(fitter, "results", None)),
                           "fit_statistics": fitter.fit_statistics, "success": fitter.success}
                    del fitter
                    if TRACE is not None:
                        TRACE.append((index, worker_index, t_taken, t_start, time.perf_counter()))
                    with cond:
                        done[index] = out
                        cond.notify_all()
                    if upcoming is None:  # window was full (or the input is exhausted): wait this time
                        upcoming = take()
                        if upcoming is not None:
                            stage(upcoming, which ^ 1)
                    taken = upcoming
                    which ^= 1
                copier.synchronize()
                stream.synchronize()
        except BaseException as exc:  # noqa: BLE001 - handed to the consumer
            with cond:
                state["error"] = exc
                state["stop"] = True
                cond.notify_all()
        finally:
            with cond:
                state["alive"] -= 1
                cond.notify_all()

    # worker contexts (model clone with its plan / device tables / scratch, CUDA stream, upload
    # buffer) are kept on `model` between calls: re-creating ~0.5 GB of device memory per worker
    # per call costs tens of milliseconds and stalls every stream (cudaMalloc / cudaFree; torch's
    # caching allocator does not hand a block cached for one stream to another)
    key = (model.num_freq, model.num_time, int(model.num_components), int(model.factor_freq_upsample),
           int(model.factor_time_upsample), bool(model.is_folded), bool(model.scintillation),
           float(model.dm_incoherent), id(model.freqs), id(model.times))
    if model.__dict__.get("_stream_slots_key") != key:
        model.__dict__["_stream_slots"] = []
        model.__dict__["_stream_slots_key"] = key
    slots = model.__dict__["_stream_slots"]
    while len(slots) < workers:
        slots.append({"clone": _clone_model(model), "stream": torch.cuda.Stream(device=device),
                      "copier": torch.cuda.Stream(device=device), "buffers": [None, None], "stagings": [None, None],
                      "ready": [torch.cuda.Event(), torch.cuda.Event()]})
    threads = [threading.Thread(target=work, args=(k,), name=f"fit_stream-{k}", daemon=True) for k in range(workers)]
    for thread in threads:
        thread.start()
    try:
        while True:
            with cond:
                while state["yielded"] not in done and state["error"] is None and state["alive"] > 0:
                    cond.wait()
                if state["error"] is not None:
                    raise state["error"]
                if state["yielded"] not in done:
                    break  # all workers finished and nothing is pending
                out = done.pop(state["yielded"])
                state["yielded"] += 1
                cond.notify_all()
            yield out
    finally:
        with cond:
            state["stop"] = True
            cond.notify_all()
        for thread in threads:
            thread.join()


class BurstFitter:
    """
    Fits independent bursts that share the axes but not the number of components (BASELINE.json
    config 3: "1-3 components, random DM/scattering"). The reference would build one
    SpectrumModeler(num_components=k) per burst (pipelines/fitburst_pipeline.py:413-424); here the
    bursts are grouped by component count, every group shares one modeler (one plan, one set of
    device tables, its fit_stream worker contexts) that lives as long as this object, and runs
    through fit_stream, so uploads and fits of different bursts overlap inside a group.
    """

    def __init__(self, freqs, times, fixed_parameters, model_kwargs=None, weighted_fit=True, weight_range=None,
                 device=None, workers=1):
        self.freqs, self.times = freqs, times
        self.fixed_parameters = list(fixed_parameters)
        self.model_kwargs = dict(model_kwargs or {})
        self.model_kwargs.pop("num_components", None)
        if device is not None:
            self.model_kwargs["device"] = device
        self.weighted_fit, self.weight_range, self.workers = weighted_fit, weight_range, int(workers)
        self.models = {}

    def model_for(self, num_components: int):
        from .analysis.model import SpectrumModeler
        if num_components not in self.models:
            self.models[num_components] = SpectrumModeler(self.freqs, self.times, num_components=num_components,
                                                          **self.model_kwargs)
        return self.models[num_components]

    def fit(self, bursts) -> list:
        """
        bursts: sequence of (data, good_freq, start_parameters); the component count of a burst is
        len(start_parameters["amplitude"]). Returns one fit_stream() dict per burst, in input
        order, with the extra key "num_components".
        """
        groups = {}
        for index, item in enumerate(bursts):
            groups.setdefault(len(item[2]["amplitude"]), []).append(index)
        out = [None] * len(bursts)
        for num_components in sorted(groups):
            members = groups[num_components]
            stream = fit_stream(self.model_for(num_components), (bursts[i] for i in members), self.fixed_parameters,
                                weighted_fit=self.weighted_fit, weight_range=self.weight_range, workers=self.workers)
            for index, result in zip(members, stream):
                result["num_components"] = num_components
                out[index] = result
        return out


def fit_bursts(freqs, times, bursts, fixed_parameters, model_kwargs=None, weighted_fit=True, weight_range=None,
               device=None, workers=1):
    """One-shot form of BurstFitter(...).fit(bursts) (plans and device tables are built and
    dropped inside the call: keep a BurstFitter when there is more than one batch)."""
    return BurstFitter(freqs, times, fixed_parameters, model_kwargs, weighted_fit, weight_range, device,
                       workers).fit(bursts)
