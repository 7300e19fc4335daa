"""
Batched workloads of the hot path on one GPU (through the C ABI):

 * chisq_grid(): chi^2 surface over a (dm, scattering_timescale) grid -- BASELINE.json config 4;
   the reference equivalent is np.sum(LSFitter.compute_residuals(x) ** 2) per grid point.
 * fit_batched(): many independent bursts of identical shape fitted back to back with device-
   resident spectra -- config 3.
Multi-GPU sharding of both lives in fitburst_b200.distributed.
"""
import ctypes

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
