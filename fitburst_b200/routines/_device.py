"""Shared plumbing of the small routine mirrors: numpy in, one kernel through the C ABI, numpy out."""
import numpy as np
import torch

from .. import _lib


def require_gpu():
    if not torch.cuda.is_available():
        raise _lib.FitburstB200Error("no CUDA device visible: fitburst_b200 has no CPU fallback")
    return _lib.load_library(), torch.device("cuda")


def to_device(array, device):
    return torch.from_numpy(np.ascontiguousarray(array, dtype=np.float64)).to(device)


def stream_of(device):
    return torch.cuda.current_stream(device).cuda_stream


def elementwise(call, freqs):
    """Runs `call(lib, in_ptr, count, out_ptr, stream)` over the flattened float64 array `freqs`
    and returns a result of the same shape (a Python float for scalar input)."""
    lib, device = require_gpu()
    array = np.asarray(freqs, dtype=np.float64)
    flat = to_device(array.reshape(-1), device)
    out = torch.empty_like(flat)
    with torch.cuda.device(device):
        _lib.check(call(lib, flat.data_ptr(), flat.numel(), out.data_ptr(), stream_of(device)))
    result = out.cpu().numpy().reshape(array.shape)
    return float(result) if array.ndim == 0 else result
