"""CPU tests of the host-side logic of fitburst_b200.batch (no kernels involved): grouping of
bursts by component count, input ordering, host-array handling of the upload path."""
import numpy as np
import torch

from fitburst_b200 import batch


def test_burst_fitter_groups_by_component_count_and_keeps_input_order(monkeypatch):
    calls = []

    def fake_fit_stream(model, bursts, fixed_parameters, weighted_fit=True, weight_range=None, workers=1):
        items = list(bursts)
        calls.append((model, len(items), list(fixed_parameters), workers))
        for data, good, start in items:
            yield {"results": ("fitted", data), "fit_statistics": {}, "success": True}

    monkeypatch.setattr(batch, "fit_stream", fake_fit_stream)
    fitter = batch.BurstFitter(np.arange(4.0), np.arange(8.0), ["dm_index"], {"factor_freq_upsample": 2}, workers=3)
    monkeypatch.setattr(fitter, "model_for", lambda count: f"model-{count}")
    counts = [2, 1, 3, 1, 2, 2, 3]
    bursts = [(f"burst-{k}", None, {"amplitude": [0.0] * n}) for k, n in enumerate(counts)]
    out = fitter.fit(bursts)
    assert [o["results"] for o in out] == [("fitted", f"burst-{k}") for k in range(len(counts))]
    assert [o["num_components"] for o in out] == counts
    # one stream per component count, in ascending order, with the members in input order
    assert calls == [("model-1", 2, ["dm_index"], 3), ("model-2", 3, ["dm_index"], 3), ("model-3", 2, ["dm_index"], 3)]
    assert "num_components" not in fitter.model_kwargs


def test_host_tensor_keeps_float32_and_widens_everything_else():
    f32 = np.arange(6, dtype=np.float32).reshape(2, 3)
    assert batch._host_tensor(f32).dtype == torch.float32
    assert batch._host_tensor(f32.astype(np.float64)).dtype == torch.float64
    ints = batch._host_tensor(np.arange(6, dtype=np.int16).reshape(2, 3))
    assert ints.dtype == torch.float64 and ints[1, 2].item() == 5.0
    strided = batch._host_tensor(f32.T)  # made contiguous, values unchanged
    assert strided.is_contiguous() and strided.shape == (3, 2) and strided[2, 1].item() == 5.0
    assert batch._host_tensor(torch.ones(2, dtype=torch.float16)).dtype == torch.float64
