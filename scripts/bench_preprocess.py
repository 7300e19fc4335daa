"""HBM roofline of the pre-fit data-preparation kernels on a raw
CHIME-sized spectrum (16384 channels x 4096 samples, 537 MB), with the numpy oracle timed beside
them on the host. Prints one JSON line per kernel."""
import ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fitburst_b200 import _lib
from oracle import preprocess_oracle as po

nf, nt = 16384, 4096
peak = 6557.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "MEASURED_PEAKS.json"
except Exception:
    peak, src = 6650.0, "fallback"
lib = _lib.load_library()
dev = torch.device("cuda")
rng = np.random.default_rng(0)
data_h = 10 + rng.normal(size=(nf, nt))
weights_h = np.ones((nf, nt)); weights_h[::7] = 0
data = torch.from_numpy(data_h).to(dev); weights = torch.from_numpy(weights_h).to(dev)
stream = torch.cuda.current_stream().cuda_stream
moments = torch.empty((nf, 4), dtype=torch.float64, device=dev)
sums = torch.empty((nf, 2), dtype=torch.float64, device=dev)
good = torch.ones(nf, dtype=torch.uint8, device=dev); good[::7] = 0
mean = torch.full((nf,), 10.0, dtype=torch.float64, device=dev)
out_ds = torch.empty((nf // 2, nt // 4), dtype=torch.float64, device=dev)
start = torch.randint(0, nt - 256, (nf,), dtype=torch.int32, device=dev)
out_w = torch.empty((nf, 256), dtype=torch.float64, device=dev)
B = nf * nt * 8
cases = {
    "k_pre_moments": (lambda: lib.fb_pre_channel_moments(data.data_ptr(), weights.data_ptr(), nf, nt, moments.data_ptr(), stream), 2 * B),
    "k_pre_baseline": (lambda: lib.fb_pre_remove_baseline(data.data_ptr(), weights.data_ptr(), good.data_ptr(), mean.data_ptr(), nf, nt, stream), 3 * B),
    "k_pre_power_sums": (lambda: lib.fb_pre_power_sums(data.data_ptr(), nf, nt, sums.data_ptr(), stream), B),
    "k_pre_downsample(2x4)": (lambda: lib.fb_pre_downsample_2d(data.data_ptr(), nf, nt, 2, 4, out_ds.data_ptr(), stream), B + B // 8),
    "k_pre_window(256)": (lambda: lib.fb_pre_window(data.data_ptr(), nf, nt, start.data_ptr(), 256, out_w.data_ptr(), stream), 2 * nf * 256 * 8),
}
for name, (fn, nbytes) in cases.items():
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(fn())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(json.dumps({"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "achieved_GBps": nbytes / ms / 1e6,
                      "peak_GBps": peak, "peak_source": src, "frac": nbytes / ms / 1e6 / peak}))
t0 = time.perf_counter(); po.preprocess(data_h[:2048], weights_h[:2048], remove_baseline=True); t1 = time.perf_counter()
print(json.dumps({"cpu_oracle_preprocess_s_full_size": (t1 - t0) * nf / 2048, "sample": "2048 of 16384 channels, scaled x8, 1 core"}))
