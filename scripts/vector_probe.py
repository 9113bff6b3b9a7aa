"""HBM-roofline check of the eigensolver's vector kernels (run under gpurun).
Algorithmic bytes per BASELINE.md: dot 16N, nrm2 8N, axpy 24N, scal 16N, multi_dot 8N(m+ceil(m/4)) (w is
re-read once per group of 4 basis vectors), multi_axpy 8N(m+2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

PEAK = 6549.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


_cuda.load()
out = {"hbm_peak_gbs": PEAK}
# copy yardstick measured the same way as MEASURED_PEAKS.json
a = torch.empty(1 << 29, dtype=torch.float64, device="cuda").normal_()
b = torch.empty_like(a)
t = timed(lambda: b.copy_(a), 10)
out["torch_copy_4GiB_gbs"] = 2 * a.numel() * 8 / t / 1e9
del a, b
for label, n in (("chi2048_N8.4M", 2048 * 2048 * 2), ("chi8192_N134M", 8192 * 8192 * 2)):
    m = 8
    V = torch.randn((m, n), dtype=torch.float64, device="cuda")
    w = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    h = torch.randn(m, dtype=torch.float64, device="cuda") * 1e-3
    res = {}
    for name, fn, nbytes in (
        ("dot", lambda: _cuda.dot(w, y), 16 * n),
        ("nrm2", lambda: _cuda.nrm2(w), 8 * n),
        ("axpy", lambda: _cuda.axpy(1e-3, w, y), 24 * n),
        ("scal", lambda: _cuda.scal(1.0000001, y), 16 * n),
        (f"multi_dot_m{m}", lambda: _cuda.multi_dot(V, w), 8 * n * (m + (m + 3) // 4)),
        (f"multi_axpy_m{m}", lambda: _cuda.multi_axpy(V, h, y), 8 * n * (m + 2)),
    ):
        t = timed(fn)
        res[name] = {"us": t * 1e6, "gbs": nbytes / t / 1e9, "frac_of_hbm_peak": nbytes / t / 1e9 / PEAK}
    out[label] = res
    del V, w, y
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/vector_probe.json", "w"), indent=1)
