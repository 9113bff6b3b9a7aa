"""EXPERIMENT probe: accuracy and speed of the Ozaki int8 tcgen05 GEMM vs the DMMA GEMM (run under gpurun)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

_cuda.load()


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


g = torch.Generator(device="cuda").manual_seed(0)
for (m, n, k) in ((128, 64, 64), (256, 192, 320), (300, 200, 130), (1024, 1024, 1024)):
    a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    a = a * torch.logspace(0, -6, m, dtype=torch.float64, device="cuda")[None, :]  # graded columns
    ref = a.t() @ b
    row = {"shape": [m, n, k]}
    for s in (6, 7, 8):
        c = _cuda.ozaki_gemm_tn(a, b, slices=s)
        err = ((c - ref).abs() / (a.abs().t() @ b.abs())).max().item()  # componentwise relative to |A|^T|B|
        row[f"err_s{s}"] = err
    row["err_dmma"] = ((_cuda.gemm_tn(a, b) - ref).abs() / (a.abs().t() @ b.abs())).max().item()
    print(json.dumps(row), flush=True)

for name, (m, n, k) in {"gemm1_chi2048": (4096, 10240, 2048), "gemm3_chi2048": (4096, 2048, 10240)}.items():
    a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    c = torch.empty((m, n), dtype=torch.float64, device="cuda")
    flops = 2.0 * m * n * k
    row = {"shape": name}
    t = timed(lambda: _cuda.gemm_tn(a, b, out=c))
    row["dmma_tflops"] = flops / t / 1e12
    ref = c.clone()
    for s in (6, 7, 8):
        _cuda.ozaki_gemm_tn(a, b, out=c, slices=s, phase=1)
        t_mm = timed(lambda: _cuda.ozaki_gemm_tn(a, b, out=c, slices=s, phase=2))
        t_all = timed(lambda: _cuda.ozaki_gemm_tn(a, b, out=c, slices=s, phase=0))
        row[f"s{s}"] = {"mma_only_tflops_equiv": flops / t_mm / 1e12, "with_slicing_tflops_equiv": flops / t_all / 1e12,
                        "max_rel_diff_vs_dmma": ((c - ref).abs().max() / ref.abs().max()).item()}
    print(json.dumps(row), flush=True)
