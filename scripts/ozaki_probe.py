"""EXPERIMENT probe: accuracy and speed of the Ozaki int8 tcgen05 GEMM vs the DMMA GEMM (run under gpurun).

OZ_VARIANTS=2,1 (default) checks the CTA-pair two-pass kernel and the single-CTA kernel side by side.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

_cuda.load()
VARIANTS = [int(v) for v in os.environ.get("OZ_VARIANTS", "2,1").split(",")]
SLICES = [int(v) for v in os.environ.get("OZ_SLICES", "8,7,6").split(",")]


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
SHAPES = ((128, 64, 64), (256, 128, 64), (256, 192, 320), (300, 200, 130), (1024, 1024, 1024), (520, 390, 4160))
for variant in VARIANTS:
    _cuda.set_ozaki_variant(variant)
    for (m, n, k) in SHAPES:
        a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
        b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
        a = a * torch.logspace(0, -6, m, dtype=torch.float64, device="cuda")[None, :]  # graded columns
        ref = a.t() @ b
        bound = a.abs().t() @ b.abs()
        row = {"variant": variant, "shape": [m, n, k]}
        for s in SLICES:
            c = _cuda.ozaki_gemm_tn(a, b, slices=s)
            row[f"err_s{s}"] = ((c - ref).abs() / bound).max().item()  # componentwise relative to |A|^T|B|
            c0 = torch.randn_like(c)
            c1 = _cuda.ozaki_gemm_tn(a, b, out=c0.clone(), slices=s, accumulate=True)
            row[f"acc_err_s{s}"] = ((c1 - c0 - c).abs() / (bound + c0.abs())).max().item()
        row["err_dmma"] = ((_cuda.gemm_tn(a, b) - ref).abs() / bound).max().item()
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
    for variant in VARIANTS:
        _cuda.set_ozaki_variant(variant)
        for s in SLICES:
            _cuda.ozaki_gemm_tn(a, b, out=c, slices=s, phase=1)
            t_mm = timed(lambda: _cuda.ozaki_gemm_tn(a, b, out=c, slices=s, phase=2))
            t_all = timed(lambda: _cuda.ozaki_gemm_tn(a, b, out=c, slices=s, phase=0))
            row[f"v{variant}_s{s}"] = {
                "mma_only_ms": t_mm * 1e3,
                "mma_only_tflops_equiv": flops / t_mm / 1e12,
                "with_slicing_tflops_equiv": flops / t_all / 1e12,
                "max_rel_diff_vs_dmma": ((c - ref).abs().max() / ref.abs().max()).item(),
            }
    print(json.dumps(row), flush=True)
