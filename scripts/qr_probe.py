"""Time the verified Cholesky-QR bond split against the Jacobi SVD split at the bench shapes (run under gpurun)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

_cuda.load()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


for rows, cols in ((512, 256), (2048, 1024), (4096, 2048), (2048, 4096), (8192, 4096)):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda")
    k = min(rows, cols)
    grade = torch.logspace(0, -12, k, dtype=torch.float64, device="cuda")
    a = (a * grade[None, :] if rows >= cols else a * grade[:, None]).contiguous()
    l0 = _cuda.launch_count()
    q, t, defect = _cuda.qr_split(a)
    launches = _cuda.launch_count() - l0
    back = q @ t if rows >= cols else t @ q
    rec = float((back - a).abs().max() / a.abs().max())
    ms_qr = timed(lambda: _cuda.qr_split(a))
    ms_qr3 = timed(lambda: _cuda.qr_split(a, shifted=True))
    ms_svd = timed(lambda: _cuda.svd(a.clone()), reps=1) if rows * cols <= 4096 * 2048 else float("nan")
    ms_lib = timed(lambda: torch.linalg.qr(a if rows >= cols else a.t()), reps=1)
    flops = 6 * 2.0 * k * k * max(rows, cols)
    print(json.dumps({"rows": rows, "cols": cols, "qr_split_ms": ms_qr, "defect": defect, "recon_err": rec,
                      "launches": launches, "qr_split_shifted_ms": ms_qr3, "gemm_tflops_if_all_gemm": flops / ms_qr / 1e9,
                      "jacobi_svd_ms": ms_svd, "cusolver_geqrf_orgqr_ms": ms_lib}), flush=True)
