"""One-shot hardware / kernel probe for the B200 box (run under gpurun).

Prints JSON lines: the cuBLAS FP64 GEMM rate (the FP64 roofline denominator the profiling recipe
asks for), the register-resident DMMA / DFMA peaks, and this library's GEMM and H_eff matvec rates at
the benchmark shapes.  Writes gpurun_out/probe.json.
"""
import ctypes
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

OUT = {}


def timed(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), sorted(ts)[len(ts) // 2]


def emit(key, value):
    OUT[key] = value
    print(json.dumps({key: value}), flush=True)


def main():
    lib = _cuda.load()
    dev = torch.device("cuda")
    emit("gpu", torch.cuda.get_device_name(0))
    # ---- cuBLAS dgemm (library yardstick, not on the product path)
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        best, med = timed(lambda: torch.matmul(a, b), warmup=2, iters=5)
        emit(f"cublas_dgemm_{n}_tflops", {"best": 2 * n**3 / best / 1e12, "median": 2 * n**3 / med / 1e12})
        del a, b
    # sustained: back-to-back for ~3 s
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cnt = 0
    while time.perf_counter() - t0 < 3.0:
        for _ in range(4):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        cnt += 4
    emit("cublas_dgemm_8192_sustained_tflops", 2 * n**3 * cnt / (time.perf_counter() - t0) / 1e12)
    del a, b
    # ---- register-resident probes
    scratch = torch.zeros(16, dtype=torch.float64, device=dev)
    res = ctypes.c_double()
    for kind, name in ((0, "dmma"), (1, "dfma")):
        for tpb, bps, ilp in ((128, 1, 8), (128, 1, 32), (256, 1, 16), (256, 1, 32), (512, 1, 16), (256, 2, 16), (1024, 1, 8)):
            rc = lib.tnpy_probe_fp64(kind, tpb, bps, ilp, 20000, ctypes.byref(res), scratch.data_ptr())
            emit(f"probe_{name}_t{tpb}_b{bps}_ilp{ilp}_tflops", res.value if rc == 0 else _cuda.last_error())
    # ---- this library's GEMM at the chi=2048 matvec shapes
    shapes = {"gemm1_chi2048": (4096, 10240, 2048), "gemm3_chi2048": (4096, 2048, 10240),
              "gemm1_chi1024": (2048, 5120, 1024), "gemm3_chi1024": (2048, 1024, 5120),
              "gemm1_chi256": (512, 1536, 256), "gemm3_chi256": (512, 256, 1536)}
    for name, (m, n, k) in shapes.items():
        a = torch.randn(k, m, dtype=torch.float64, device=dev)
        b = torch.randn(k, n, dtype=torch.float64, device=dev)
        c = torch.empty(m, n, dtype=torch.float64, device=dev)
        ref = None
        for tile in (0, 1, 2, -1):
            lib.tnpy_set_gemm_tile(tile)
            try:
                best, med = timed(lambda: _cuda.gemm_tn(a, b, out=c, algo=_cuda.GEMM_DMMA), warmup=1, iters=3)
            finally:
                lib.tnpy_set_gemm_tile(-1)
            if ref is None:
                ref = torch.matmul(a.t(), b)
            err = float((c - ref).abs().max() / ref.abs().max())
            emit(f"{name}_tile{tile}", {"tflops_best": 2 * m * n * k / best / 1e12, "tflops_median": 2 * m * n * k / med / 1e12, "rel_err": err})
        best, med = timed(lambda: torch.matmul(a.t(), b, out=c), warmup=1, iters=3)
        emit(f"{name}_cublas", {"tflops_best": 2 * m * n * k / best / 1e12})
        del a, b, c, ref
    # ---- H_eff matvec
    for chi, w in ((256, 6), (1024, 5), (2048, 5)):
        d = 2
        L = torch.randn(chi, w, chi, dtype=torch.float64, device=dev)
        R = torch.randn(chi, w, chi, dtype=torch.float64, device=dev)
        W = torch.randn(w, w, d, d, dtype=torch.float64, device=dev)
        x = torch.randn(chi, d, chi, dtype=torch.float64, device=dev)
        y = torch.empty_like(x)
        best, med = timed(lambda: _cuda.heff_apply(L, W, R, x, y), warmup=2, iters=5)
        flops = 4 * w * d * chi**3 + 2 * w * w * d * d * chi**2
        emit(f"heff_chi{chi}", {"ms_best": best * 1e3, "ms_median": med * 1e3, "tflops_best": flops / best / 1e12, "tflops_median": flops / med / 1e12})
        del L, R, W, x, y
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(OUT, f, indent=1)


if __name__ == "__main__":
    main()
