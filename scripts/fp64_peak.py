"""FP64 roofline denominators measured on the box (MEASURED_PEAKS.json carries none): cuBLAS dgemm 8192^3 burst
(best of 10) and sustained (back to back for 4 s), the register-resident DMMA (mma.sync m8n8k4.f64) and DFMA probes
of this library, with the SM clock sampled during the sustained run.  Writes gpurun_out/r02_fp64_peak.json; the
tracked copy is profiles/r02_fp64_peak.json (cited by bench.py's roofline.peak_source)."""
import ctypes
import json
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402


def main():
    lib = _cuda.load()
    out = {"gpu": torch.cuda.get_device_name(0), "how": __doc__.split("\n\n")[0].replace("\n", " ")}
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    out["cublas_dgemm_8192_burst_tflops"] = 2 * n**3 / min(ts) / 1e12
    smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                           stdout=subprocess.PIPE, text=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cnt = 0
    while time.perf_counter() - t0 < 4.0:
        for _ in range(4):
            torch.matmul(a, b)
        torch.cuda.synchronize()
        cnt += 4
    out["cublas_dgemm_8192_sustained_tflops"] = 2 * n**3 * cnt / (time.perf_counter() - t0) / 1e12
    smi.terminate()
    rows = [r.split(",") for r in smi.stdout.read().strip().splitlines() if "," in r]
    clocks = sorted(float(r[0]) for r in rows)
    out["sm_mhz_during_sustained"] = {"median": clocks[len(clocks) // 2] if clocks else None, "samples": len(clocks),
                                      "power_w_max": max((float(r[1]) for r in rows), default=None)}
    del a, b
    scratch = torch.zeros(16, dtype=torch.float64, device="cuda")
    res = ctypes.c_double()
    for kind, name in ((0, "dmma"), (1, "dfma")):
        best = 0.0
        for tpb, bps, ilp in ((128, 1, 32), (256, 1, 16), (256, 1, 32), (512, 1, 16), (256, 2, 16), (1024, 1, 8)):
            if lib.tnpy_probe_fp64(kind, tpb, bps, ilp, 20000, ctypes.byref(res), scratch.data_ptr()) == 0:
                best = max(best, res.value)
        out[f"probe_{name}_register_resident_tflops"] = best
    out["hardware_dmma_peak_tflops_at_1957_mhz_ncu"] = 37.07  # profiles/r01_gemm_tn_dmma_ncu_full_raw.csv
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r02_fp64_peak.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
