"""Time the on-device bond SVD at the bench shapes (run under gpurun)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

lib = _cuda.load()
SHAPES = ((120, 60, "rand"), (512, 256, "rand"), (2048, 1024, "rand"), (4096, 2048, "rand"), (4096, 2048, "near"), (2048, 4096, "rand"))
if os.environ.get("SVD_PROBE_BIG_ONLY"):
    SHAPES = ((4096, 2048, "rand"), (4096, 2048, "near"))
for rows, cols, kind in SHAPES:
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda")
    if kind == "near":  # nearly orthogonal columns, as in late DMRG sweeps
        q, _ = torch.linalg.qr(a)
        a = q * torch.logspace(0, -10, cols, dtype=torch.float64, device="cuda")[None, :]
        a = (a + 1e-6 * torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda") * a.abs().mean()).contiguous()
    _cuda.svd(a.clone())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    u, s, vt = _cuda.svd(a.clone())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt_ref = float("nan")
    if not os.environ.get("SVD_PROBE_NO_REF"):
        t1 = time.perf_counter()
        torch.linalg.svd(a, full_matrices=False)
        torch.cuda.synchronize()
        dt_ref = time.perf_counter() - t1
    err = float(((u * s) @ vt - a).abs().max() / s[0])
    print(json.dumps({"rows": rows, "cols": cols, "kind": kind, "ms": dt * 1e3, "sweeps": lib.tnpy_last_svd_sweeps(),
                      "recon_err": err, "cusolver_gesvd_ms": dt_ref * 1e3}), flush=True)
