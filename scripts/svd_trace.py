"""Jacobi-sweep statistics of the bond SVDs inside warm DMRG sweeps (run under gpurun)."""
import ctypes
import json
import logging
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import random_right_canonical_device  # noqa: E402
from tnpy_b200 import _cuda  # noqa: E402
from tnpy_b200.finite_dmrg import FiniteDMRG  # noqa: E402
from tnpy_b200.matrix_product_state import Direction  # noqa: E402
from tnpy_b200.model import XXZ  # noqa: E402

logging.getLogger("tnpy").setLevel(logging.WARNING)
n, chi = int(sys.argv[1]), int(sys.argv[2])
lib = _cuda.load()
dmrg = FiniteDMRG(XXZ(n=n, delta=0.5).mpo, bond_dim=chi, mps=random_right_canonical_device(n, chi, 2, 0), compute_variance=False)
env = dmrg.environment
orig = env.split_tensor
records = []


def traced(site, direction):
    torch.cuda.synchronize()
    t = time.perf_counter()
    s = orig(site, direction)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    buf = (ctypes.c_uint * 64)()
    k = lib.tnpy_last_svd_trace(buf, 64)
    records.append({"site": site, "ms": dt * 1e3, "sweeps": lib.tnpy_last_svd_sweeps(), "trace": [("c:" if v & 0x80000000 else "") + str(v & 0x7FFFFFFF) for v in buf[:k]],
                    "s_min": float(s.min()), "s_max": float(s.max())})
    return s


env.split_tensor = traced
for i, direction in enumerate((Direction.RIGHTWARD, Direction.LEFTWARD, Direction.RIGHTWARD)):
    records.clear()
    t = time.perf_counter()
    e = dmrg.sweep(direction, tol=1e-8)
    torch.cuda.synchronize()
    mid = [r for r in records if r["sweeps"] >= 0]
    print(json.dumps({"sweep": i, "energy": e, "s": time.perf_counter() - t, "matvecs": sum(x.get("n_matvec", 0) for x in dmrg.solver_stats),
                      "svd_ms_total": sum(r["ms"] for r in records), "block_svds": len(mid),
                      "sample": mid[len(mid) // 2] if mid else None}), flush=True)
