"""A/B of the eigensolver's inexact-Krylov slice schedule: GEMM error by slice count against FP64, then
measured sweeps of XXZ n=40 chi=2048 with matvec, reduced-slice and failed-check counts and the energies.

    python scripts/inexact_slices_ab.py [--sweeps 3]                      # schedule on (default)
    TNPY_INEXACT_SLICES=0 python scripts/inexact_slices_ab.py             # every product at the base slice count
    TNPY_INEXACT_FACTOR=0.005 python scripts/inexact_slices_ab.py         # another safety factor (default 0.00125)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--sweeps", type=int, default=3)
    ap.add_argument("--no-gemm", action="store_true")
    args = ap.parse_args()

    import torch

    import bench
    from tnpy_b200 import _cuda as cu

    out = {"TNPY_INEXACT_SLICES": os.environ.get("TNPY_INEXACT_SLICES", "1"),
           "TNPY_INEXACT_FACTOR": os.environ.get("TNPY_INEXACT_FACTOR", "0.00125")}
    if not args.no_gemm:
        g = torch.Generator(device="cuda").manual_seed(0)
        a = torch.randn((2048, 1024), generator=g, dtype=torch.float64, device="cuda")
        b = torch.randn((2048, 512), generator=g, dtype=torch.float64, device="cuda")
        ref = a.t() @ b
        out["gemm_rel_err_by_slices"] = {s: float((cu.ozaki_gemm_tn(a, b, slices=s) - ref).abs().max() / ref.abs().max())
                                         for s in (5, 6, 7, 8)}
    out["sweeps"] = bench.measure_sweeps(args.n, args.chi, 1e-8, args.sweeps)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
