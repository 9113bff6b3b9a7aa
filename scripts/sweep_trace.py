"""Per-site local-solve statistics of late sweeps for both bond-split gauges (run under gpurun).

    python scripts/sweep_trace.py [--chi 512] [--n 40] [--sweeps 5] [--from-sweep 3]
"""
import argparse
import json
import logging
import os
import sys
from itertools import cycle

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=512)
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--sweeps", type=int, default=5)
    ap.add_argument("--from-sweep", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-8)
    args = ap.parse_args()
    import torch

    from bench import random_right_canonical_device
    from tnpy_b200 import model as models
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import Direction

    logging.getLogger("tnpy").setLevel(logging.ERROR)
    mdl = models.XXZ(n=args.n, delta=0.5)
    for split in ("svd", "qr"):
        init = random_right_canonical_device(args.n, args.chi, 2, seed=0)
        f = FiniteDMRG(mdl.mpo, bond_dim=args.chi, mps=init, compute_variance=False, split=split)
        for k, direction in zip(range(1, args.sweeps + 1), cycle([Direction.RIGHTWARD, Direction.LEFTWARD])):
            e = f.sweep(direction, tol=args.tol)
            torch.cuda.synchronize()
            rec = {"split": split, "sweep": k, "energy": e,
                   "matvecs": sum(s.get("n_matvec", 0) for s in f.solver_stats)}
            if k >= args.from_sweep:
                rec["sites"] = [[s["site"], s.get("n_matvec", 0), s.get("n_restart", 0), float("%.3g" % s.get("resid", 0.0)),
                                 float("%.6g" % s.get("anorm", 0.0))] for s in f.solver_stats if not s.get("dense")]
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
