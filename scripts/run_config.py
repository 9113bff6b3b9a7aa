"""Run a BASELINE.json configuration end to end on the GPU and (optionally) the CPU oracle from the
same initial MPS; report energies per sweep, wall time, overlap and singular-value differences.

    python scripts/run_config.py --config 1 [--sweeps 4] [--oracle]
      1: XXZ n=100 delta=0.5 chi=60 tol=1e-8          (README example)
      2: Thirring n=100 chi=256 (scripts/thirring_fdmrg.py parameters)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tnpy_oracle as oracle  # noqa: E402  (checker only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--chi", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=0, help="fixed number of sweeps (0 = run to |dE| < tol)")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ncv", type=int, default=0, help="Lanczos basis size (0 = library default)")
    ap.add_argument("--split", default="qr", choices=["qr", "svd"], help="bond split inside the sweeps")
    args = ap.parse_args()

    import logging

    import torch

    from tnpy_b200 import _cuda, model as models
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import MatrixProductState

    logging.getLogger("tnpy").setLevel(logging.WARNING)
    if args.config == 1:
        n, chi = args.n or 100, args.chi or 60
        mdl = models.XXZ(n=n, delta=0.5)
    elif args.config == 2:
        n, chi = args.n or 100, args.chi or 256
        mdl = models.Thirring(n=n, delta=0.5, ma=1.0, penalty=100.0, s_target=0)
    else:  # 3: XXZ n=100 chi=2048 (the roofline configuration); device-born synthetic start
        n, chi = args.n or 100, args.chi or 2048
        mdl = models.XXZ(n=n, delta=0.5)
    if args.config == 3:
        from bench import random_right_canonical_device

        init = None
        device_init = random_right_canonical_device(n, chi, 2, seed=args.seed)
    else:
        init = oracle.random_mps(n, chi, 2, seed=args.seed)
    kw = dict(tol=args.tol)
    if args.sweeps:
        kw.update(tol=args.tol, max_sweep=args.sweeps)
    out = {"config": args.config, "n": n, "chi": chi, "tol": args.tol}

    if init is None:
        gpu = FiniteDMRG(mdl.mpo, bond_dim=chi, mps=device_init, compute_variance=False, split=args.split)
        del device_init
    else:
        gpu = FiniteDMRG(mdl.mpo, bond_dim=chi, mps=MatrixProductState([a.copy() for a in init]), split=args.split)
    gpu.phase_seconds = {}
    torch.cuda.synchronize()
    l0 = _cuda.launch_count()
    t0 = time.perf_counter()
    if args.sweeps:
        from itertools import cycle

        from tnpy_b200.matrix_product_state import Direction

        e_gpu, per_sweep = [], []
        for _, direction in zip(range(args.sweeps), cycle([Direction.RIGHTWARD, Direction.LEFTWARD])):
            t = time.perf_counter()
            e_gpu.append(gpu.sweep(direction, tol=args.tol, **({"ncv": args.ncv} if args.ncv else {})))
            torch.cuda.synchronize()
            per_sweep.append(time.perf_counter() - t)
            out.setdefault("gpu_matvecs_per_sweep", []).append(sum(s.get("n_matvec", 0) for s in gpu.solver_stats))
            out.setdefault("gpu_looks_per_sweep", []).append(sum(s.get("looks", 0) for s in gpu.solver_stats))
            out.setdefault("gpu_extra_gs_passes_per_sweep", []).append(sum(s.get("extra_gs_passes", 0) for s in gpu.solver_stats))
            out.setdefault("gpu_reduced_slice_matvecs_per_sweep", []).append(sum(s.get("reduced_slice_matvecs", 0) for s in gpu.solver_stats))
            out.setdefault("gpu_failed_residual_checks_per_sweep", []).append(sum(s.get("failed_residual_checks", 0) for s in gpu.solver_stats))
            out.setdefault("gpu_phase_s_cumulative", []).append(dict(gpu.phase_seconds))
        out["gpu_sweep_s"] = per_sweep
        out["gpu_matvecs_last_sweep"] = sum(s.get("n_matvec", 0) for s in gpu.solver_stats)
    else:
        totals = {"n_matvec": 0, "looks": 0, "sweeps": 0}
        plain_sweep = gpu.sweep

        def counted_sweep(*a, **k):
            energy = plain_sweep(*a, **k)
            totals["n_matvec"] += sum(s.get("n_matvec", 0) for s in gpu.solver_stats)
            totals["looks"] += sum(s.get("looks", 0) for s in gpu.solver_stats)
            totals["sweeps"] += 1
            return energy

        gpu.sweep = counted_sweep
        e_gpu = gpu.run(**kw)
        out["gpu_totals"] = totals
        out["gpu_phase_s"] = dict(gpu.phase_seconds)  # synchronised per-phase wall seconds, summed over the run
    torch.cuda.synchronize()
    out["gpu_wall_s"] = time.perf_counter() - t0
    out["gpu_energies"] = e_gpu
    out["gpu_launches"] = _cuda.launch_count() - l0
    out["split"] = args.split
    out["ncv"] = args.ncv
    out["split_counts"] = dict(gpu.environment.split_counts)

    if args.oracle:
        ref = oracle.FiniteDMRG(mdl.mpo.arrays, chi, mps=[a.copy() for a in init])
        t0 = time.perf_counter()
        if args.sweeps:
            from itertools import cycle

            e_ref, per_sweep = [], []
            for _, direction in zip(range(args.sweeps), cycle([oracle.RIGHTWARD, oracle.LEFTWARD])):
                t = time.perf_counter()
                e_ref.append(ref.sweep(direction, tol=args.tol))
                per_sweep.append(time.perf_counter() - t)
            out["cpu_sweep_s"] = per_sweep
        else:
            e_ref = ref.run(with_variance=False, **kw)
        out["cpu_wall_s"] = time.perf_counter() - t0
        out["cpu_energies"] = e_ref
        out["cpu_cores"] = os.cpu_count()
        out["cpu_matvecs"] = ref.n_matvec
        a, b = ref.mps, gpu.mps.arrays
        ov = abs(oracle.mps_overlap(a, b)) / np.sqrt(oracle.mps_overlap(a, a) * oracle.mps_overlap(b, b))
        out["one_minus_overlap"] = 1 - ov
        out["energy_rel_diff"] = abs(e_gpu[-1] - e_ref[-1]) / abs(e_ref[-1])
        sv = gpu.bond_singular_values
        out["max_singular_value_diff"] = max(float(np.abs(sv[b_] - s).max()) for b_, s in ref.bond_singular_values.items())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
