"""ShiftInvertDMRG beyond the Jacobi-dense range: pencils of 8 192 .. 32 768 unknowns through the Cholesky + Lanczos
route (tnpy_geig_chol_lowest).  Reports seconds per sweep, the sizes and Lanczos matvecs of the local pencils, the
energy per sweep and, at the end, the energy variance of the restored state on the unshifted Hamiltonian.

    python scripts/shift_invert_large.py --n 16 --chi 64 --sweeps 4 [--h 10.5] [--offset 0.1]
"""
import argparse
import json
import logging
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--chi", type=int, default=64)
    ap.add_argument("--h", type=float, default=10.5)
    ap.add_argument("--offset", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=2022)
    ap.add_argument("--sweeps", type=int, default=4)
    ap.add_argument("--tol", type=float, default=1e-8)
    args = ap.parse_args()

    import torch

    from tnpy_b200.finite_dmrg import ShiftInvertDMRG
    from tnpy_b200.matrix_product_state import Direction
    from tnpy_b200.model import RandomHeisenberg

    logging.getLogger("tnpy").setLevel(logging.WARNING)
    model = RandomHeisenberg(n=args.n, h=args.h, seed=args.seed)
    shifted = RandomHeisenberg(n=args.n, h=args.h, seed=args.seed, offset=args.offset)
    solver = ShiftInvertDMRG(shifted.mpo, bond_dim=args.chi, offset=args.offset, seed=1)
    out = {"model": f"RandomHeisenberg n={args.n} h={args.h} seed={args.seed}", "chi": args.chi, "offset": args.offset,
           "sweeps": []}
    plain_solve = solver._solve_on_device

    def timed_solve(site, tol, **kw):  # per-pencil seconds on stderr as they happen (a run cut short still tells)
        torch.cuda.synchronize()
        t = time.perf_counter()
        energy = plain_solve(site, tol, **kw)
        torch.cuda.synchronize()
        st = solver.solver_stats[-1]
        print(json.dumps({"site": site, "unknowns": int(solver.environment.device_tensor(site).numel()),
                          "seconds": round(time.perf_counter() - t, 4), "lanczos_matvecs": st.get("n_matvec", 0)}),
              file=sys.stderr, flush=True)
        return energy

    solver._solve_on_device = timed_solve
    for k in range(args.sweeps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lam = solver.sweep(Direction.RIGHTWARD if k % 2 == 0 else Direction.LEFTWARD, tol=args.tol)
        torch.cuda.synchronize()
        sizes = [int(solver.environment.device_tensor(st["site"]).numel()) for st in solver.solver_stats]
        out["sweeps"].append({
            "seconds": time.perf_counter() - t0,
            "energy": 1.0 / lam + args.offset,
            "largest_pencil": max(sizes),
            "pencils_over_2048": sum(1 for s in sizes if s > 2048),
            "lanczos_matvecs": [st.get("n_matvec", 0) for st in solver.solver_stats if st.get("n_matvec", 0) > 0],
        })
    solver._restore_mps()
    restored = solver.restored_mps

    def expectation(mpo):  # <psi|O|psi> by transfer matrices, on the device (bond chi * w: too slow for the host)
        e = torch.ones((1, 1, 1), dtype=torch.float64, device="cuda")
        for site in range(restored.n_sites):
            a = torch.from_numpy(restored.three_leg(site)).cuda()
            w = torch.from_numpy(mpo.as_four_leg(site)).cuda()
            t = torch.einsum("lam,lpr->ampr", e, a)
            t = torch.einsum("ampr,abpq->mrbq", t, w)
            e = torch.einsum("mrbq,mqs->rbs", t, a)
        return float(e[0, 0, 0])

    e1 = expectation(model.mpo)
    e2 = expectation(model.mpo.square())
    out["restored_state"] = {"energy": e1, "energy_from_the_pencil": out["sweeps"][-1]["energy"], "variance": e2 - e1 * e1}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
