"""chi-sharded local eigensolve (BASELINE configs[4]) -- launched with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 \
        scripts/sharded_solve.py --chi 2048 [--n 26] [--check] [--tol 1e-8] [--max-matvec 60]

Every rank builds the same XXZ environment at a mid-chain site from the same seed (bit-identical kernels), brings the
state into the mixed-canonical form of a sweep, keeps its row block of L and of the site tensor, and the ranks solve
the local problem together (tnpy_eig_lowest_rows: all-gather of the Lanczos vector + all-reduces of the Gram-Schmidt
coefficients per step, everything else local).  --check also runs the unsharded tnpy_eig_lowest on every rank and
compares energy, residual and the eigenvector rows.  Rank 0 prints one JSON line (device time: max over ranks)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import f_mv, mixed_canonicalize, random_right_canonical_device  # noqa: E402
from tnpy_b200 import _cuda  # noqa: E402
from tnpy_b200.finite_dmrg import FiniteDMRG  # noqa: E402
from tnpy_b200.model import XXZ  # noqa: E402
from tnpy_b200.parallel import make_comm, sharded_eig_lowest  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-matvec", type=int, default=1000)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--general", action="store_true", help="do not use the measured identity channels (general chain)")
    args = ap.parse_args()
    import logging

    logging.getLogger("tnpy").setLevel(logging.WARNING)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo" if world == 1 else "nccl", **({} if world == 1 else {"device_id": torch.device("cuda", local)}))
    _cuda.load()
    comm = make_comm()
    chi = args.chi
    n = args.n or max(12, 2 * (chi.bit_length() - 1) + 2)  # shortest chain whose middle bond reaches chi
    dmrg = FiniteDMRG(XXZ(n=n, delta=0.5).mpo, bond_dim=chi, mps=random_right_canonical_device(n, chi, 2, seed=0),
                      compute_variance=False)
    env = dmrg.environment
    site = n // 2
    mixed_canonicalize(dmrg, site)
    L, W, R = env.operands(site)
    psi = env.device_tensor(site)
    l, d, r = psi.shape
    flags = 0 if args.general else env.gauge_flags(site)
    # warm-up: NCCL sets its transports up lazily on the first collectives, the library configures its kernels
    sharded_eig_lowest(comm, L, W, R, psi, tol=args.tol, flags=flags, image=True, max_matvec=3)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stats, psi_rows, image_rows = sharded_eig_lowest(comm, L, W, R, psi, tol=args.tol, flags=flags, image=True,
                                                     max_matvec=args.max_matvec)
    e1.record()
    torch.cuda.synchronize()
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    comm.allreduce_sum(dt.clone())  # exercise the exposed collective too
    dts = [torch.zeros_like(dt) for _ in range(world)]
    if world > 1:
        dist.all_gather(dts, dt)
    else:
        dts = [dt]
    out = {"world": world, "n": n, "chi": chi, "site": site, "shape": [l, d, r], "flags": flags, "tol": args.tol,
           "theta": stats["theta"], "resid": stats["resid"], "n_matvec": stats["n_matvec"], "converged": stats["converged"],
           "heff_mode": stats["heff_mode"], "slices": stats["slices"], "int8_error_bound": stats["int8_error_bound"],
           "solve_s_max_over_ranks": max(float(t.item()) for t in dts),
           "tflops_algorithmic": f_mv(l, r, W.shape[0], W.shape[1], d) * stats["n_matvec"] / max(float(t.item()) for t in dts) / 1e12}
    if args.check:
        ref = psi.clone()
        image = torch.empty_like(ref)
        _cuda.eig_lowest(L, W, R, psi.clone(), tol=args.tol, flags=flags, max_matvec=3)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref_stats = _cuda.eig_lowest(L, W, R, ref, tol=args.tol, flags=flags, image=image, max_matvec=args.max_matvec)
        torch.cuda.synchronize()
        out["unsharded_solve_s"] = time.perf_counter() - t0
        lo = rank * (l // world)
        rows = ref[lo:lo + l // world]
        sign = 1.0 if float((rows * psi_rows).sum()) >= 0 else -1.0
        diff = torch.tensor([float((rows - sign * psi_rows).abs().max()), float((image[lo:lo + l // world] - sign * image_rows).abs().max())],
                            dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        out.update(unsharded_theta=ref_stats["theta"], unsharded_n_matvec=ref_stats["n_matvec"],
                   theta_diff=abs(ref_stats["theta"] - stats["theta"]), max_abs_diff_psi_rows=float(diff[0].item()),
                   max_abs_diff_image_rows=float(diff[1].item()))
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
