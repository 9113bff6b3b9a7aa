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
from tnpy_b200.matrix_product_state import Direction  # noqa: E402
from tnpy_b200.model import XXZ  # noqa: E402
from tnpy_b200.parallel import make_comm, sharded_eig_lowest, sharded_local_update  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=2048)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-matvec", type=int, default=1000)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--general", action="store_true", help="do not use the measured identity channels (general chain)")
    ap.add_argument("--update", action="store_true",
                    help="time one whole sharded local update (eigensolve + perturbation + split + environment update)")
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
    if args.update:
        lo = l // world
        L_rows = L[:, :, rank * lo:(rank + 1) * lo].contiguous()
        psi_rows = psi[rank * lo:(rank + 1) * lo].contiguous().clone()
        nb = env.device_tensor(site + 1)
        dist.barrier()
        upd = sharded_local_update(comm, L_rows, W, R, psi_rows, nb, l, tol=args.tol, flags=flags, max_matvec=args.max_matvec)
        ph = torch.tensor([upd["phase_s"][k] for k in ("eigensolve", "perturb", "gather_and_split", "env_update")],
                          dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ph, op=dist.ReduceOp.MAX)
        out["local_update"] = {"phase_s_max_over_ranks": dict(zip(("eigensolve", "perturb", "gather_and_split", "env_update"),
                                                                  [float(v) for v in ph])),
                               "total_s": float(ph.sum()), "n_matvec": upd["stats"]["n_matvec"], "theta": upd["stats"]["theta"]}
        if args.check:
            # gauge-independent checks of the three post-solve steps (the split's Q is a sensitive function of psi, so
            # it is not compared with another run's Q): the new site tensor is left-orthonormal, site x neighbour
            # still is psi x old neighbour, and the summed row-block environment update equals the unsharded update
            # of the *same* site tensor
            q = upd["site_tensor"].reshape(l * d, r)
            gram = _cuda.gemm_tn(q, q)
            orth = float((gram - torch.eye(r, dtype=torch.float64, device="cuda")).abs().max())
            psi_full = torch.empty((l, d, r), dtype=torch.float64, device="cuda")
            comm.allgather(psi_rows.contiguous(), psi_full)
            r2 = nb.shape[2]
            two_new = upd["site_tensor"].reshape(l * d, r) @ upd["neighbour"].reshape(r, d * r2)
            two_old = psi_full.reshape(l * d, r) @ nb.reshape(r, d * r2)
            state = float((two_new - two_old).abs().max() / two_old.abs().max())
            ref_next = _cuda.env_update_left(L, upd["site_tensor"], W, flags=flags & _cuda.LEFT_IDENTITY)
            ro = r // world
            d_env = float((upd["next_left_rows"] - ref_next[:, :, rank * ro:(rank + 1) * ro]).abs().max() / ref_next.abs().max())
            t0 = time.perf_counter()
            e_ref = dmrg._solve_on_device(site, args.tol, maxiter=args.max_matvec)
            dmrg.perturb_wave_function(site)
            env.split_tensor(site, Direction.RIGHTWARD)
            env.update(site, Direction.RIGHTWARD)
            torch.cuda.synchronize()
            out["local_update"]["unsharded_total_s"] = time.perf_counter() - t0
            diff2 = torch.tensor([d_env, orth, state, abs(e_ref - upd["stats"]["theta"])], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(diff2, op=dist.ReduceOp.MAX)
            out["local_update"].update(max_rel_diff_next_left_env=float(diff2[0]), site_tensor_orthogonality_defect=float(diff2[1]),
                                       two_site_state_rel_diff=float(diff2[2]), theta_diff=float(diff2[3]))
    if rank == 0:
        print(json.dumps(out), flush=True)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
