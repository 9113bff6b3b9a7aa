"""BASELINE configs[3]: random-disorder Heisenberg, a batch of disorder realisations over the GPUs of
one box -- replicas only (one process per GPU, no collective on the data path).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_disorder_batch.py \
        --sites 64 --chi 1024 --realisations 64 --sweeps 2
    python scripts/run_disorder_batch.py --sites 16 --chi 32 --realisations 4        # single GPU
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
    ap.add_argument("--sites", "--n", dest="n", type=int, default=64,
                    help="chain length (spell it --sites under torchrun, whose own parser finds --n ambiguous)")
    ap.add_argument("--chi", type=int, default=1024)
    ap.add_argument("--h", type=float, default=1.0)
    ap.add_argument("--realisations", type=int, default=64)
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--verify", action="store_true",
                    help="rank 0 also runs 4 realisations at n=16 chi=32 (6 sweeps) on the GPU and on the CPU oracle and reports the largest energy difference")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import Direction
    from tnpy_b200.model import RandomHeisenberg
    from tnpy_b200.parallel import assign_realisations

    logging.getLogger("tnpy").setLevel(logging.WARNING)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine = assign_realisations(args.realisations, world, rank)
    results = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for seed in mine:
        model = RandomHeisenberg(n=args.n, h=args.h, seed=seed)
        dmrg = FiniteDMRG(model.mpo, bond_dim=args.chi, seed=seed, compute_variance=False)
        energy = None
        for k in range(args.sweeps):
            energy = dmrg.sweep(Direction.RIGHTWARD if k % 2 == 0 else Direction.LEFTWARD, tol=args.tol)
        results[seed] = energy
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (results, elapsed))  # the only communication: final gather
        dist.destroy_process_group()
    else:
        gathered = [(results, elapsed)]
    verify = None
    if rank == 0 and args.verify:
        from oracle import tnpy_oracle as oracle  # checker only
        from tnpy_b200.matrix_product_state import MatrixProductState

        worst = 0.0
        for seed in range(4):
            model = RandomHeisenberg(n=16, h=args.h, seed=seed)
            init = oracle.random_mps(16, 32, 2, seed=seed)
            ref = oracle.FiniteDMRG(model.mpo.arrays, 32, mps=[a.copy() for a in init])
            gpu = FiniteDMRG(model.mpo, bond_dim=32, mps=MatrixProductState([a.copy() for a in init]), compute_variance=False)
            e_ref = e_gpu = None
            for k in range(6):
                e_ref = ref.sweep(oracle.RIGHTWARD if k % 2 == 0 else oracle.LEFTWARD, tol=1e-12)
                e_gpu = gpu.sweep(Direction.RIGHTWARD if k % 2 == 0 else Direction.LEFTWARD, tol=1e-12)
            worst = max(worst, abs(e_gpu - e_ref) / abs(e_ref))
        verify = {"realisations": 4, "n": 16, "chi": 32, "sweeps": 6, "max_rel_energy_diff_vs_oracle": worst}
    if rank == 0:
        energies = {}
        for res, _ in gathered:
            energies.update(res)
        wall = max(t for _, t in gathered)
        print(json.dumps({
            "config": "RandomHeisenberg n=%d chi=%d h=%g, %d realisations on %d GPU(s), %d sweeps each" % (
                args.n, args.chi, args.h, args.realisations, world, args.sweeps),
            "wall_s": wall, "realisations_per_s": args.realisations / wall,
            "energies": [energies[s] for s in sorted(energies)], "verify": verify,
        }), flush=True)


if __name__ == "__main__":
    main()
