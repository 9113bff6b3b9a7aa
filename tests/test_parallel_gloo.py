"""world_size-2 (and 3) gloo runs of the chi-row sharding logic on CPU: the row blocks tile the bond,
all-gather reassembles x, and the sharded matvec (oracle arithmetic per rank) equals the full one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tnpy_oracle as oracle
from tnpy_b200.parallel import all_row_blocks, assign_realisations, row_block


@pytest.mark.parametrize("chi,world", [(8, 2), (60, 8), (2048, 8), (8192, 4), (7, 2), (10, 3), (3, 4)])
def test_row_blocks_tile_the_bond(chi, world):
    blocks = all_row_blocks(chi, world)
    assert blocks[0][0] == 0 and blocks[-1][1] == chi
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a1 >= a0
    sizes = [hi - lo for lo, hi in blocks]
    assert max(sizes) - min(sizes) <= 2
    if chi % 2 == 0 and chi >= 2 * world:
        assert all(s % 2 == 0 for s in sizes)
    with pytest.raises(ValueError):
        row_block(chi, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, chi, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tnpy_b200.parallel import gather_rows, row_block, shard_left_env, sharded_dot

        rng = np.random.default_rng(0)  # same operands on every rank
        w, d = 5, 2
        L = rng.standard_normal((chi, w, chi))
        R = rng.standard_normal((chi, w, chi))
        W = rng.standard_normal((w, w, d, d))
        x = rng.standard_normal((chi, d, chi))
        lo, hi = row_block(chi, world, rank)
        x_rows = torch.from_numpy(x[lo:hi].copy())
        x_full = gather_rows(x_rows, chi).numpy()
        assert np.array_equal(x_full, x)
        L_rows = shard_left_env(torch.from_numpy(L), world, rank).numpy()
        # per-rank chain with the row-sharded left environment (what tnpy_heff_apply_rows computes)
        t1 = np.tensordot(L_rows, x_full, axes=(0, 0))
        t2 = np.einsum("abpq,ampr->bmqr", W, t1, optimize=True)
        y_rows = np.einsum("bmqr,rbs->mqs", t2, R, optimize=True)
        y_full = oracle.heff_apply(L, W, R, x)
        err = np.abs(y_rows - y_full[lo:hi]).max() / np.abs(y_full).max()
        dot = float(sharded_dot(torch.from_numpy(y_rows), x_rows))
        q.put((rank, err, dot, float(np.vdot(y_full, x))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,chi", [(2, 12), (3, 10)])
def test_sharded_matvec_gloo(world, chi):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, chi, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, dot, dot_ref in results:
        assert err < 1e-13
        assert abs(dot - dot_ref) < 1e-10 * abs(dot_ref)


def test_realisations_partition():
    """configs[3]: 64 disorder realisations over 8 GPUs -- every seed exactly once, 8 per rank."""
    seen = []
    for rank in range(8):
        mine = assign_realisations(64, 8, rank)
        assert len(mine) == 8
        seen += mine
    assert sorted(seen) == list(range(64))
    assert assign_realisations(5, 2, 1) == [1, 3]
    with pytest.raises(ValueError):
        assign_realisations(4, 2, 2)
