"""chi-row sharding of the H_eff matvec across the GPUs of one box.

Rank g of G owns the row block ``rows_g`` of the bra bond: ``L[:, :, rows_g]`` (stored contiguously),
the matching rows of x / y and of every Lanczos vector, plus a full copy of R.  One matvec is

    all-gather(x rows) -> x           (torch.distributed, NCCL over NVLink; the only exchange step)
    y[rows_g] = tnpy_heff_apply_rows(L[:, :, rows_g], W, R, x)        (no reduction needed)

Scalars of the eigensolver (dots, norms) are all-reduced.  The host-side index logic lives here so it
can be tested on CPU with the gloo backend.
"""
from __future__ import annotations

from typing import List, Tuple


def row_block(chi: int, world: int, rank: int) -> Tuple[int, int]:
    """Half-open row range [lo, hi) of ``rank``: blocks differ by at most one row, 2-row aligned
    when chi allows it (keeps every shard's leading dimension even for the TMA path)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    unit = 2 if chi % 2 == 0 and chi >= 2 * world else 1
    n_units = chi // unit
    base, extra = divmod(n_units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return lo_u * unit, hi_u * unit


def all_row_blocks(chi: int, world: int) -> List[Tuple[int, int]]:
    return [row_block(chi, world, g) for g in range(world)]


def shard_left_env(L, world: int, rank: int):
    """Contiguous copy of this rank's bra-row block of a left environment (l, w, l)."""
    lo, hi = row_block(L.shape[2], world, rank)
    return L[:, :, lo:hi].contiguous()


def gather_rows(x_rows, chi: int, group=None):
    """All-gather the row blocks of a site tensor into the full (chi, d, r) tensor on every rank."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = all_row_blocks(chi, world)
    full = torch.empty((chi,) + tuple(x_rows.shape[1:]), dtype=x_rows.dtype, device=x_rows.device)
    assert blocks[rank][1] - blocks[rank][0] == x_rows.shape[0]
    if all(hi - lo == blocks[0][1] - blocks[0][0] for lo, hi in blocks):
        dist.all_gather_into_tensor(full, x_rows.contiguous(), group=group)
        return full
    # ragged blocks: pad every block to the largest one, gather, then drop the padding
    most = max(hi - lo for lo, hi in blocks)
    padded = torch.zeros((most,) + tuple(x_rows.shape[1:]), dtype=x_rows.dtype, device=x_rows.device)
    padded[: x_rows.shape[0]] = x_rows
    stacked = torch.empty((world * most,) + tuple(x_rows.shape[1:]), dtype=x_rows.dtype, device=x_rows.device)
    dist.all_gather_into_tensor(stacked, padded, group=group)
    for g, (lo, hi) in enumerate(blocks):
        full[lo:hi] = stacked[g * most : g * most + hi - lo]
    return full


def sharded_dot(a_rows, b_rows, group=None):
    """Global <a|b> from row blocks: local partial + all-reduce of one scalar."""
    import torch
    import torch.distributed as dist

    part = (a_rows * b_rows).sum().reshape(1)
    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part[0]


def assign_realisations(n_realisations: int, world: int, rank: int) -> List[int]:
    """Disorder batches (BASELINE configs[3]) are replicas only: realisation s runs on rank s % world,
    no collective on the data path; results are gathered once at the end."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_realisations, world))


# ---------------------------------------------------------------------------------------------------
# chi-sharded local solve (the eigensolve of one DMRG site over the GPUs of a box)
# ---------------------------------------------------------------------------------------------------
def make_comm(group=None):
    """The library's own NCCL communicator for this process (``tnpy_comm_*``), created on the current CUDA device
    from an initialised ``torch.distributed`` group; the 128-byte NCCL id travels by an object broadcast, so the
    group's backend can be gloo or nccl."""
    from tnpy_b200 import _cuda

    return _cuda.Comm.from_torch_distributed(group)


def sharded_eig_lowest(comm, L, W, R, psi, tol: float = 1e-8, flags: int = 0, image: bool = False, **opts):
    """Lowest eigenpair of H_eff(L, W, R) with the bra rows of the left bond split evenly over ``comm``'s ranks.

    Every rank passes the *full* device tensors it holds (L (l, wl, l), psi (l, d, r) as start vector; W and R are
    needed in full anyway); this helper cuts out the rank's row block, runs ``tnpy_eig_lowest_rows`` -- the whole
    Lanczos iteration on the devices, one all-gather of the current vector and a few all-reduces of <= 64 doubles per
    step over NVLink -- and returns ``(stats, psi_rows, image_rows)`` with this rank's rows of the eigenvector (and of
    H_eff psi when ``image``).  Callers that keep their data sharded call ``_cuda.eig_lowest_rows`` directly."""
    import torch

    from tnpy_b200 import _cuda

    l = psi.shape[0]
    if l % comm.world:
        raise ValueError(f"left bond {l} does not split evenly over {comm.world} ranks")
    lo, hi = row_block(l, comm.world, comm.rank)
    if hi - lo != l // comm.world or lo != comm.rank * (l // comm.world):
        raise ValueError("row blocks must be equal (left bond a multiple of 2 x world)")
    L_rows = L[:, :, lo:hi].contiguous()
    psi_rows = psi[lo:hi].contiguous().clone()
    image_rows = torch.empty_like(psi_rows) if image else None
    stats = _cuda.eig_lowest_rows(comm, L_rows, W, R, psi_rows, l, lo, tol=tol, flags=flags, image_rows=image_rows, **opts)
    return stats, psi_rows, image_rows


def sharded_local_update(comm, L_rows, W, R, psi_rows, neighbour, l: int, tol: float = 1e-8, flags: int = 0,
                         alpha: float = 1e-5, **opts):
    """One *rightward* local update of a row-sharded site (finite_dmrg.py:164-170 for one site, over ``comm``'s GPUs):

      1. eigensolve            ``tnpy_eig_lowest_rows`` on this rank's rows (collective; H_eff psi comes with it)
      2. perturbation          psi += alpha * H_eff psi on the rows (no renormalisation, as in the reference)
      3. split                 all-gather psi (the one vector-sized exchange), then the verified Cholesky-QR split of
                               the full (l d) x r matrix, replicated on every rank (deterministic: same input, same
                               kernels), and the absorb into the replicated neighbour
      4. environment update    every rank contracts *its* bra rows of L with the new site tensor
                               (``tnpy_env_update_left_rows``), the contributions are summed by one all-reduce and
                               the rank keeps its row block of the next left environment.

    ``L_rows`` (l, wl, l_rows) and ``psi_rows`` (l_rows, d, r) are this rank's blocks (rank g holds rows
    [g l / G, (g + 1) l / G)); ``W``, ``R`` (r, wr, r) and ``neighbour`` (r, d, r2) are full.  Returns a dict with the
    solver stats, the new site tensor (full, left-orthonormal), the new neighbour, this rank's rows of the next left
    environment, and the seconds of every phase (device-synchronised)."""
    import time

    import torch

    from tnpy_b200 import _cuda
    from tnpy_b200.matrix_product_state import Direction, _split_on_device

    lo, d, r = psi_rows.shape
    if lo * comm.world != l:
        raise ValueError(f"left bond {l} is not {comm.world} x {lo} rows")
    if r % comm.world:
        raise ValueError(f"right bond {r} does not split evenly over {comm.world} ranks")
    row0 = comm.rank * lo
    phases = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        phases[name] = time.perf_counter() - t0
        return time.perf_counter()

    torch.cuda.synchronize()
    t = time.perf_counter()
    image_rows = torch.empty_like(psi_rows)
    stats = _cuda.eig_lowest_rows(comm, L_rows, W, R, psi_rows, l, row0, tol=tol, flags=flags, image_rows=image_rows, **opts)
    t = tick("eigensolve", t)
    _cuda.axpy(alpha, image_rows, psi_rows)
    t = tick("perturb", t)
    psi_full = torch.empty((l, d, r), dtype=torch.float64, device=psi_rows.device)
    comm.allgather(psi_rows.contiguous(), psi_full)
    site_tensor, new_neighbour, spectrum = _split_on_device(psi_full, neighbour, Direction.RIGHTWARD, "qr")
    site_tensor = site_tensor.contiguous()
    t = tick("gather_and_split", t)
    left_flag = flags & _cuda.LEFT_IDENTITY
    partial = _cuda.env_update_left_rows(L_rows, site_tensor, W, row0, flags=left_flag)
    comm.allreduce_sum(partial)
    ro = r // comm.world
    next_rows = partial[:, :, comm.rank * ro:(comm.rank + 1) * ro].contiguous()
    tick("env_update", t)
    return {"stats": stats, "site_tensor": site_tensor, "neighbour": new_neighbour.contiguous(), "spectrum": spectrum,
            "next_left_rows": next_rows, "phase_s": phases}
