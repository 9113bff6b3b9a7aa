"""CPU oracle for the finite-DMRG local-update hot path of tanlin2013/tnpy.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tnpy_b200/`` may import this module; it is
imported by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, always as the checker / CPU arm and never as the thing shipped.

What it is
----------
A plain NumPy/SciPy restatement of the reference's algorithm (pure Python package
``/root/reference/tnpy``, v0.1.1a3).  Every function cites the reference ``file:line`` it follows.
The reference delegates its arithmetic to third-party packages that are neither vendored under
``/root/reference`` nor installable in the build container (no network, not in the wheelhouse):

    quimb 1.4.0 (pyproject.toml:25)         tensor-network contraction + index bookkeeping
    opt-einsum 3.3.0 (poetry.lock:2080)     pairwise contraction path -> numpy.tensordot
    primme 3.2.1 (pyproject.toml:18)        Davidson-type eigensolver behind linalg.eigshmv
    tensornetwork 0.4.6 (pyproject.toml:21) neighbour absorb in split_tensor (a matmul)

so the reference itself cannot be imported here.  Their published behaviour is restated:
contractions as ``numpy.einsum`` / ``tensordot`` (mathematically path-independent), the eigensolver as
SciPy ARPACK ``eigsh`` (or dense ``eigh`` when asked) converged to primme's documented stopping rule
``||A x - theta x|| <= tol * ||A||`` with ``||A||`` estimated by the largest |Ritz value|.

Parity pinning
--------------
Pinned against every golden vector the reference's own tests hold for this path
(``tests/test_oracle_golden.py``): the dense Hamiltonians of tests/test_operators.py:46-88 (exact),
SpinOperators identities :9-16, ``square()`` shapes and H*H :20-28, the RandomHeisenberg seed/offset
rule tests/model/test_random_heisenberg.py:21-26, the compressed bond dimensions
tests/test_matrix_product_state.py:13-37, split_tensor invariance :83-89 and the end-to-end energy
tests/test_finite_dmrg.py:22-23 (fDMRG == dense ED at n=10, chi=32, atol 1e-8).
**Parity unpinned at the kernel level**: no reference test asserts anything about
``one_site_matvec`` / ``update_left`` / ``update_right`` / ``eigshmv`` in isolation
(tests/test_matrix_product_state.py:98-112 are prints and ``pass``); those are pinned only through
the end-to-end energy above plus the invariants checked in the tests (H_eff symmetric, identity
channels in canonical gauge, <psi|H_eff|psi> == <H>).
"""
from __future__ import annotations

import itertools
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as spla
import scipy.sparse.linalg as spsla

RIGHTWARD = 1  # matrix_product_state.py:24
LEFTWARD = -1  # matrix_product_state.py:25


# --------------------------------------------------------------------------------------
# operators.py / model/*.py  -- MPO construction (bit-exact host arithmetic)
# --------------------------------------------------------------------------------------
def spin_operators(spin: float = 0.5):
    """operators.py:55-60 -- Sp, Sm, Sz, I2, O2 for the given spin prefactor."""
    sp = spin * np.array([[0, 2], [0, 0]], dtype=float)
    sm = spin * np.array([[0, 0], [2, 0]], dtype=float)
    sz = spin * np.array([[1, 0], [0, -1]], dtype=float)
    return sp, sm, sz, np.identity(2, dtype=float), np.zeros((2, 2), dtype=float)


def _apply_boundary(full: np.ndarray, site: int, n: int, row: int = 0, col: int = -1) -> np.ndarray:
    """model/utils.py:22-30 -- left end keeps MPO row ``row``, right end keeps column ``col``."""
    if site == 0:
        return full[row, :, :, :]
    if site == n - 1:
        return full[:, col, :, :]
    return full


def _drop_penalty_channel(full: np.ndarray, penalty: float, row: int = 3, col: int = 3) -> np.ndarray:
    """model/utils.py:52-60 -- delete the penalty row/column when penalty == 0."""
    if penalty == 0:
        return np.delete(np.delete(full, row, axis=0), col, axis=1)
    return full


def xxz_mpo(n: int, delta: float) -> List[np.ndarray]:
    """model/xxz.py:19-30 -- w=5 upper-triangular MPO, layout (w_l, w_r, up, down)."""
    sp, sm, sz, i2, o2 = spin_operators()
    out = []
    for site in range(n):
        full = np.array(
            [
                [i2, -0.5 * sp, -0.5 * sm, -delta * sz, o2],
                [o2, o2, o2, o2, sm],
                [o2, o2, o2, o2, sp],
                [o2, o2, o2, o2, sz],
                [o2, o2, o2, o2, i2],
            ]
        )
        out.append(_apply_boundary(full, site, n))
    return out


def thirring_mpo(n: int, delta: float, ma: float, penalty: float, s_target: int) -> List[np.ndarray]:
    """model/thirring.py:40-65 -- w=6 (5 when penalty == 0)."""
    sp, sm, sz, i2, o2 = spin_operators()
    out = []
    for site in range(n):
        beta = delta + ((-1.0) ** site * ma) - 2.0 * penalty * s_target
        gamma = penalty * (0.25 + s_target**2 / n) + 0.25 * delta
        full = np.array(
            [
                [i2, -0.5 * sp, -0.5 * sm, 2.0 * np.sqrt(penalty) * sz, delta * sz, gamma * i2 + beta * sz],
                [o2, o2, o2, o2, o2, sm],
                [o2, o2, o2, o2, o2, sp],
                [o2, o2, o2, i2, o2, np.sqrt(penalty) * sz],
                [o2, o2, o2, o2, o2, sz],
                [o2, o2, o2, o2, o2, i2],
            ],
            dtype=float,
        )
        out.append(_apply_boundary(_drop_penalty_channel(full, penalty), site, n))
    return out


def random_heisenberg_fields(n: int, h: float, seed: Optional[int]) -> np.ndarray:
    """model/random_heisenberg.py:57-58 -- RandomState(seed).uniform(-h, h, size=n)."""
    return np.random.RandomState(seed).uniform(-h, h, size=n)


def random_heisenberg_mpo(
    n: int, h: float, penalty: float = 0, s_target: int = 0, offset: float = 0, seed: Optional[int] = None
) -> List[np.ndarray]:
    """model/random_heisenberg.py:84-109 -- w=6 (5 when penalty == 0), seeded on-site fields."""
    sp, sm, sz, i2, o2 = spin_operators()
    fields = random_heisenberg_fields(n, h, seed)
    out = []
    for site in range(n):
        alpha = penalty * (0.25 + s_target**2 / n) - offset / n
        beta = fields[site] - 2.0 * penalty * s_target
        full = np.array(
            [
                [i2, 0.5 * sp, 0.5 * sm, 2.0 * penalty * sz, sz, alpha * i2 + beta * sz],
                [o2, o2, o2, o2, o2, sm],
                [o2, o2, o2, o2, o2, sp],
                [o2, o2, o2, i2, o2, sz],
                [o2, o2, o2, o2, o2, sz],
                [o2, o2, o2, o2, o2, i2],
            ],
            dtype=float,
        )
        out.append(_apply_boundary(_drop_penalty_channel(full, penalty), site, n))
    return out


def mpo_square(mpo: Sequence[np.ndarray]) -> List[np.ndarray]:
    """operators.py:91-116 -- merge two MPO layers; first layer's ``b`` joins second layer's ``k``.

    Fused bond order is (first-layer bond, second-layer bond), first layer slow
    (operators.py:105-114: fuse [inds[0], inds[2]] / [inds[0], inds[3]], [inds[1], inds[4]]).
    """
    n = len(mpo)
    out = []
    for site, w in enumerate(mpo):
        if site == 0 or site == n - 1:
            t = np.einsum("apx,bxq->abpq", w, w)
            out.append(t.reshape(w.shape[0] ** 2, w.shape[1], w.shape[2]))
        else:
            t = np.einsum("acpx,bdxq->abcdpq", w, w)
            out.append(t.reshape(w.shape[0] ** 2, w.shape[1] ** 2, w.shape[2], w.shape[3]))
    return out


def full_hamiltonian(mpo: Sequence[np.ndarray]) -> np.ndarray:
    """operators.py:119-156 -- dense H[(k0..kn-1),(b0..bn-1)] from the MPO (n <= 12 guard :140)."""
    n = len(mpo)
    d = mpo[0].shape[-1]
    if d**n > 2**12:
        raise ResourceWarning(f"Requesting more than {n} sites with physical dim {d}.")
    acc = mpo[0]  # (w, k0, b0)
    acc = np.transpose(acc, (1, 2, 0))  # (K, B, w)
    for site in range(1, n):
        w = mpo[site]
        if site == n - 1:
            acc = np.einsum("KBa,apq->KpBq", acc, w)
            acc = acc.reshape(acc.shape[0] * acc.shape[1], acc.shape[2] * acc.shape[3])
        else:
            acc = np.einsum("KBa,abpq->KpBqb", acc, w)
            acc = acc.reshape(acc.shape[0] * acc.shape[1], acc.shape[2] * acc.shape[3], acc.shape[4])
    return acc


# --------------------------------------------------------------------------------------
# matrix_product_state.py -- MPS helpers
# --------------------------------------------------------------------------------------
def compressed_bond_dims(n: int, bond_dim: int, phys_dim: int) -> List[int]:
    """Bond dims of ``MatrixProductState.random`` after quimb's compress()
    (matrix_product_state.py:170-185; pinned by tests/test_matrix_product_state.py:13-37)."""
    return [int(min(phys_dim ** (i + 1), bond_dim, phys_dim ** (n - 1 - i))) for i in range(n - 1)]


def random_mps(n: int, bond_dim: int, phys_dim: int = 2, seed: int = 0) -> List[np.ndarray]:
    """Harness-side initial state: default_rng(seed) normals at the compressed bond
    dims, right-canonicalised by QR from site n-1 down to 1, site 0 normalised.  Layout 'lpr':
    site 0 (d, r), bulk (l, d, r), last (l, d)  (matrix_product_state.py:40-45).

    Stands in for quimb's ``MPS_rand_state(...).compress()`` (matrix_product_state.py:183-184),
    whose RNG stream is third-party and not reproducible here; the state class is the same
    (right-canonical, unit norm, same shapes).
    """
    rng = np.random.default_rng(seed)
    chi = compressed_bond_dims(n, bond_dim, phys_dim)
    dims = [1] + chi + [1]
    arrays = [rng.standard_normal((dims[i], phys_dim, dims[i + 1])) for i in range(n)]
    for site in range(n - 1, 0, -1):
        l, d, r = arrays[site].shape
        q, rr = np.linalg.qr(arrays[site].reshape(l, d * r).T)  # (d r, l) = q (d r, l) rr (l, l)
        arrays[site] = q.T.reshape(l, d, r)
        arrays[site - 1] = np.einsum("lpr,sr->lps", arrays[site - 1], rr)
    arrays[0] /= np.linalg.norm(arrays[0])
    arrays[0] = arrays[0].reshape(phys_dim, dims[1])
    arrays[-1] = arrays[-1].reshape(dims[n - 1], phys_dim)
    return arrays


def _as3(a: np.ndarray, site: int, n: int) -> np.ndarray:
    """View an edge site tensor with an explicit unit bond: always (l, d, r)."""
    if a.ndim == 3:
        return a
    return a[None, :, :] if site == 0 else a[:, :, None]


def _w4(w: np.ndarray, site: int, n: int) -> np.ndarray:
    """View an edge MPO tensor with an explicit unit bond: always (w_l, w_r, d, d)."""
    if w.ndim == 4:
        return w
    return w[None, :, :, :] if site == 0 else w[:, None, :, :]


def mps_to_dense(mps: Sequence[np.ndarray]) -> np.ndarray:
    n = len(mps)
    acc = _as3(mps[0], 0, n)[0]  # (d, r)
    for site in range(1, n):
        acc = np.tensordot(acc, _as3(mps[site], site, n), axes=(acc.ndim - 1, 0))
        acc = acc.reshape(-1, acc.shape[-1])
    return acc.reshape(-1)


def mps_overlap(a: Sequence[np.ndarray], b: Sequence[np.ndarray]) -> float:
    """<a|b> by transfer matrices (real tensors)."""
    n = len(a)
    e = np.ones((1, 1))
    for site in range(n):
        x, y = _as3(a[site], site, n), _as3(b[site], site, n)
        e = np.einsum("lm,lpr,mps->rs", e, x, y)
    return float(e[0, 0])


def mps_expectation(mps: Sequence[np.ndarray], mpo: Sequence[np.ndarray]) -> float:
    """<psi|O|psi> (not divided by the norm) -- MatrixProductStateMeasurements.expectation_value,
    matrix_product_state.py:447-453."""
    n = len(mps)
    e = np.ones((1, 1, 1))
    for site in range(n):
        a, w = _as3(mps[site], site, n), _w4(mpo[site], site, n)
        e = np.einsum("lam,lpr,abpq,mqs->rbs", e, a, w, a, optimize=True)
    return float(e[0, 0, 0])


def mps_sz_profile(mps: Sequence[np.ndarray]) -> Tuple[np.ndarray, float]:
    """(<Sz_i> / <psi|psi> for every site, <psi|psi>) of an 'lpr' spin-1/2 MPS: left / right overlap environments
    (the measurement matrix_product_state.py:447-453 makes through quimb, specialised to one-site operators)."""
    n = len(mps)
    a3 = [_as3(a, i, n) for i, a in enumerate(mps)]
    left = [np.ones((1, 1))]
    for a in a3:
        left.append(np.einsum("lm,lpr,mps->rs", left[-1], a, a, optimize=True))
    right = [np.ones((1, 1))] * (n + 1)
    for i in range(n - 1, -1, -1):
        right[i] = np.einsum("rs,lpr,mps->lm", right[i + 1], a3[i], a3[i], optimize=True)
    sz = np.diag([0.5, -0.5])
    norm2 = float(left[-1][0, 0])
    prof = [float(np.einsum("lm,lpr,pq,mqs,rs->", left[i], a3[i], sz, a3[i], right[i + 1], optimize=True)) / norm2
            for i in range(n)]
    return np.array(prof), norm2


def split_tensor(mps: List[np.ndarray], site: int, direction: int) -> np.ndarray:
    """matrix_product_state.py:187-225 + linalg.py:9-23 -- thin SVD with cutoff = current bond
    (never truncates), A[site] <- U or Vt, neighbour absorbs diag(s) Vt / U diag(s).
    Returns the singular values of the bond."""
    n = len(mps)
    a = mps[site]
    d = a.shape[0] if site == 0 else a.shape[1]
    if direction == RIGHTWARD:
        psi = a if site == 0 else a.reshape(d * a.shape[0], -1)
        cutoff = a.shape[-1]
        u, s, vt = np.linalg.svd(psi, full_matrices=False)
        u, s, vt = u[:, :cutoff], s[:cutoff], vt[:cutoff, :]
        mps[site] = u.reshape(a.shape)
        residual = np.diagflat(s) @ vt
        mps[site + 1] = np.tensordot(residual, mps[site + 1], axes=(1, 0))
    elif direction == LEFTWARD:
        psi = a if site == n - 1 else a.reshape(-1, d * a.shape[2])
        cutoff = a.shape[0]
        u, s, vt = np.linalg.svd(psi, full_matrices=False)
        u, s, vt = u[:, :cutoff], s[:cutoff], vt[:cutoff, :]
        mps[site] = vt.reshape(a.shape)
        residual = u @ np.diagflat(s)
        mps[site - 1] = np.tensordot(mps[site - 1], residual, axes=(mps[site - 1].ndim - 1, 0))
    else:
        raise KeyError("MatrixProductState only supplies left or right direction.")
    return s


# --------------------------------------------------------------------------------------
# matrix_product_state.py:231-440 -- Environment
# --------------------------------------------------------------------------------------
def heff_apply(L: Optional[np.ndarray], W: np.ndarray, R: Optional[np.ndarray], x: np.ndarray) -> np.ndarray:
    """The H_eff.psi contraction, matrix_product_state.py:423-438.

    Bulk: y[m,q,s] = sum L[l,a,m] W[a,b,p,q] R[r,b,s] x[l,p,r]; L = (ket, mpo, bra), W = 'lrud',
    R = (ket, mpo, bra).  Edges: site 0 has no L and W is (w_r,d,d), x (d,r) (:425-427); the last
    site has no R, W is (w_l,d,d), x (l,d) (:428-430).  Lowered the way opt_einsum would: two big
    tensordots around the small W mixing.
    """
    if L is None and R is None:
        raise ValueError("one-site chain not supported")
    if L is None:
        t = np.tensordot(x, R, axes=(1, 0))  # (p, b, s)
        return np.einsum("bpq,pbs->qs", W, t)
    if R is None:
        t = np.tensordot(L, x, axes=(0, 0))  # (a, m, p)
        return np.einsum("apq,amp->mq", W, t)
    # three pairwise tensordots (transpose-copy + dgemm each), the lowering opt_einsum gives quimb
    t1 = np.tensordot(L, x, axes=(0, 0))  # (a, m, p, r)
    t2 = np.tensordot(W, t1, axes=((0, 2), (0, 2)))  # (b, q, m, r)
    y = np.tensordot(t2, R, axes=((0, 3), (1, 0)))  # (q, m, s)
    return np.ascontiguousarray(np.transpose(y, (1, 0, 2)))


def env_update_left(L: Optional[np.ndarray], A: np.ndarray, W: np.ndarray) -> np.ndarray:
    """matrix_product_state.py:296-315: L'[r,b,s] = sum L[l,a,m] A[l,p,r] W[a,b,p,q] A[m,q,s]
    (site == 1 drops L: A is (d,r), W is (w_r,d,d))."""
    if L is None:
        return np.einsum("pr,bpq,qs->rbs", A, W, A, optimize=True)
    t1 = np.tensordot(L, A, axes=(0, 0))  # (a, m, p, r)
    t2 = np.einsum("abpq,ampr->mqrb", W, t1, optimize=True)
    return np.tensordot(t2, A, axes=((0, 1), (0, 1)))  # (r, b, s)


def env_update_right(R: Optional[np.ndarray], A: np.ndarray, W: np.ndarray) -> np.ndarray:
    """matrix_product_state.py:317-336: R'[l,a,m] = sum R[r,b,s] A[l,p,r] W[a,b,p,q] A[m,q,s]
    (site == n-2 drops R: A is (l,d), W is (w_l,d,d))."""
    if R is None:
        return np.einsum("lp,apq,mq->lam", A, W, A, optimize=True)
    t1 = np.tensordot(A, R, axes=(2, 0))  # (l, p, b, s)
    t2 = np.einsum("abpq,lpbs->laqs", W, t1, optimize=True)
    return np.tensordot(t2, A, axes=((2, 3), (1, 2)))  # (l, a, m)


class Environment:
    """matrix_product_state.py:231-440.  ``left[site]`` for site in 1..n-1, ``right[site]`` for
    site in 0..n-2, both stacks built up front (:247-250).  Real data: the bra copy is the ket."""

    def __init__(self, mpo: Sequence[np.ndarray], mps: Sequence[np.ndarray]):
        self.mpo = [np.asarray(w, dtype=float) for w in mpo]
        self.mps = [np.array(a, dtype=float) for a in mps]
        self.n_sites = len(self.mpo)
        self.left: Dict[int, np.ndarray] = {}
        self.right: Dict[int, np.ndarray] = {}
        for site in range(1, self.n_sites):
            self.update_left(site)
        for site in range(self.n_sites - 2, -1, -1):
            self.update_right(site)

    def update_left(self, site: int):
        prev = None if site == 1 else self.left[site - 1]
        self.left[site] = env_update_left(prev, self.mps[site - 1], self.mpo[site - 1])

    def update_right(self, site: int):
        prev = None if site == self.n_sites - 2 else self.right[site + 1]
        self.right[site] = env_update_right(prev, self.mps[site + 1], self.mpo[site + 1])

    def update(self, site: int, direction: int):
        """:338-353"""
        if direction == RIGHTWARD:
            self.update_left(site + 1)
        elif direction == LEFTWARD:
            self.update_right(site - 1)

    def update_mps(self, site: int, data: np.ndarray):
        """:355-357"""
        self.mps[site] = np.array(data, dtype=float)

    def split_tensor(self, site: int, direction: int) -> np.ndarray:
        """:359-365"""
        return split_tensor(self.mps, site, direction)

    def _lr(self, site: int):
        L = None if site == 0 else self.left[site]
        R = None if site == self.n_sites - 1 else self.right[site]
        return L, R

    def matvec(self, site: int, x: np.ndarray) -> np.ndarray:
        """:423-438 -- accepts (N,) or (N,1), returns (N,1)."""
        L, R = self._lr(site)
        shape = self.mps[site].shape
        return heff_apply(L, self.mpo[site], R, np.asarray(x).reshape(shape)).reshape(-1, 1)

    def one_site_matvec(self, site: int) -> spsla.LinearOperator:
        """:411-440"""
        size = self.mps[site].size
        return spsla.LinearOperator(shape=(size, size), matvec=lambda x: self.matvec(site, x), dtype=float)

    def one_site_full_matrix(self, site: int) -> np.ndarray:
        """:372-409 -- dense H_eff[(l,p,r),(m,q,s)] ('k' fuse rows, 'b' fuse columns)."""
        L, R = self._lr(site)
        W = self.mpo[site]
        if site == 0:
            t = np.einsum("bpq,rbs->prqs", W, R, optimize=True)
        elif site == self.n_sites - 1:
            t = np.einsum("lam,apq->lpmq", L, W, optimize=True)
        else:
            t = np.einsum("lam,abpq,rbs->lprmqs", L, W, R, optimize=True)
        size = self.mps[site].size
        return t.reshape(size, size)

    def variance(self) -> float:
        """:367-370 -- <H^2> - <H>^2 on the (un-normalised) state."""
        return mps_expectation(self.mps, mpo_square(self.mpo)) - mps_expectation(self.mps, self.mpo) ** 2


# --------------------------------------------------------------------------------------
# linalg.py -- eigensolvers
# --------------------------------------------------------------------------------------
def eigh_lowest(matrix: np.ndarray) -> Tuple[float, np.ndarray]:
    """linalg.py:42-61 with k=1, backend 'numpy': (evals[0], evecs[:, 0])."""
    evals, evecs = np.linalg.eigh(matrix)
    return float(evals[0]), evecs[:, 0]


def eigshmv(
    matvec: Callable[[np.ndarray], np.ndarray],
    v0: np.ndarray,
    tol: float = 0.0,
    dense_below: int = 0,
    maxiter: Optional[int] = None,
) -> Tuple[float, np.ndarray]:
    """linalg.py:64-87 -- lowest eigenpair ('SA', k=1), returns (eval, evec of shape (N,1)).

    primme is not available; the stopping rule it documents is reproduced instead:
    ||A x - theta x|| <= tol * ||A||, ||A|| ~ largest |Ritz value| (tol = 0 means
    machine-epsilon * 1e4 as in primme's default).  ARPACK's own ``tol`` is a relative Ritz-value
    accuracy, so it is tightened until the residual rule holds.  ``dense_below`` > N switches to
    the exact dense answer (used by tests as the solver-independent oracle).
    """
    v0 = np.asarray(v0, dtype=float).reshape(-1)
    size = v0.size
    if tol == 0:
        tol = np.finfo(float).eps * 1e4
    if size <= max(dense_below, 3):
        dense = np.column_stack([np.asarray(matvec(e)).reshape(-1) for e in np.eye(size)])
        dense = 0.5 * (dense + dense.T)
        e, v = eigh_lowest(dense)
        return e, v.reshape(-1, 1)
    op = spsla.LinearOperator(shape=(size, size), matvec=lambda x: np.asarray(matvec(x)).reshape(-1), dtype=float)
    arpack_tol = min(tol, 1e-6) ** 2 * 1e2
    ncv = min(size - 1, 24)
    for _ in range(4):
        try:
            evals, evecs = spsla.eigsh(op, k=1, which="SA", v0=v0, tol=arpack_tol, ncv=ncv, maxiter=maxiter)
        except spsla.ArpackNoConvergence as exc:  # pragma: no cover - pathological
            evals, evecs = exc.eigenvalues, exc.eigenvectors
            if len(evals) == 0:
                raise
        x = evecs[:, 0]
        theta = float(evals[0])
        ax = np.asarray(matvec(x)).reshape(-1)
        resid = np.linalg.norm(ax - theta * x)
        anorm = max(abs(theta), np.linalg.norm(ax))
        if resid <= tol * anorm:
            break
        arpack_tol = max(arpack_tol * 1e-3, 1e-30)
        v0 = x
    return theta, x.reshape(-1, 1)


# --------------------------------------------------------------------------------------
# finite_dmrg.py -- the driver
# --------------------------------------------------------------------------------------
class FiniteDMRG:
    """finite_dmrg.py:32-263 restated on NumPy arrays.

    ``mps`` must be right-canonical (the reference's random() is, via quimb compress()).
    """

    def __init__(
        self,
        mpo: Sequence[np.ndarray],
        bond_dim: int,
        mps: Optional[Sequence[np.ndarray]] = None,
        exact_solver_dim: int = 200,
        seed: int = 0,
        exact_local_solver: bool = False,
    ):
        self.n_sites = len(mpo)
        self.bond_dim = bond_dim
        self.phys_dim = mpo[0].shape[-1]
        self.exact_solver_dim = exact_solver_dim
        self.exact_local_solver = exact_local_solver
        if mps is None:
            mps = random_mps(self.n_sites, bond_dim, self.phys_dim, seed=seed)
        self.env = Environment(mpo, mps)
        self.energies: List[float] = [np.nan]  # :75
        self.variances: List[float] = [np.nan]  # :76
        self.bond_singular_values: Dict[int, np.ndarray] = {}
        self.n_matvec = 0

    @property
    def mps(self) -> List[np.ndarray]:
        return self.env.mps

    def one_site_solver(self, site: int, tol: float = 1e-8) -> Tuple[float, np.ndarray]:
        """:97-111"""
        v0 = self.mps[site].reshape(-1, 1)
        if v0.size < self.exact_solver_dim:
            return eigh_lowest(self.env.one_site_full_matrix(site))

        def mv(x):
            self.n_matvec += 1
            return self.env.matvec(site, x)

        dense_below = 4096 if self.exact_local_solver else 0
        return eigshmv(mv, v0, tol=tol, dense_below=dense_below)

    def perturb_wave_function(self, site: int, alpha: float = 1e-5):
        """:116-141 -- psi += alpha * H_eff psi, no renormalisation."""
        psi = self.mps[site].flatten()
        psi += alpha * self.env.matvec(site, psi).reshape(-1)
        self.env.update_mps(site, psi.reshape(self.mps[site].shape))

    def sweep(self, direction: int = RIGHTWARD, tol: float = 1e-8) -> float:
        """:143-171"""
        sites = range(self.n_sites - 1) if direction == RIGHTWARD else range(self.n_sites - 1, 0, -1)
        energy = None
        for site in sites:
            energy, psi = self.one_site_solver(site, tol)
            self.env.update_mps(site, np.asarray(psi).reshape(self.mps[site].shape))
            self.perturb_wave_function(site)
            s = self.env.split_tensor(site, direction)
            bond = site if direction == RIGHTWARD else site - 1
            self.bond_singular_values[bond] = s
            self.env.update(site, direction)
        return float(energy)

    def run(self, tol: float = 1e-8, max_sweep: int = 100, metric: str = "ENERGY", with_variance: bool = True) -> List[float]:
        """:214-257 -- alternate sweeps until |dE| < tol (first diff is against nan => >= 2 sweeps)."""
        directions = itertools.cycle([RIGHTWARD, LEFTWARD])
        for n_sweep, direction in zip(range(1, max_sweep + 1), directions):
            energy = self.sweep(direction, tol=tol)
            self.energies.append(energy)
            self.variances.append(self.env.variance() if with_variance else np.nan)
            series = self.variances if metric == "VARIANCE" else self.energies
            gradient = np.diff(series[-2:])[0]
            if abs(gradient) < tol:
                break
        return self.energies[1:]


def exact_ground_energy(mpo: Sequence[np.ndarray]) -> float:
    """exact_diagonalization.py:57-58 -- dense eigh of the full Hamiltonian."""
    return float(np.linalg.eigvalsh(full_hamiltonian(mpo))[0])


# --------------------------------------------------------------------------------------
# finite_dmrg.py:266-407 -- ShiftInvertDMRG
# --------------------------------------------------------------------------------------
class ShiftInvertDMRG(FiniteDMRG):
    """Restatement of the reference's shift-invert driver.  The generalised local problem
    A x = lambda M x (A from the MPO of H - eps, M from its square) is solved densely with
    scipy.linalg.eigh(a, b) at every site (the reference does that below ``exact_solver_dim`` and
    hands larger sites to primme.eigsh(A, M=M); dense is the solver-independent answer)."""

    def __init__(self, mpo, bond_dim, offset=0.0, mps=None, seed=0):
        super().__init__(mpo, bond_dim, mps=mps, seed=seed)
        self.env2 = Environment(mpo_square(mpo), self.env.mps)
        self.env2.mps = self.env.mps  # one shared list of site tensors (finite_dmrg.py:379-385 keeps them in step)
        self.offset = offset
        self.restored_mps = None

    def one_site_solver(self, site, tol=1e-8):
        a = self.env.one_site_full_matrix(site)
        b = self.env2.one_site_full_matrix(site)
        a, b = 0.5 * (a + a.T), 0.5 * (b + b.T)
        evals, evecs = spla.eigh(a, b, subset_by_index=[0, 0])
        return float(evals[0]), evecs[:, 0]  # b-normalised: x^T b x = 1, as primme / scipy return it

    def sweep(self, direction=RIGHTWARD, tol=1e-8):
        sites = range(self.n_sites - 1) if direction == RIGHTWARD else range(self.n_sites - 1, 0, -1)
        energy = None
        for site in sites:
            energy, psi = self.one_site_solver(site, tol)
            self.env.update_mps(site, np.asarray(psi).reshape(self.mps[site].shape))
            self.perturb_wave_function(site)
            self.env.split_tensor(site, direction)
            self.env.update(site, direction)
            self.env2.update(site, direction)
        return float(energy)

    def restore_mps(self):
        """finite_dmrg.py:313-339: |psi> = (H - eps)|phi>, fused bonds (MPS bond slow, MPO bond fast)."""
        n = self.n_sites
        arrays = []
        for site in range(n):
            a, w = _as3(self.mps[site], site, n), _w4(self.env.mpo[site], site, n)
            t = np.einsum("lbr,xykb->lxkry", a, w)
            l, wl, k, r, wr = t.shape
            arrays.append(t.reshape(l * wl, k, r * wr))
        arrays[0] = arrays[0][0]
        arrays[-1] = arrays[-1][:, :, 0]
        self.restored_mps = arrays
        return arrays

    def run(self, tol=1e-7, max_sweep=100):
        energies = super().run(tol=tol, max_sweep=max_sweep, with_variance=False)
        self.restore_mps()
        return (np.reciprocal(energies) + self.offset).tolist()
