"""NumPy restatement of the algorithm behind ``tnpy_qr_split`` (csrc/qr.cu): the numerical claims DESIGN.md
section 4a makes about it, checked on the CPU -- the blocked right-looking Cholesky with 64-wide blocks and the
recursive-doubling triangular inverse follow the kernel's index arithmetic block by block, so a change of the
blocking logic can be tried here first.  (The CUDA kernels themselves are tested under -m gpu.)"""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import tnpy_oracle as oracle

NB = 64


def padded(n):
    blocks, p = -(-n // NB), 1
    while p < blocks:
        p *= 2
    return p * NB


def blocked_cholesky_inverse(g):
    """g: (np, np) SPD, np = 64 * 2^k.  Returns (L, L^-1) the way cholesky_inverse() in qr.cu builds them."""
    n = g.shape[0]
    g = g.copy()
    nblk = n // NB
    dinv = []
    for kb in range(nblk):
        lo, hi = kb * NB, (kb + 1) * NB
        lkk = np.linalg.cholesky(np.tril(g[lo:hi, lo:hi]) + np.tril(g[lo:hi, lo:hi], -1).T)  # chol_diag_kernel
        g[lo:hi, lo:hi] = lkk
        dinv.append(np.linalg.inv(lkk))
        if hi < n:
            g[hi:, lo:hi] = g[hi:, lo:hi] @ dinv[-1].T                      # panel solve (mm_kernel, in place)
            g[hi:, hi:] -= np.tril(g[hi:, lo:hi] @ g[hi:, lo:hi].T)         # trailing update, lower tiles only
    low = np.tril(g)
    inv = np.zeros_like(low)
    for kb in range(nblk):
        inv[kb * NB:(kb + 1) * NB, kb * NB:(kb + 1) * NB] = dinv[kb]
    s = NB
    while s < n:                                                            # recursive doubling
        for p in range(n // (2 * s)):
            o = 2 * s * p
            t = low[o + s:o + 2 * s, o:o + s] @ inv[o:o + s, o:o + s]
            inv[o + s:o + 2 * s, o:o + s] = -inv[o + s:o + 2 * s, o + s:o + 2 * s] @ t
        s *= 2
    return low, inv


def cholqr(x, shifted):
    """rows of x = the vectors.  Returns (q, defect) or (None, inf) on a non-positive pivot."""
    n = x.shape[0]
    n_pad = padded(n)
    passes = 3 if shifted else 2
    cur = x
    for p in range(passes):
        g = np.eye(n_pad)
        gram = cur @ cur.T
        d = 1.0 / np.sqrt(np.diag(gram)) if p == 0 else np.ones(n)
        g[:n, :n] = gram * d[:, None] * d[None, :]
        if p == 0 and shifted:
            g[:n, :n] += 100 * 1.1102230246251565e-16 * n * np.eye(n)
        try:
            _, inv = blocked_cholesky_inverse(g)
        except np.linalg.LinAlgError:
            return None, np.inf
        cur = (inv[:n, :n] * d[None, :]) @ cur
    return cur, np.abs(cur @ cur.T - np.eye(n)).max()


@pytest.mark.parametrize("n", [64, 100, 256, 320])
def test_blocked_cholesky_and_recursive_inverse(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, 2 * n))
    g = np.eye(padded(n))
    g[:n, :n] = a @ a.T / (2 * n)
    low, inv = blocked_cholesky_inverse(g)
    assert np.abs(low @ low.T - g).max() < 1e-13
    assert np.abs(inv @ low - np.eye(len(g))).max() < 1e-11
    assert np.abs(np.triu(inv, 1)).max() == 0.0


@pytest.mark.parametrize("cond,two_pass_ok,three_pass_ok", [(1e4, True, True), (1e7, True, True), (1e10, False, True),
                                                             (1e13, False, True)])
def test_conditioning_limits(cond, two_pass_ok, three_pass_ok):
    """Two passes reach rounding-level orthogonality up to cond ~1e8 of the *normalised* vectors, the shifted
    three-pass variant up to ~1e14; beyond that the defect says so (and the caller goes to the SVD)."""
    rng = np.random.default_rng(3)
    n, m = 128, 256
    u, _ = np.linalg.qr(rng.standard_normal((m, n)))
    v, _ = np.linalg.qr(rng.standard_normal((n, n)))
    x = (v * np.logspace(0, -np.log10(cond), n)) @ u.T
    for shifted, ok in ((False, two_pass_ok), (True, three_pass_ok)):
        q, defect = cholqr(x, shifted)
        assert (defect <= 1e-13) == ok, (cond, shifted, defect)
        if ok:
            t = x @ q.T
            assert np.abs(t @ q - x).max() < 1e-13
            assert np.abs(np.linalg.svd(t, compute_uv=False) - np.linalg.svd(x, compute_uv=False)).max() < 1e-13


def test_norm_scaling_removes_the_grading():
    """Vectors graded over 12 decades but otherwise well conditioned (a warm DMRG site tensor): the scaling by
    the vector norms is what lets two passes succeed."""
    rng = np.random.default_rng(5)
    n, m = 96, 192
    x = rng.standard_normal((n, m)) * np.logspace(0, -12, n)[:, None]
    q, defect = cholqr(x, shifted=False)
    assert defect < 1e-13
    assert np.linalg.cond(x) > 1e11  # without the scaling this would be far outside two-pass territory


def test_warm_sweep_site_tensors_are_easy_cold_ones_are_not():
    """What DESIGN.md 4a says about DMRG tensors: in the third sweep the normalised vectors of every split are
    conditioned at O(1..100); in the first sweep from a random state some are beyond 1e8."""
    n, chi = 16, 32
    mpo = oracle.xxz_mpo(n, 0.5)
    f = oracle.FiniteDMRG(mpo, chi, mps=oracle.random_mps(n, chi, 2, seed=0))
    conds = {0: [], 2: []}
    orig = oracle.split_tensor
    state = {"sweep": 0}

    def spy(mps, site, direction):
        a = mps[site]
        if a.ndim == 3 and min(a.shape[0], a.shape[2]) >= 16 and state["sweep"] in conds:
            x = a.reshape(a.shape[0] * a.shape[1], -1).T if direction == oracle.RIGHTWARD else a.reshape(a.shape[0], -1)
            conds[state["sweep"]].append(np.linalg.cond(x / np.linalg.norm(x, axis=1)[:, None]))
        return orig(mps, site, direction)

    oracle.split_tensor = spy
    try:
        direction = oracle.RIGHTWARD
        for sweep in range(3):
            state["sweep"] = sweep
            f.sweep(direction, tol=1e-8)
            direction = -direction
    finally:
        oracle.split_tensor = orig
    assert max(conds[2]) < 1e3
    assert max(conds[0]) > 1e4
