"""Parity of every CUDA kernel of the path against the CPU oracle, through the C ABI (-m gpu).

Tolerances: all arithmetic is float64; contraction results are compared at 1e-12 relative to the
largest entry (summation-order differences only), eigenvalues at 1e-10 relative (north_star),
singular values at 1e-9 (north_star).
"""
import numpy as np
import pytest

from oracle import tnpy_oracle as oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


def rel_err(got, want):
    scale = max(np.abs(want).max(), 1e-300)
    return np.abs(got - want).max() / scale


@pytest.fixture(scope="module")
def cu():
    from tnpy_b200 import _cuda

    _cuda.load()
    return _cuda


# ------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    (1, 1, 1), (3, 5, 7), (64, 64, 16), (120, 300, 60), (128, 128, 128), (130, 254, 100),
    (256, 640, 128), (512, 384, 200), (1024, 1280, 256), (96, 2048, 48),
]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
@pytest.mark.parametrize("algo", ["generic", "dmma", "auto"])
def test_gemm_tn(cu, m, n, k, algo):
    rng = np.random.default_rng(m * 1000003 + n * 1009 + k)
    a, b = rng.standard_normal((k, m)), rng.standard_normal((k, n))
    code = {"generic": cu.GEMM_GENERIC, "dmma": cu.GEMM_DMMA, "auto": cu.GEMM_AUTO}[algo]
    if algo == "dmma" and (m % 2 or n % 2):
        with pytest.raises(RuntimeError):
            cu.gemm_tn(dev(a), dev(b), algo=code)
        return
    got = cu.gemm_tn(dev(a), dev(b), algo=code).cpu().numpy()
    want = a.T @ b
    assert rel_err(got, want) < 1e-13 * max(k, 16)


@pytest.mark.parametrize("tile", [0, 1, 2])
def test_gemm_tn_every_tile_config(cu, tile):
    rng = np.random.default_rng(tile)
    m, n, k = 384, 448, 272
    a, b = rng.standard_normal((k, m)), rng.standard_normal((k, n))
    cu.load().tnpy_set_gemm_tile(tile)
    try:
        got = cu.gemm_tn(dev(a), dev(b), algo=cu.GEMM_DMMA).cpu().numpy()
        c0 = rng.standard_normal((m, n))
        acc = cu.gemm_tn(dev(a), dev(b), out=dev(c0), accumulate=True, algo=cu.GEMM_DMMA).cpu().numpy()
    finally:
        cu.load().tnpy_set_gemm_tile(-1)
    assert rel_err(got, a.T @ b) < 1e-12
    assert rel_err(acc, c0 + a.T @ b) < 1e-12


def test_gemm_linearity_large(cu):
    """Size-independent property at a bench-like shape: (A1 + A2)^T B == A1^T B + A2^T B."""
    g = torch.Generator(device="cuda").manual_seed(0)
    k, m, n = 1024, 2048, 2560
    a1 = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    a2 = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    lhs = cu.gemm_tn(a1 + a2, b)
    rhs = cu.gemm_tn(a1, b) + cu.gemm_tn(a2, b)
    assert float((lhs - rhs).abs().max() / lhs.abs().max()) < 1e-12


# ------------------------------------------------------------------------- contraction chains
def random_operands(rng, l, r, wl, wr, d, sparse_w=True):
    L = rng.standard_normal((l, wl, l))
    R = rng.standard_normal((r, wr, r))
    W = rng.standard_normal((wl, wr, d, d))
    if sparse_w:
        W *= rng.random((wl, wr, d, d)) < 0.35
    x = rng.standard_normal((l, d, r))
    return L, W, R, x


CHAIN_DIMS = [
    (4, 8, 5, 5, 2), (16, 16, 5, 5, 2), (60, 60, 5, 5, 2), (32, 64, 6, 6, 2), (64, 32, 5, 6, 2),
    (128, 128, 5, 5, 2), (256, 256, 6, 6, 2), (7, 9, 3, 4, 3), (33, 17, 5, 5, 2), (96, 96, 25, 25, 2),
]


@pytest.mark.parametrize("l,r,wl,wr,d", CHAIN_DIMS)
def test_heff_apply_bulk(cu, l, r, wl, wr, d):
    rng = np.random.default_rng(l * 7919 + r)
    L, W, R, x = random_operands(rng, l, r, wl, wr, d)
    want = oracle.heff_apply(L, W, R, x)
    got = cu.heff_apply(dev(L), dev(W), dev(R), dev(x)).cpu().numpy()
    assert rel_err(got, want) < 1e-12


@pytest.mark.parametrize("chi,w,d", [(2, 5, 2), (16, 5, 2), (64, 6, 2), (5, 3, 3)])
def test_heff_apply_edges(cu, chi, w, d):
    rng = np.random.default_rng(chi)
    # site 0: no L, W (w_r, d, d), x (d, r)
    R = rng.standard_normal((chi, w, chi))
    W0 = rng.standard_normal((w, d, d))
    x0 = rng.standard_normal((d, chi))
    want = oracle.heff_apply(None, W0, R, x0)
    got = cu.heff_apply(None, dev(W0[None]), dev(R), dev(x0[None])).cpu().numpy()[0]
    assert rel_err(got, want) < 1e-12
    # last site: no R, W (w_l, d, d), x (l, d)
    L = rng.standard_normal((chi, w, chi))
    x1 = rng.standard_normal((chi, d))
    want = oracle.heff_apply(L, W0, None, x1)
    got = cu.heff_apply(dev(L), dev(W0[:, None]), None, dev(x1[:, :, None])).cpu().numpy()[:, :, 0]
    assert rel_err(got, want) < 1e-12


@pytest.mark.parametrize("l,r,wl,wr,d", CHAIN_DIMS)
def test_env_updates(cu, l, r, wl, wr, d):
    rng = np.random.default_rng(l * 31 + r * 17 + wl)
    L, W, R, A = random_operands(rng, l, r, wl, wr, d)
    got = cu.env_update_left(dev(L), dev(A), dev(W)).cpu().numpy()
    assert rel_err(got, oracle.env_update_left(L, A, W)) < 1e-12
    got = cu.env_update_right(dev(R), dev(A), dev(W)).cpu().numpy()
    assert rel_err(got, oracle.env_update_right(R, A, W)) < 1e-12


def test_env_updates_edges(cu):
    rng = np.random.default_rng(5)
    d, chi, w = 2, 8, 5
    A0, W0 = rng.standard_normal((d, chi)), rng.standard_normal((w, d, d))
    got = cu.env_update_left(None, dev(A0[None]), dev(W0[None])).cpu().numpy()
    assert rel_err(got, oracle.env_update_left(None, A0, W0)) < 1e-12
    An = rng.standard_normal((chi, d))
    got = cu.env_update_right(None, dev(An[:, :, None]), dev(W0[:, None])).cpu().numpy()
    assert rel_err(got, oracle.env_update_right(None, An, W0)) < 1e-12


@pytest.mark.parametrize("l,r,wl,wr,d", [(2, 4, 5, 5, 2), (4, 8, 5, 6, 2), (1, 2, 1, 5, 2), (3, 3, 4, 4, 3),
                                          (16, 12, 6, 5, 2), (8, 9, 25, 25, 2), (1, 40, 1, 5, 2), (33, 1, 5, 1, 2)])
def test_heff_dense(cu, l, r, wl, wr, d):
    rng = np.random.default_rng(l + 10 * r)
    L, W, R, _ = random_operands(rng, l, r, wl, wr, d)
    want = np.einsum("lam,abpq,rbs->lprmqs", L, W, R, optimize=True).reshape(l * d * r, l * d * r)
    got = cu.heff_dense(dev(L), dev(W), dev(R), l, r).cpu().numpy()
    assert rel_err(got, want) < 1e-13


def test_mirror(cu):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 3, 50))
    assert np.array_equal(cu.mirror_lpr(dev(a)).cpu().numpy(), np.transpose(a, (2, 1, 0)))


# -------------------------------------------------------------------------------- vector kernels
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 7200, 131072, 1 << 20, (1 << 20) + 3])
def test_blas1(cu, n):
    rng = np.random.default_rng(n)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    assert abs(cu.dot(dev(x), dev(y)).item() - x @ y) <= 1e-13 * (np.abs(x) @ np.abs(y))
    assert abs(cu.nrm2(dev(x)).item() - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
    assert np.allclose(cu.axpy(0.37, dev(x), dev(y)).cpu().numpy(), y + 0.37 * x, rtol=1e-15, atol=1e-15)
    assert np.allclose(cu.scal(-1.5, dev(x)).cpu().numpy(), -1.5 * x, rtol=1e-15)


@pytest.mark.parametrize("n,m", [(7200, 1), (7200, 3), (131072, 9), (100001, 20), (1 << 20, 5)])
def test_multi_dot_axpy(cu, n, m):
    rng = np.random.default_rng(n + m)
    ld = n + (n & 1)
    V = np.zeros((m, ld))
    V[:, :n] = rng.standard_normal((m, n))
    w = rng.standard_normal(n)
    Vd, wd = dev(V), dev(w)
    h = cu.multi_dot(Vd, wd).cpu().numpy()
    want = V[:, :n] @ w
    assert np.abs(h - want).max() <= 1e-13 * (np.abs(V[:, :n]) @ np.abs(w)).max()
    out = cu.multi_axpy(Vd, dev(want), wd).cpu().numpy()
    assert np.allclose(out, w - want @ V[:, :n], rtol=1e-12, atol=1e-12)


def test_reductions_are_deterministic(cu):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(1 << 22, generator=g, dtype=torch.float64, device="cuda")
    vals = {cu.nrm2(x).item() for _ in range(5)}
    assert len(vals) == 1


def test_two_streams_keep_their_own_reduction_scratch(cu):
    """The two-stage reductions keep partial sums and a ticket counter per (device, stream): kernels enqueued on two
    streams at once give bit-identical results to the same calls made one after the other (one global
    scratch let concurrent streams corrupt each other's dot products)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    n, m = 1 << 22, 8
    V1 = torch.randn((m, n), generator=g, dtype=torch.float64, device="cuda")
    V2 = torch.randn((m, n), generator=g, dtype=torch.float64, device="cuda")
    w1 = torch.randn(n, generator=g, dtype=torch.float64, device="cuda")
    w2 = torch.randn(n, generator=g, dtype=torch.float64, device="cuda")
    want1, want2 = cu.multi_dot(V1, w1).clone(), cu.multi_dot(V2, w2).clone()
    nrm1, nrm2 = cu.nrm2(w1).clone(), cu.nrm2(w2).clone()
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(10):
        with torch.cuda.stream(s1):
            a, na = cu.multi_dot(V1, w1), cu.nrm2(w1)
        with torch.cuda.stream(s2):
            b, nb = cu.multi_dot(V2, w2), cu.nrm2(w2)
        torch.cuda.synchronize()
        assert torch.equal(a, want1) and torch.equal(b, want2) and torch.equal(na, nrm1) and torch.equal(nb, nrm2)


# ------------------------------------------------------------------------------------ eigensolver
def canonical_problem(n_sites, chi, site, seed=0, model="xxz"):
    # penalty 1.0: with the script's 100.0 a *random* environment gives a relative gap of ~5e-7
    # (||A|| = 758, gap 4e-4), out of reach of any unpreconditioned Krylov solver in a unit test
    mpo = oracle.xxz_mpo(n_sites, 0.5) if model == "xxz" else oracle.thirring_mpo(n_sites, 0.5, 1.0, 1.0, 0)
    mps = oracle.random_mps(n_sites, chi, 2, seed=seed)
    env = oracle.Environment(mpo, mps)
    return env, mpo, mps


@pytest.mark.parametrize("n_sites,chi,site,model", [(10, 16, 4, "xxz"), (12, 24, 6, "thirring"), (10, 32, 5, "xxz")])
def test_eig_lowest_matches_dense(cu, n_sites, chi, site, model):
    env, mpo, mps = canonical_problem(n_sites, chi, site, model=model)
    H = env.one_site_full_matrix(site)
    evals = np.linalg.eigvalsh(0.5 * (H + H.T))
    psi = dev(mps[site])
    W = dev(mpo[site])
    stats = cu.eig_lowest(dev(env.left[site]), W, dev(env.right[site]), psi, tol=1e-10)
    assert stats["converged"]
    assert abs(stats["theta"] - evals[0]) <= 1e-10 * max(abs(evals).max(), 1.0)
    x = psi.cpu().numpy().reshape(-1)
    assert abs(np.linalg.norm(x) - 1.0) < 1e-12
    resid = np.linalg.norm(H.T @ x - stats["theta"] * x)
    assert resid <= 1e-9 * abs(evals).max()


def test_eig_lowest_restarts(cu):
    """A tiny basis forces thick restarts; the answer must not change."""
    env, mpo, mps = canonical_problem(12, 32, 6, seed=3)
    site = 6
    H = env.one_site_full_matrix(site)
    e0 = np.linalg.eigvalsh(0.5 * (H + H.T))[0]
    psi = dev(np.random.default_rng(0).standard_normal(mps[site].shape))
    stats = cu.eig_lowest(dev(env.left[site]), dev(mpo[site]), dev(env.right[site]), psi, tol=1e-10, ncv=6,
                          max_matvec=2000)
    assert stats["converged"] and stats["n_restart"] > 0
    assert abs(stats["theta"] - e0) <= 1e-9 * abs(e0)


@pytest.mark.parametrize("ncv,max_matvec,tol", [(0, 1000, 1e-10), (6, 2000, 1e-10), (0, 7, 1e-10), (0, 1000, 1e-3), (0, 1, 1e-10)])
def test_eig_lowest_image_is_heff_of_psi(cu, ncv, max_matvec, tol):
    """tnpy_eig_lowest_image: the H_eff psi it returns from the Lanczos relation equals a fresh matvec of the
    returned psi -- converged, restarted, cut off after a few steps, loosely converged, one-step solves."""
    env, mpo, mps = canonical_problem(12, 32, 6, seed=3)
    site = 6
    L, W, R = dev(env.left[site]), dev(mpo[site]), dev(env.right[site])
    psi = dev(np.random.default_rng(1).standard_normal(mps[site].shape))
    image = torch.empty_like(psi)
    stats = cu.eig_lowest(L, W, R, psi, tol=tol, ncv=ncv, max_matvec=max_matvec, image=image)
    fresh = cu.heff_apply(L, W, R, psi)
    scale = float(fresh.abs().max())
    assert float((image - fresh).abs().max()) <= 1e-12 * max(scale, stats["anorm"])
    # and the residual the solver reported is the norm of H psi - theta psi
    r = fresh - stats["theta"] * psi
    assert abs(float(r.norm()) - stats["resid"]) <= 1e-10 * stats["anorm"]


@pytest.mark.parametrize("n_sites,chi,site,model", [(10, 16, 4, "xxz"), (12, 24, 6, "thirring"), (14, 60, 7, "xxz"),
                                                     (12, 33, 1, "xxz"), (16, 128, 8, "thirring"), (18, 200, 9, "xxz")])
@pytest.mark.parametrize("ncv", [0, 6])
def test_fused_small_site_steps_match_the_general_solver(cu, n_sites, chi, site, model, ncv):
    """Small sites run whole Lanczos steps in one cooperative launch (csrc/lanczos_steps.cu), mid-size sites the
    Gram-Schmidt half of a step: same eigenvalue and eigenvector as the general multi-kernel solver on the same start
    vector, the image still equal to a fresh matvec, far fewer launches; edge-of-chain shapes (left bond 2), restarts
    (ncv = 6), the longest vector the whole-step path takes (chi = 128: 32768 elements), a mid-size site (chi = 200:
    80000 elements), and bit-identical results on repetition."""
    env, mpo, mps = canonical_problem(n_sites, chi, site, model=model, seed=5)
    L, W, R = dev(env.left[site]), dev(mpo[site]), dev(env.right[site])
    start = np.random.default_rng(2).standard_normal(mps[site].shape)
    out = {}
    for fused in (True, False):
        previous = cu.set_fused_steps(fused)
        try:
            psi = dev(start)
            image = torch.empty_like(psi)
            before = cu.launch_count()
            stats = cu.eig_lowest(L, W, R, psi, tol=1e-10, ncv=ncv, max_matvec=4000, image=image)
            out[fused] = (stats, psi, image, cu.launch_count() - before)
        finally:
            cu.set_fused_steps(previous)
    (sf, pf, imf, lf), (sg, pg, img, lg) = out[True], out[False]
    assert sf["converged"] and sg["converged"]
    assert sf["heff_mode"] == cu.HEFF_FP64_CHAIN
    anorm = sg["anorm"]
    assert abs(sf["theta"] - sg["theta"]) <= 1e-11 * anorm
    # residuals at 1e-10 ||A||: the eigenvectors agree to that divided by the gap; compare through the overlap
    assert abs(abs(float((pf * pg).sum())) - 1.0) <= 1e-8
    fresh = cu.heff_apply(L, W, R, pf)
    assert float((imf - fresh).abs().max()) <= 1e-12 * max(float(fresh.abs().max()), anorm)
    r = fresh - sf["theta"] * pf
    assert float(r.norm()) <= 1e-10 * anorm * 1.01
    assert abs(float(r.norm()) - sf["resid"]) <= 1e-10 * anorm
    if sg["n_matvec"] >= 20:
        # whole steps in one launch below 32768 unknowns; above, the matvec keeps its GEMM launches and the
        # Gram-Schmidt half of a step is one launch instead of nine
        assert lf * (3 if psi.numel() <= 32768 else 2) < lg, (lf, lg)
    # deterministic
    psi2 = dev(start)
    cu.eig_lowest(L, W, R, psi2, tol=1e-10, ncv=ncv, max_matvec=4000)
    assert torch.equal(psi2, pf)


def test_fused_steps_handle_a_vector_shorter_than_the_basis(cu):
    """N = 8 unknowns: the Krylov space is exhausted (exact breakdown or m == N) before the basis is full."""
    env, mpo, mps = canonical_problem(8, 2, 4, seed=1)
    site = 4
    H = env.one_site_full_matrix(site)
    evals = np.linalg.eigvalsh(0.5 * (H + H.T))
    psi = dev(np.random.default_rng(3).standard_normal(mps[site].shape))
    stats = cu.eig_lowest(dev(env.left[site]), dev(mpo[site]), dev(env.right[site]), psi, tol=1e-12)
    assert stats["converged"]
    assert abs(stats["theta"] - evals[0]) <= 1e-12 * abs(evals).max()


@pytest.mark.parametrize("n", [4, 16, 64, 100, 199])
def test_eigh_lowest(cu, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a = a + a.T
    ev, vec = cu.eigh_lowest(dev(a))
    w, v = np.linalg.eigh(a)
    assert abs(ev.item() - w[0]) <= 1e-12 * np.abs(w).max()
    x = vec.cpu().numpy()
    assert np.linalg.norm(a @ x - ev.item() * x) <= 1e-10 * np.abs(w).max()


# -------------------------------------------------------------------------------------------- SVD
@pytest.mark.parametrize("rows,cols", [(2, 2), (4, 2), (2, 4), (16, 8), (8, 16), (64, 32), (120, 60), (60, 120),
                                        (512, 256), (256, 512), (300, 300)])
def test_svd(cu, rows, cols):
    rng = np.random.default_rng(rows * 1000 + cols)
    a = rng.standard_normal((rows, cols)) * np.logspace(0, -6, cols)[None, :]
    u, s, vt = (t.cpu().numpy() for t in cu.svd(dev(a)))
    k = min(rows, cols)
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.all(np.diff(s) <= 0)
    assert np.abs(s - s_ref).max() <= 1e-12 * s_ref[0]
    assert np.abs(u @ np.diag(s) @ vt - a).max() <= 1e-12 * s_ref[0]
    assert np.abs(u.T @ u - np.eye(k)).max() <= 1e-11
    assert np.abs(vt @ vt.T - np.eye(k)).max() <= 1e-11


@pytest.mark.parametrize("rows,cols", [(2048, 1024), (1024, 2048), (1000, 700)])
def test_svd_block_path_large(cu, rows, cols):
    """Block-Jacobi path at a size where a host SVD would take seconds: checked through
    size-independent properties (orthogonality, reconstruction, sortedness) and against the
    singular values of cuSOLVER (yardstick only)."""
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    a = torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda")
    a = a * torch.logspace(0, -8, cols, dtype=torch.float64, device="cuda")[None, :]
    u, s, vt = cu.svd(a.clone())
    k = min(rows, cols)
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    assert float((u.t() @ u - eye).abs().max()) < 1e-11
    assert float((vt @ vt.t() - eye).abs().max()) < 1e-11
    assert float(((u * s) @ vt - a).abs().max()) < 1e-12 * float(s[0])
    assert bool((s[1:] <= s[:-1]).all())
    s_ref = torch.linalg.svdvals(a)
    assert float((s - s_ref).abs().max()) < 1e-12 * float(s_ref[0])
    assert cu.load().tnpy_last_svd_sweeps() < 60  # 60 = the sweep cap: convergence was reached


def test_svd_graded_spectrum(cu):
    """DMRG wave functions have singular values over 16 decades; small ones must keep absolute 1e-9."""
    rng = np.random.default_rng(7)
    q1, _ = np.linalg.qr(rng.standard_normal((128, 64)))
    q2, _ = np.linalg.qr(rng.standard_normal((64, 64)))
    s_true = np.logspace(0, -15, 64)
    a = (q1 * s_true) @ q2.T
    _, s, _ = cu.svd(dev(a))
    assert np.abs(s.cpu().numpy() - s_true).max() < 1e-14


def test_absorb(cu):
    rng = np.random.default_rng(3)
    k, n, cols, rows = 24, 24, 80, 70
    s, vt, nb = rng.random(k), rng.standard_normal((k, n)), rng.standard_normal((n, cols))
    got = cu.absorb_right(dev(s), dev(vt), dev(nb)).cpu().numpy()
    assert rel_err(got, np.diag(s) @ vt @ nb) < 1e-13
    u, nb2 = rng.standard_normal((n, k)), rng.standard_normal((rows, n))
    got = cu.absorb_left(dev(u), dev(s), dev(nb2)).cpu().numpy()
    assert rel_err(got, nb2 @ u @ np.diag(s)) < 1e-13


# ------------------------------------------------------------- BASELINE-size, size-independent checks
@pytest.mark.parametrize("chi,w", [(2048, 5), (1024, 6)])
def test_heff_properties_at_bench_size(cu, chi, w):
    """At BASELINE's chi the oracle would need minutes, so the chain is checked through properties:
    linearity, the transpose identity <x|H(L,W,R) y> = <H(L',W',R') x|y> with ket/bra legs swapped
    (H_eff is symmetric only for genuine environments; this identity holds for any operands), and
    agreement of the row-sharded entry point with the full matvec."""
    d = 2
    g = torch.Generator(device="cuda").manual_seed(chi)
    rnd = lambda *s: torch.randn(s, generator=g, dtype=torch.float64, device="cuda")  # noqa: E731
    L, R, W = rnd(chi, w, chi), rnd(chi, w, chi), rnd(w, w, d, d)
    x, y = rnd(chi, d, chi), rnd(chi, d, chi)
    hx, hy = cu.heff_apply(L, W, R, x), cu.heff_apply(L, W, R, y)
    hxy = cu.heff_apply(L, W, R, x + 0.5 * y)
    scale = float(hx.abs().max())
    assert float((hxy - (hx + 0.5 * hy)).abs().max()) < 1e-12 * scale
    Lt, Rt = L.permute(2, 1, 0).contiguous(), R.permute(2, 1, 0).contiguous()
    Wt = W.permute(0, 1, 3, 2).contiguous()
    lhs = float((x * hy).sum())
    rhs = float((cu.heff_apply(Lt, Wt, Rt, x) * y).sum())
    assert abs(lhs - rhs) < 1e-11 * abs(lhs) + 1e-9 * scale
    lo, hi = chi // 4, chi // 2
    rows = cu.heff_apply_rows(L[:, :, lo:hi].contiguous(), W, R, x, row0=lo)
    assert float((rows - hx[lo:hi]).abs().max()) < 1e-12 * scale


def _bench_size_operands(chi, w, canonical):
    """Seeded operands at a BASELINE bulk-site size with the model's real MPO tensor (XXZ for w = 5, Thirring
    with the script's penalty for w = 6); ``canonical`` plants the identity channels of the mixed-canonical gauge."""
    rng = np.random.default_rng(7 * chi + w)
    W = np.ascontiguousarray(oracle.xxz_mpo(4, 0.5)[1] if w == 5 else oracle.thirring_mpo(4, 0.5, 1.0, 100.0, 0)[1])
    assert W.shape[0] == w
    L, R = rng.standard_normal((chi, w, chi)), rng.standard_normal((chi, w, chi))
    if canonical:
        L[:, 0, :] = np.eye(chi)
        R[:, w - 1, :] = np.eye(chi)
    x = rng.standard_normal((chi, 2, chi))
    return L, W, R, x


@pytest.mark.parametrize("chi,w", [(2048, 5), (1024, 6)])
@pytest.mark.parametrize("flags", [0, 3])
def test_chains_match_oracle_at_bench_size(cu, chi, w, flags):
    """BASELINE configs[2] (chi=2048, w=5) and configs[3]'s bond (chi=1024, here with the w=6 Thirring tensor):
    the H_eff matvec and both environment updates against the CPU oracle itself (a few seconds of host BLAS per
    case), on the native FP64 DMMA path and on the tcgen05 path, general operands (flags 0) and the
    mixed-canonical gauge with both identity channels flagged (flags 3).  Tolerance 1e-12 max|y|."""
    L, W, R, x = _bench_size_operands(chi, w, canonical=bool(flags))
    want_y = oracle.heff_apply(L, W, R, x)
    want_l = oracle.env_update_left(L, x, W)
    want_r = oracle.env_update_right(R, x, W)
    dL, dW, dR, dx = dev(L), dev(W), dev(R), dev(x)
    for algo in (cu.GEMM_FP64, cu.GEMM_OZAKI):
        cu.set_gemm_algo(algo)
        try:
            got_y = cu.heff_apply(dL, dW, dR, dx, flags=flags).cpu().numpy()
            got_l = cu.env_update_left(dL, dx, dW, flags=flags & cu.LEFT_IDENTITY).cpu().numpy()
            got_r = cu.env_update_right(dR, dx, dW, flags=flags & cu.RIGHT_IDENTITY).cpu().numpy()
        finally:
            cu.set_gemm_algo(cu.GEMM_AUTO)
        for got, want, what in ((got_y, want_y, "heff"), (got_l, want_l, "env_left"), (got_r, want_r, "env_right")):
            err = np.abs(got - want).max() / np.abs(want).max()
            assert err < 1e-12, (what, algo, err)


def test_env_update_identity_channel_at_bench_size(cu):
    """Left-canonical site tensor + identity incoming channel => identity outgoing channel
    (the canonical-gauge invariant of SURVEY 8c) at chi = 1024."""
    from tnpy_b200.model import XXZ

    chi, d, w = 1024, 2, 5
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn((chi * d, chi), generator=g, dtype=torch.float64, device="cuda")
    q, _ = torch.linalg.qr(a)
    A = q.reshape(chi, d, chi).contiguous()
    W = torch.from_numpy(np.ascontiguousarray(XXZ(n=4, delta=0.5).mpo.as_four_leg(1))).cuda()
    L = torch.randn((chi, w, chi), generator=g, dtype=torch.float64, device="cuda")
    L[:, 0, :] = torch.eye(chi, dtype=torch.float64, device="cuda")
    out = cu.env_update_left(L, A, W)
    eye = torch.eye(chi, dtype=torch.float64, device="cuda")
    assert float((out[:, 0, :] - eye).abs().max()) < 1e-12
    # mirror symmetry: update_right of the mirrored tensor equals update_left
    Am = A.permute(2, 1, 0).contiguous()
    Wm = W.permute(1, 0, 2, 3).contiguous()
    out_r = cu.env_update_right(L, Am, Wm)
    assert float((out_r - out).abs().max()) < 1e-11 * float(out.abs().max())


def test_svd_round_trip_at_bench_size(cu):
    rows, cols = 4096, 2048
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda")
    u, s, vt = cu.svd(a.clone())
    eye = torch.eye(cols, dtype=torch.float64, device="cuda")
    assert float((u.t() @ u - eye).abs().max()) < 1e-11
    assert float((vt @ vt.t() - eye).abs().max()) < 1e-11
    assert float(((u * s) @ vt - a).abs().max()) < 1e-11 * float(s[0])
    assert bool((s[1:] <= s[:-1]).all())
    # checksum of checksums: sum of squared singular values == squared Frobenius norm
    assert abs(float((s * s).sum()) / float((a * a).sum()) - 1) < 1e-12


# ------------------------------------------------------------- canonical-gauge identity channels
@pytest.mark.parametrize("l,r,wl,wr,d", [(16, 16, 5, 5, 2), (64, 96, 5, 6, 2), (128, 128, 6, 6, 2), (33, 17, 5, 5, 2), (8, 8, 1, 5, 2)])
@pytest.mark.parametrize("flags", [1, 2, 3])
def test_identity_channel_shortcuts(cu, l, r, wl, wr, d, flags):
    """With L[:, 0, :] = I and R[:, wr-1, :] = I the flagged chains must reproduce the dense ones."""
    rng = np.random.default_rng(l + r + flags)
    L, W, R, x = random_operands(rng, l, r, wl, wr, d)
    L[:, 0, :] = np.eye(l)
    R[:, wr - 1, :] = np.eye(r)
    Ld, Wd, Rd, xd = dev(L), dev(W), dev(R), dev(x)
    assert cu.identity_defect(Ld, 0) == 0.0 and cu.identity_defect(Rd, wr - 1) == 0.0
    assert cu.identity_defect(Rd, 0) > 0.1
    want = oracle.heff_apply(L, W, R, x)
    got = cu.heff_apply(Ld, Wd, Rd, xd, flags=flags).cpu().numpy()
    assert rel_err(got, want) < 1e-12
    if flags & 1:
        got = cu.env_update_left(Ld, xd, Wd, flags=1).cpu().numpy()
        assert rel_err(got, oracle.env_update_left(L, x, W)) < 1e-12
    if flags & 2:
        got = cu.env_update_right(Rd, xd, Wd, flags=2).cpu().numpy()
        assert rel_err(got, oracle.env_update_right(R, x, W)) < 1e-12


def test_identity_flags_are_measured_not_assumed():
    """Environment sets the flags only where the channel really is the identity: a right-canonical
    random MPS has identity right channels and non-identity left channels."""
    from tnpy_b200.matrix_product_state import Environment, MatrixProductState
    from tnpy_b200.model import XXZ

    env = Environment(XXZ(n=10, delta=0.5).mpo, MatrixProductState.random(10, 16, 2, seed=2))
    assert env.gauge_flags(5) == cu_flags("right")
    assert env.gauge_flags(0) == cu_flags("right")
    off = Environment(XXZ(n=10, delta=0.5).mpo, MatrixProductState.random(10, 16, 2, seed=2), use_identity_channels=False)
    assert off.gauge_flags(5) == 0


def cu_flags(which):
    from tnpy_b200 import _cuda

    return {"left": _cuda.LEFT_IDENTITY, "right": _cuda.RIGHT_IDENTITY}[which]


@pytest.mark.parametrize("rows,cols,rank", [(16, 8, 5), (8, 16, 3), (300, 200, 120), (200, 300, 0)])
def test_svd_exact_null_space(cu, rows, cols, rank):
    """Exactly rank-deficient input (product-state-like site tensors): zero singular values are reported
    as 0 and U / Vt are still complete isometries (LAPACK's orthonormal completion)."""
    rng = np.random.default_rng(rows + cols + rank)
    a = rng.standard_normal((rows, rank)) @ rng.standard_normal((rank, cols)) if rank else np.zeros((rows, cols))
    k = min(rows, cols)
    if rank:  # make the null space exact: zero out whole columns / rows of a factor instead of relying on rounding
        a = np.zeros((rows, cols))
        a[:, :rank] = rng.standard_normal((rows, rank))
        if rows < cols:
            a = np.zeros((rows, cols))
            a[:rank, :] = rng.standard_normal((rank, cols))
    u, s, vt = (t.cpu().numpy() for t in cu.svd(dev(a)))
    s_ref = np.linalg.svd(a, compute_uv=False)
    assert np.abs(s - s_ref).max() <= 1e-12 * max(s_ref[0], 1e-300) + 1e-130
    assert np.all(s[rank:] < 1e-100)
    assert np.abs(u.T @ u - np.eye(k)).max() < 1e-11
    assert np.abs(vt @ vt.T - np.eye(k)).max() < 1e-11
    assert np.abs(u @ np.diag(s) @ vt - a).max() <= 1e-12 * max(s_ref[0], 1.0)


# --------------------------------------------------- tcgen05 int8 (Ozaki) FP64 GEMM
@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (130, 70, 100), (512, 384, 200), (1000, 999, 777), (2048, 1024, 4096)])
def test_ozaki_gemm_matches_fp64(cu, m, n, k):
    """8 slices: componentwise error relative to |A|^T |B| at the level of a plain FP64 GEMM, also for
    operands whose columns span many orders of magnitude."""
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    a = a * torch.logspace(0, -9, m, dtype=torch.float64, device="cuda")[None, :]
    b = b * torch.logspace(3, -3, n, dtype=torch.float64, device="cuda")[None, :]
    ref = a.t() @ b
    bound = a.abs().t() @ b.abs()
    got = cu.ozaki_gemm_tn(a, b, slices=8)
    assert float(((got - ref).abs() / bound).max()) < 4e-15
    c0 = torch.randn((m, n), generator=g, dtype=torch.float64, device="cuda")
    acc = cu.ozaki_gemm_tn(a, b, out=c0.clone(), slices=8, accumulate=True)
    assert float(((acc - (c0 + ref)).abs() / (bound + c0.abs())).max()) < 4e-15
    loose = cu.ozaki_gemm_tn(a, b, slices=6)
    assert float(((loose - ref).abs() / bound).max()) < 1e-10


def test_ozaki_handles_zero_columns_and_ragged_k(cu):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn((333, 256), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((333, 128), generator=g, dtype=torch.float64, device="cuda")
    a[:, 5] = 0.0
    b[:, 7] = 0.0
    got = cu.ozaki_gemm_tn(a, b)
    ref = a.t() @ b
    assert float((got - ref).abs().max()) < 1e-12
    assert float(got[5].abs().max()) == 0.0 and float(got[:, 7].abs().max()) == 0.0


@pytest.mark.parametrize("chi,w", [(1024, 5), (1280, 6)])
def test_chain_with_ozaki_gemm(cu, chi, w):
    """The matvec / environment chains on the default selection (large GEMMs on the tcgen05 path) agree with the
    native FP64 chains."""
    d = 2
    g = torch.Generator(device="cuda").manual_seed(chi)
    rnd = lambda *s: torch.randn(s, generator=g, dtype=torch.float64, device="cuda")  # noqa: E731
    L, R, W, x = rnd(chi, w, chi), rnd(chi, w, chi), rnd(w, w, d, d), rnd(chi, d, chi)
    n0 = cu.launch_count()
    got = cu.heff_apply(L, W, R, x)
    launches_default = cu.launch_count() - n0
    got_env = cu.env_update_left(L, x, W)
    got_env_r = cu.env_update_right(R, x, W)
    cu.set_gemm_algo(cu.GEMM_FP64)
    try:
        n0 = cu.launch_count()
        ref = cu.heff_apply(L, W, R, x)
        launches_fp64 = cu.launch_count() - n0
        ref_env = cu.env_update_left(L, x, W)
        ref_env_r = cu.env_update_right(R, x, W)
    finally:
        cu.set_gemm_algo(cu.GEMM_AUTO)
    assert launches_default > launches_fp64  # the default really took the sliced path
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-13
    assert float((got_env - ref_env).abs().max() / ref_env.abs().max()) < 1e-13
    assert float((got_env_r - ref_env_r).abs().max() / ref_env_r.abs().max()) < 1e-13


@pytest.mark.parametrize("m,n,k", [(256, 128, 64), (700, 300, 1000), (1024, 1024, 1024), (520, 390, 4160), (4096, 2048, 640)])
def test_ozaki_gemm_shapes_slices_and_bound(cu, m, n, k):
    """The CTA-pair kernel with and without the K-split of its last wave (1024^3 and 520x390x4160 exercise the
    split, 4096x2048x640 a full wave): exact slice GEMMs, so the result agrees with FP64 at the DMMA level with 8
    slices, degrades by 2^7 per dropped slice, and always stays inside the rigorous bound the library reports."""
    g = torch.Generator(device="cuda").manual_seed(7 * m + n + k)
    a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    ref = a.t() @ b
    bound = a.abs().t() @ b.abs()
    for s, tol in ((8, 4e-15), (7, 2e-13), (6, 3e-11), (5, 4e-9)):
        got = cu.ozaki_gemm_tn(a, b, slices=s)
        assert float(((got - ref).abs() / bound).max()) < tol
        rigorous = cu.ozaki_error_bound(m, n, k, slices=s)
        err = float((got - ref).norm())
        assert err <= rigorous + 1e-14 * float(ref.norm()), (s, err, rigorous)  # + the FP64 rounding of ref itself
        assert rigorous < 1e-12 * 128.0 ** (8 - s) * float(bound.norm())  # and it is not vacuous
    c0 = torch.randn((m, n), generator=g, dtype=torch.float64, device="cuda")
    acc = cu.ozaki_gemm_tn(a, b, out=c0.clone(), slices=8, accumulate=True)
    assert float(((acc - (c0 + ref)).abs() / (bound + c0.abs())).max()) < 4e-15


def test_heff_plan_slices_environments_once(cu):
    """A prepared H_eff slices the environments at creation; every application then only slices x-side operands.
    Results are bit-identical to the one-shot entry point, and the plan notices nothing it should not (a changed
    environment needs a new plan -- HeffOperator tracks that through Environment.operand_version)."""
    chi, w, d = 1024, 5, 2
    g = torch.Generator(device="cuda").manual_seed(3)
    rnd = lambda *s: torch.randn(s, generator=g, dtype=torch.float64, device="cuda")  # noqa: E731
    L, R, W, x = rnd(chi, w, chi), rnd(chi, w, chi), rnd(w, w, d, d), rnd(chi, d, chi)
    plain = cu.heff_apply(L, W, R, x).clone()
    n0 = cu.launch_count()
    cu.heff_apply(L, W, R, x)
    launches_plain = cu.launch_count() - n0
    plan = cu.HeffPlan(L, W, R, chi, chi)
    assert plan.mode == cu.HEFF_OZ_CHAIN  # a dense random W has interior blocks: the chain, on tcgen05
    first = plan.apply(x).clone()
    n0 = cu.launch_count()
    second = plan.apply(2.0 * x).clone()
    launches_plan = cu.launch_count() - n0
    assert torch.equal(first, plain)
    assert float((second - 2.0 * plain).abs().max() / plain.abs().max()) < 1e-14
    assert launches_plan == launches_plain - 4  # colmax + slice kernels of L and of R are not repeated
    assert 0.0 < plan.error_bound() < 1e-9 * float(plain.norm())
    fp64 = cu.HeffPlan(L, W, R, chi, chi, algo=cu.GEMM_FP64)
    assert fp64.mode == cu.HEFF_FP64_CHAIN and fp64.error_bound() == 0.0
    ref = fp64.apply(x)
    assert float((first - ref).abs().max() / ref.abs().max()) < 1e-13
    plan.close(); fp64.close()


@pytest.mark.parametrize("l,r,model", [(1024, 1024, "xxz"), (1200, 1000, "xxz"), (1000, 1204, "rh"), (1024, 1024, "thirring")])
def test_direct_path_in_canonical_gauge(cu, l, r, model):
    """Mixed-canonical gauge + an MPO tensor without interior-to-interior blocks => the direct path (two independent
    tcgen05 GEMMs on operands premixed and sliced from x, no FP64 intermediate).  Checked against the oracle and the
    FP64 chain, with 8 and 7 slices, on aligned and ragged bonds; Thirring with its penalty channel (W[3,3] = I)
    must be sent to the chain instead."""
    mpo = {"xxz": oracle.xxz_mpo(4, 0.5), "rh": oracle.random_heisenberg_mpo(4, 1.0, seed=2022),
           "thirring": oracle.thirring_mpo(4, 0.5, 1.0, 100.0, 0)}[model]
    W = np.ascontiguousarray(mpo[1])
    w, d = W.shape[0], W.shape[2]
    rng = np.random.default_rng(l + 3 * r)
    L, R = rng.standard_normal((l, w, l)), rng.standard_normal((r, w, r))
    L[:, 0, :] = np.eye(l)
    R[:, w - 1, :] = np.eye(r)
    x = rng.standard_normal((l, d, r))
    want = oracle.heff_apply(L, W, R, x)
    dL, dW, dR, dx = dev(L), dev(W), dev(R), dev(x)
    plan = cu.HeffPlan(dL, dW, dR, l, r, flags=3, w_host=W)
    assert plan.mode == (cu.HEFF_OZ_CHAIN if model == "thirring" else cu.HEFF_OZ_DIRECT)
    got8 = plan.apply(dx).cpu().numpy()
    scale = np.abs(want).max()
    assert np.abs(got8 - want).max() < 1e-12 * scale
    b8 = plan.error_bound()
    got7 = plan.apply(dx, slices=7).cpu().numpy()
    assert np.abs(got7 - want).max() < 1e-10 * scale
    assert np.linalg.norm(got8 - want) <= b8 + 1e-13 * np.linalg.norm(want)
    assert np.linalg.norm(got7 - want) <= plan.error_bound() + 1e-13 * np.linalg.norm(want)
    # W_host not given: the library reads W back itself and reaches the same decision
    plan2 = cu.HeffPlan(dL, dW, dR, l, r, flags=3)
    assert plan2.mode == plan.mode and torch.equal(plan2.apply(dx), plan.apply(dx))
    chain = cu.HeffPlan(dL, dW, dR, l, r, flags=3, algo=cu.GEMM_FP64)
    assert np.abs(chain.apply(dx).cpu().numpy() - got8).max() < 1e-12 * scale
    for p_ in (plan, plan2, chain):
        p_.close()


def test_inexact_slice_schedule_is_accepted_on_the_true_residual(cu):
    """On the tcgen05 path the eigensolver runs the later steps of a solve with fewer int8 slices and accepts the result
    only on the true residual (one matvec at full accuracy).  XXZ n=22 chi=1024 in the mixed-canonical gauge: the
    schedule is used, no check fails, the eigenvalue equals the full-accuracy solve's to 1e-12 |theta|, the true
    residual holds the tolerance, and the returned image is H_eff psi."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    import bench
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.model import XXZ

    n, chi, tol = 22, 1024, 1e-8
    dmrg = FiniteDMRG(XXZ(n=n, delta=0.5).mpo, bond_dim=chi, mps=bench.random_right_canonical_device(n, chi, 2, seed=3),
                      compute_variance=False)
    site = n // 2
    bench.mixed_canonicalize(dmrg, site)
    env = dmrg.environment
    L, W, R = env.operands(site)
    flags = env.gauge_flags(site)
    assert flags == 3
    start = env.device_tensor(site).clone()
    out = {}
    for on in (True, False):
        previous = cu.set_inexact_slices(on)
        try:
            psi = start.clone()
            image = torch.empty_like(psi)
            stats = cu.eig_lowest(L, W, R, psi, tol=tol, max_matvec=3000, flags=flags, image=image)
            out[on] = (stats, psi, image)
        finally:
            cu.set_inexact_slices(previous)
    (s_on, p_on, im_on), (s_off, p_off, _) = out[True], out[False]
    assert s_on["converged"] and s_off["converged"]
    assert s_on["heff_mode"] == cu.HEFF_OZ_DIRECT
    assert s_on["reduced_slice_matvecs"] > 0 and s_off["reduced_slice_matvecs"] == 0
    assert s_on["failed_residual_checks"] == 0
    assert abs(s_on["theta"] - s_off["theta"]) <= 1e-12 * abs(s_off["theta"])
    assert abs(abs(float((p_on * p_off).sum())) - 1.0) <= 1e-8
    # the true residual, from a fresh FP64 matvec
    fresh = cu.heff_apply(L, W, R, p_on, flags=flags)
    r = fresh - s_on["theta"] * p_on
    assert float(r.norm()) <= 1.05 * tol * s_on["anorm"]
    assert abs(float(r.norm()) - s_on["resid"]) <= 1e-3 * tol * s_on["anorm"]
    assert float((im_on - fresh).abs().max()) <= 1e-11 * max(float(fresh.abs().max()), 1.0)
    assert s_on["n_matvec"] <= s_off["n_matvec"] + 8


def test_row_block_of_the_direct_path(cu):
    """One rank's row block of the chi-sharded matvec in the mixed-canonical gauge: the identity flags are honoured
    for a block (L[li, 0, row0 + m] = delta), the R-side term reads only the block's own rows of x, and the result is
    the corresponding rows of the full matvec -- on the direct path, the tcgen05 chain and the FP64 chain."""
    chi, w, d = 2048, 5, 2
    L, W, R, x = _bench_size_operands(chi, w, canonical=True)
    dL, dW, dR, dx = dev(L), dev(W), dev(R), dev(x)
    full = cu.HeffPlan(dL, dW, dR, chi, chi, flags=3, w_host=W)
    want = full.apply(dx)
    full.close()
    lo, hi = 512, 1536
    L_rows = dL[:, :, lo:hi].contiguous()
    for algo, flags, mode in ((cu.GEMM_AUTO, 3, cu.HEFF_OZ_DIRECT), (cu.GEMM_AUTO, 0, cu.HEFF_OZ_CHAIN), (cu.GEMM_FP64, 3, cu.HEFF_FP64_CHAIN)):
        plan = cu.HeffPlan(L_rows, dW, dR, chi, chi, flags=flags, algo=algo, w_host=W, l_rows=hi - lo, row0=lo)
        assert plan.mode == mode
        got = plan.apply(dx)
        assert tuple(got.shape) == (hi - lo, d, chi)
        assert float((got - want[lo:hi]).abs().max() / want.abs().max()) < 1e-13
        plan.close()
    one_shot = cu.heff_apply_rows(L_rows, dW, dR, dx, row0=lo, flags=3)
    assert float((one_shot - want[lo:hi]).abs().max() / want.abs().max()) < 1e-13


# ---- a8/a9 as an orthogonal split: tnpy_qr_split (Cholesky-QR twice, verified on the device) ---------------
@pytest.mark.parametrize("rows,cols", [(8, 4), (4, 8), (64, 64), (128, 64), (64, 128), (200, 100), (130, 300),
                                       (520, 260), (1000, 700), (2048, 1024), (1024, 2048), (4096, 2048)])
def test_qr_split(cu, rows, cols):
    """A = Q T (tall) / T Q (wide): Q orthonormal to rounding, exact reconstruction, the singular values of the
    small factor T are those of A (graded columns / rows over 10 decades, as a DMRG site tensor has)."""
    g = torch.Generator(device="cuda").manual_seed(3 * rows + cols)
    a = torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda")
    k = min(rows, cols)
    grade = torch.logspace(0, -10, k, dtype=torch.float64, device="cuda")
    a = a * grade[None, :] if rows >= cols else a * grade[:, None]
    keep = a.clone()
    q, t, defect = cu.qr_split(a)
    assert torch.equal(a, keep)  # input untouched
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    gram = q.t() @ q if rows >= cols else q @ q.t()
    measured = float((gram - eye).abs().max())
    assert measured < 1e-13
    assert defect < 1e-13 and abs(defect - measured) < 5e-15  # the device-side check measures the same thing
    back = q @ t if rows >= cols else t @ q
    assert float((back - a).abs().max()) < 1e-13 * float(a.abs().max())
    s_ref = torch.linalg.svdvals(a)
    s_t = torch.linalg.svdvals(t)
    assert float((s_t - s_ref).abs().max()) < 1e-12 * float(s_ref[0])


@pytest.mark.parametrize("n", [64, 100, 512])
@pytest.mark.parametrize("t_first", [False, True])
def test_qr_split_square_factor_order(cu, n, t_first):
    """A square matrix is split as Q T or, with TNPY_QR_T_FIRST, as T Q -- the order a leftward split of a
    square site tensor (l == d r, the ramp of the chain) needs; the two are different factorisations."""
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn((n, n), generator=g, dtype=torch.float64, device="cuda")
    grade = torch.logspace(0, -8, n, dtype=torch.float64, device="cuda")
    a = (a * grade[:, None] if t_first else a * grade[None, :]).contiguous()  # graded along the vectors being orthonormalised
    q, t, defect = cu.qr_split(a, t_first=t_first)
    assert defect < 1e-13
    good, bad = (t @ q, q @ t) if t_first else (q @ t, t @ q)
    assert float((good - a).abs().max()) < 1e-13
    assert float((bad - a).abs().max()) > 1e-3


def test_qr_split_reports_breakdown(cu):
    """A zero vector cannot be normalised: the defect comes back as inf, which is what sends split_tensor to
    the SVD (LAPACK-style null-space completion lives there).  Exactly dependent columns either break down or
    give a valid split (orthonormal Q, exact reconstruction) -- never a silently wrong one."""
    g = torch.Generator(device="cuda").manual_seed(5)
    b = torch.randn((128, 256), generator=g, dtype=torch.float64, device="cuda")
    b[17, :] = 0.0
    assert cu.qr_split(b)[2] == float("inf")
    assert cu.qr_split(b, shifted=True)[2] == float("inf")
    a = torch.randn((256, 128), generator=g, dtype=torch.float64, device="cuda")
    a[:, 70] = a[:, 3]
    q, t, defect = cu.qr_split(a)
    if defect <= 1e-13:
        assert float((q.t() @ q - torch.eye(128, dtype=torch.float64, device="cuda")).abs().max()) < 1e-13
        assert float((q @ t - a).abs().max()) < 1e-12


@pytest.mark.parametrize("rows,cols,cond", [(512, 256, 1e10), (2048, 1024, 1e12), (1024, 2048, 1e11)])
def test_qr_split_shifted_handles_ill_conditioned(cu, rows, cols, cond):
    """Vectors whose *normalised* Gram matrix has condition cond^2 (a cold-sweep site tensor: no grading to
    scale away): two Cholesky-QR passes are not enough -- and say so through the defect -- the shifted
    three-pass variant is."""
    k = min(rows, cols)
    g = torch.Generator(device="cuda").manual_seed(rows + 7)
    u, _ = torch.linalg.qr(torch.randn((max(rows, cols), k), generator=g, dtype=torch.float64, device="cuda"))
    v, _ = torch.linalg.qr(torch.randn((k, k), generator=g, dtype=torch.float64, device="cuda"))
    s = torch.logspace(0, -float(np.log10(cond)), k, dtype=torch.float64, device="cuda")
    a = (u * s) @ v.t()
    a = a.contiguous() if rows >= cols else a.t().contiguous()
    assert cu.qr_split(a)[2] > 1e-13
    q, t, defect = cu.qr_split(a, shifted=True)
    assert defect < 1e-13
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    gram = q.t() @ q if rows >= cols else q @ q.t()
    assert float((gram - eye).abs().max()) < 1e-13
    back = q @ t if rows >= cols else t @ q
    assert float((back - a).abs().max()) < 1e-13
    assert float((torch.linalg.svdvals(t) - s).abs().max()) < 1e-12  # absolute, s_max = 1 (north_star asks 1e-9)


def test_qr_split_matches_svd_split():
    """split_tensor in "qr" mode against the reference-literal "svd" mode on the same site tensor: same
    product A[site].A[site+1], same bond spectrum (deferred small SVD), isometric site tensor, both directions."""
    from tnpy_b200.matrix_product_state import DeferredSpectrum, Direction, _split_on_device

    g = torch.Generator(device="cuda").manual_seed(9)
    l, d, r = 96, 2, 128
    a = torch.randn((l, d, r), generator=g, dtype=torch.float64, device="cuda")
    a = a * torch.logspace(0, -9, r, dtype=torch.float64, device="cuda")[None, None, :]
    nb = torch.randn((r, d, 80), generator=g, dtype=torch.float64, device="cuda")
    theta = torch.einsum("lpr,rqs->lpqs", a, nb)
    for mode in ("qr", "svd"):
        q, new_nb, s = _split_on_device(a, nb, Direction.RIGHTWARD, mode, 16)
        assert isinstance(s, DeferredSpectrum) == (mode == "qr")
        assert float((torch.einsum("lpr,rqs->lpqs", q, new_nb) - theta).abs().max()) < 1e-12 * float(theta.abs().max())
        iso = q.reshape(l * d, r)
        assert float((iso.t() @ iso - torch.eye(r, dtype=torch.float64, device="cuda")).abs().max()) < 1e-12
        sv = s.cpu().numpy()
        s_ref = torch.linalg.svdvals(a.reshape(l * d, r)).cpu().numpy()
        assert np.abs(np.sort(sv)[::-1] - s_ref).max() < 1e-12 * s_ref[0]
    # square site tensors (l == d r leftward, l d == r rightward): the ramp sites of a chain
    sq = torch.randn((128, 2, 64), generator=g, dtype=torch.float64, device="cuda")
    nb_l = torch.randn((70, d, 128), generator=g, dtype=torch.float64, device="cuda")
    th = torch.einsum("lpr,rqs->lpqs", nb_l, sq)
    q, new_nb, s = _split_on_device(sq, nb_l, Direction.LEFTWARD, "qr", 16)
    assert isinstance(s, DeferredSpectrum)
    assert float((torch.einsum("lpr,rqs->lpqs", new_nb, q) - th).abs().max()) < 1e-12 * float(th.abs().max())
    sq = torch.randn((32, 2, 64), generator=g, dtype=torch.float64, device="cuda")
    nb_r = torch.randn((64, d, 50), generator=g, dtype=torch.float64, device="cuda")
    th = torch.einsum("lpr,rqs->lpqs", sq, nb_r)
    q, new_nb, s = _split_on_device(sq, nb_r, Direction.RIGHTWARD, "qr", 16)
    assert isinstance(s, DeferredSpectrum)
    assert float((torch.einsum("lpr,rqs->lpqs", q, new_nb) - th).abs().max()) < 1e-12 * float(th.abs().max())
    a2 = a.permute(2, 1, 0).contiguous()  # (128, 2, 96): wide as (l, d r)
    nb2 = torch.randn((40, d, 128), generator=g, dtype=torch.float64, device="cuda")
    theta2 = torch.einsum("lpr,rqs->lpqs", nb2, a2)
    for mode in ("qr", "svd"):
        q, new_nb, s = _split_on_device(a2, nb2, Direction.LEFTWARD, mode, 16)
        assert float((torch.einsum("lpr,rqs->lpqs", new_nb, q) - theta2).abs().max()) < 1e-12 * float(theta2.abs().max())
        iso = q.reshape(128, d * 96)
        assert float((iso @ iso.t() - torch.eye(128, dtype=torch.float64, device="cuda")).abs().max()) < 1e-12
