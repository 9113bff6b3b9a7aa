"""ShiftInvertDMRG (reference finite_dmrg.py:266-407, pinned by tests/test_finite_dmrg.py:26-72)."""
import numpy as np
import pytest
import scipy.linalg as spla

from oracle import tnpy_oracle as oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


@pytest.mark.parametrize("site", [0, 3, 7])
def test_geig_lowest_matches_dense_pencil(site):
    from tnpy_b200 import _cuda

    n, chi, offset = 8, 16, 0.1
    mpo = oracle.random_heisenberg_mpo(n, 10.5, seed=2022, offset=offset)
    mps = oracle.random_mps(n, chi, 2, seed=5)
    env, env2 = oracle.Environment(mpo, mps), oracle.Environment(oracle.mpo_square(mpo), mps)
    a, b = env.one_site_full_matrix(site), env2.one_site_full_matrix(site)
    want = spla.eigh(0.5 * (a + a.T), 0.5 * (b + b.T), eigvals_only=True, subset_by_index=[0, 0])[0]

    def ops(e, s):
        L = None if s == 0 else dev(e.left[s])
        R = None if s == n - 1 else dev(e.right[s])
        return L, dev(oracle._w4(e.mpo[s], s, n)), R

    # dense route (what ShiftInvertDMRG uses up to dense_pencil_dim unknowns)
    l, d, r = oracle._as3(mps[site], site, n).shape
    ad = _cuda.heff_dense(*ops(env, site), l, r)
    bd = _cuda.heff_dense(*ops(env2, site), l, r)
    theta, xd = _cuda.geig_dense_lowest(ad, bd)
    assert abs(theta.item() - want) <= 1e-7 * abs(want)
    x = xd.cpu().numpy()
    assert abs(x @ b.T @ x - 1.0) < 1e-6  # M-normalised
    assert np.linalg.norm(a.T @ x - theta.item() * (b.T @ x)) <= 1e-6 * np.linalg.norm(a.T @ x)
    # iterative route: exact when the basis can span the whole space (N <= 40)
    if a.shape[0] <= 40:
        psi = dev(oracle._as3(mps[site], site, n)).clone()
        stats = _cuda.geig_lowest(*ops(env, site), *ops(env2, site), psi, tol=1e-10)
        assert stats["converged"]
        assert abs(stats["theta"] - want) <= 1e-7 * abs(want)


@pytest.mark.parametrize("offset", [-0.1, 0.1, 0.2])
def test_shift_invert_reference_anchors(offset):
    """tests/test_finite_dmrg.py:56-72 at n=8, chi=16: nearest eigenvalue below the offset (atol 1e-6),
    the restored state equals the ED eigenvector up to a sign (atol 1e-6), and the returned energy
    equals <H> on the restored state (atol 1e-6)."""
    from tnpy_b200.finite_dmrg import ShiftInvertDMRG
    from tnpy_b200.model import RandomHeisenberg

    n, h, seed = 8, 10.5, 2022
    model = RandomHeisenberg(n=n, h=h, seed=seed)
    evals, evecs = np.linalg.eigh(oracle.full_hamiltonian(model.mpo.arrays))
    idx = np.where(evals < offset)[0].max()
    shifted = RandomHeisenberg(n=n, h=h, seed=seed, offset=offset)
    sidmrg = ShiftInvertDMRG(shifted.mpo, bond_dim=2**4, offset=offset, seed=1)
    with pytest.raises(RuntimeError):
        sidmrg.measurements
    energies = sidmrg.run(tol=1e-8)
    np.testing.assert_allclose(energies[-1], evals[idx], atol=1e-6)
    vec = sidmrg.restored_mps.to_dense()
    if not np.allclose(vec, evecs[:, idx], atol=1e-6):
        np.testing.assert_allclose(-vec, evecs[:, idx], atol=1e-6)
    np.testing.assert_allclose(energies[-1], sidmrg.measurements.expectation_value(model.mpo), atol=1e-6)


def test_shift_invert_parity_with_oracle():
    from tnpy_b200.finite_dmrg import ShiftInvertDMRG
    from tnpy_b200.matrix_product_state import MatrixProductState
    from tnpy_b200.model import RandomHeisenberg

    n, chi, offset = 8, 12, 0.1
    shifted = RandomHeisenberg(n=n, h=10.5, seed=2022, offset=offset)
    init = oracle.random_mps(n, chi, 2, seed=3)
    ref = oracle.ShiftInvertDMRG(shifted.mpo.arrays, chi, offset=offset, mps=[a.copy() for a in init])
    e_ref = ref.run(tol=0.0, max_sweep=4)  # tol 0: the |dE| < tol stop never fires, both sides do 4 sweeps
    gpu = ShiftInvertDMRG(shifted.mpo, bond_dim=chi, offset=offset, mps=MatrixProductState([a.copy() for a in init]))
    e_gpu = gpu.run(tol=0.0, max_sweep=4)
    assert len(e_gpu) == len(e_ref)
    np.testing.assert_allclose(e_gpu, e_ref, rtol=1e-8)
    a, b = oracle.mps_to_dense(ref.restored_mps), gpu.restored_mps.to_dense()
    assert abs(a @ b) / np.sqrt((a @ a) * (b @ b)) > 1 - 1e-8


def test_shift_invert_at_the_reference_test_size():
    """tests/test_finite_dmrg.py:26-72 as written: n=10, h=10.5, seed=2022, bond_dim 2**6 (bonds compress to at most 32,
    so the two mid-chain sites have 1024 unknowns: inside the dense-pencil range)."""
    from tnpy_b200.finite_dmrg import ShiftInvertDMRG
    from tnpy_b200.model import RandomHeisenberg

    n, h, seed, offset = 10, 10.5, 2022, 0.1
    model = RandomHeisenberg(n=n, h=h, seed=seed)
    evals, evecs = np.linalg.eigh(oracle.full_hamiltonian(model.mpo.arrays))
    idx = np.where(evals < offset)[0].max()
    shifted = RandomHeisenberg(n=n, h=h, seed=seed, offset=offset)
    sidmrg = ShiftInvertDMRG(shifted.mpo, bond_dim=2**6, offset=offset, seed=1)
    energies = sidmrg.run(tol=1e-8)
    assert all(st["dense"] for st in sidmrg.solver_stats)
    np.testing.assert_allclose(energies[-1], evals[idx], atol=1e-6)
    vec = sidmrg.restored_mps.to_dense()
    if not np.allclose(vec, evecs[:, idx], atol=1e-6):
        np.testing.assert_allclose(-vec, evecs[:, idx], atol=1e-6)
    np.testing.assert_allclose(energies[-1], sidmrg.measurements.expectation_value(model.mpo), atol=1e-6)


def test_unconverged_iterative_pencil_raises_and_keeps_the_state():
    """Above dense_pencil_dim the generalised Davidson is tried; if it does not converge the sweep stops with an
    error and the site tensor is what it was (never an unconverged vector passed off as the solution)."""
    from tnpy_b200.finite_dmrg import ShiftInvertDMRG
    from tnpy_b200.model import RandomHeisenberg

    shifted = RandomHeisenberg(n=10, h=10.5, seed=2022, offset=0.1)
    sidmrg = ShiftInvertDMRG(shifted.mpo, bond_dim=2**6, offset=0.1, seed=1)
    sidmrg.dense_pencil_dim = 512  # the two mid-chain sites (1024 unknowns) now iterate
    sidmrg.cholesky_pencil_dim = 0
    site = 5
    before = sidmrg.environment.device_tensor(site).clone()
    try:
        sidmrg._solve_on_device(site, 1e-8, maxiter=60)
    except RuntimeError as exc:
        assert "did not converge" in str(exc)
        assert torch.equal(sidmrg.environment.device_tensor(site), before)
    else:  # it converged within 60 iterations: then the answer must be the dense one
        theta_iter = sidmrg.solver_stats[-1]["theta"]
        sidmrg.environment.device_tensor(site).copy_(before)
        sidmrg.dense_pencil_dim = 2048
        theta_dense = sidmrg._solve_on_device(site, 1e-8)
        assert sidmrg.solver_stats[-1]["dense"]
        assert abs(theta_iter - theta_dense) <= 1e-6 * abs(theta_dense)


def test_cholesky_pencil_matches_lapack():
    """tnpy_geig_chol_lowest (Cholesky factor of the right-hand matrix + the on-device Lanczos solver on the reduced
    matrix) against scipy.linalg.eigh(a, b) -- LAPACK's sygvd, the same reduction -- on a projected pencil of 2048
    unknowns: eigenvalue, M-normalisation and the pencil residual; a and b are left intact."""
    from tnpy_b200 import _cuda

    n, chi, offset, site = 12, 32, 0.1, 6
    mpo = oracle.random_heisenberg_mpo(n, 10.5, seed=2022, offset=offset)
    mps = oracle.random_mps(n, chi, 2, seed=5)
    env, env2 = oracle.Environment(mpo, mps), oracle.Environment(oracle.mpo_square(mpo), mps)
    l, d, r = mps[site].shape
    assert l * d * r == 2048
    L, W, R = dev(env.left[site]), dev(oracle._w4(mpo[site], site, n)), dev(env.right[site])
    L2, W2, R2 = dev(env2.left[site]), dev(oracle._w4(env2.mpo[site], site, n)), dev(env2.right[site])
    ad = _cuda.heff_dense(L, W, R, l, r)
    bd = _cuda.heff_dense(L2, W2, R2, l, r)
    a, b = ad.cpu().numpy(), bd.cpu().numpy()
    a, b = 0.5 * (a + a.T), 0.5 * (b + b.T)
    want = spla.eigh(a, b, eigvals_only=True, subset_by_index=[0, 0])[0]
    a_before, b_before = ad.clone(), bd.clone()
    theta, xd, stats = _cuda.geig_chol_lowest(ad, bd, tol=1e-12)
    assert stats["converged"] and stats["n_matvec"] < 400
    assert torch.equal(ad, a_before) and torch.equal(bd, b_before)
    cond = np.linalg.cond(b)
    assert abs(theta.item() - want) <= max(1e-8, 10 * cond * 2.2e-16) * abs(want)
    x = xd.cpu().numpy()
    assert abs(x @ b @ x - 1.0) < 1e-6
    assert np.linalg.norm(a @ x - theta.item() * (b @ x)) <= 1e-6 * np.linalg.norm(a @ x)
    # and it agrees with the SVD / Jacobi route of the smaller sites
    theta_svd, _ = _cuda.geig_dense_lowest(ad.clone(), bd.clone())
    assert abs(theta.item() - theta_svd.item()) <= max(1e-8, 10 * cond * 2.2e-16) * abs(want)


def test_cholesky_pencil_reports_an_indefinite_right_hand_side():
    from tnpy_b200 import _cuda

    rng = np.random.default_rng(0)
    a = rng.standard_normal((300, 300))
    b = np.eye(300)
    b[17, 17] = -1.0
    with pytest.raises(RuntimeError, match="positive definite"):
        _cuda.geig_chol_lowest(dev(a + a.T), dev(b))


def test_shift_invert_through_the_cholesky_pencil_beyond_the_dense_limit():
    """n = 12, bond_dim 64 (exact: the state space is 4096-dimensional): the mid-chain sites have 4096 unknowns,
    beyond dense_pencil_dim, and go through tnpy_geig_chol_lowest.  Same anchors as the reference's test
    (tests/test_finite_dmrg.py:56-72): nearest eigenvalue below the offset, the restored state is the ED
    eigenvector, the energy is <H> on it -- all atol 1e-6."""
    from tnpy_b200.finite_dmrg import ShiftInvertDMRG
    from tnpy_b200.model import RandomHeisenberg

    n, h, seed, offset = 12, 10.5, 2022, 0.1
    model = RandomHeisenberg(n=n, h=h, seed=seed)
    evals, evecs = np.linalg.eigh(oracle.full_hamiltonian(model.mpo.arrays))
    idx = np.where(evals < offset)[0].max()
    shifted = RandomHeisenberg(n=n, h=h, seed=seed, offset=offset)
    sidmrg = ShiftInvertDMRG(shifted.mpo, bond_dim=2**6, offset=offset, seed=1)
    energies = sidmrg.run(tol=1e-8)
    sizes = [st.get("n_matvec", 0) for st in sidmrg.solver_stats]
    assert all(st["dense"] for st in sidmrg.solver_stats) and max(sizes) > 0  # some sites went through Lanczos
    np.testing.assert_allclose(energies[-1], evals[idx], atol=1e-6)
    vec = sidmrg.restored_mps.to_dense()
    if not np.allclose(vec, evecs[:, idx], atol=1e-6):
        np.testing.assert_allclose(-vec, evecs[:, idx], atol=1e-6)
    np.testing.assert_allclose(energies[-1], sidmrg.measurements.expectation_value(model.mpo), atol=1e-6)
