"""Pin the CPU oracle to every golden vector / known answer the reference's own tests hold for the
fDMRG path.  The literals below are the expected values written in
/root/reference/tests (cited per test); nothing here reads /root/reference at run time."""
import numpy as np
import pytest

from oracle import tnpy_oracle as oracle

# /root/reference/tests/test_operators.py:46-88 (assert_array_equal => exact)
H_RH_N2 = np.array([[0.25, 0, 0, 0], [0, -0.25, 0.5, 0], [0, 0.5, -0.25, 0], [0, 0, 0, 0.25]])
H_RH_N3 = np.array(
    [
        [0.5, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 0.5, 0, 0, 0, 0, 0],
        [0, 0.5, -0.5, 0, 0.5, 0, 0, 0],
        [0, 0, 0, 0, 0, 0.5, 0, 0],
        [0, 0, 0.5, 0, 0, 0, 0, 0],
        [0, 0, 0, 0.5, 0, -0.5, 0.5, 0],
        [0, 0, 0, 0, 0, 0.5, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0.5],
    ]
)
H_XXZ_N2 = np.array([[-0.125, 0, 0, 0], [0, 0.125, -0.5, 0], [0, -0.5, 0.125, 0], [0, 0, 0, -0.125]])
H_XXZ_N3 = np.array(
    [
        [-0.25, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, -0.5, 0, 0, 0, 0, 0],
        [0, -0.5, 0.25, 0, -0.5, 0, 0, 0],
        [0, 0, 0, 0, 0, -0.5, 0, 0],
        [0, 0, -0.5, 0, 0, 0, 0, 0],
        [0, 0, 0, -0.5, 0, 0.25, -0.5, 0],
        [0, 0, 0, 0, 0, -0.5, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, -0.25],
    ]
)


@pytest.mark.parametrize(
    "mpo,expected",
    [
        (lambda: oracle.random_heisenberg_mpo(2, 0), H_RH_N2),
        (lambda: oracle.random_heisenberg_mpo(3, 0), H_RH_N3),
        (lambda: oracle.xxz_mpo(2, 0.5), H_XXZ_N2),
        (lambda: oracle.xxz_mpo(3, 0.5), H_XXZ_N3),
    ],
)
def test_full_hamiltonian_golden(mpo, expected):
    np.testing.assert_array_equal(oracle.full_hamiltonian(mpo()), expected)


def test_spin_half_ops():
    """tests/test_operators.py:9-16"""
    sp, sm, sz, i2, o2 = oracle.spin_operators()
    np.testing.assert_array_equal(np.array([[0, 1], [1, 0]]), sp + sm)
    np.testing.assert_array_equal(np.array([[0, -1j], [1j, 0]]), -1j * (sp - sm))


@pytest.mark.parametrize("h", [0, 0.5])
def test_square(h):
    """tests/test_operators.py:20-28"""
    mpo = oracle.random_heisenberg_mpo(4, h, seed=None if h == 0 else 3)
    sq = oracle.mpo_square(mpo)
    assert [t.shape for t in sq] == [(25, 2, 2), (25, 25, 2, 2), (25, 25, 2, 2), (25, 2, 2)]
    ham = oracle.full_hamiltonian(mpo)
    np.testing.assert_allclose(ham @ ham, oracle.full_hamiltonian(sq), atol=1e-12)


def test_random_heisenberg_seed_and_offset():
    """tests/model/test_random_heisenberg.py:21-26 + the RandomState rule (random_heisenberg.py:57-58)."""
    fields = oracle.random_heisenberg_fields(6, 10.5, 2022)
    np.testing.assert_allclose(fields[:3], [-10.30346911, -0.01978597, -8.11894251], atol=5e-9)
    base = oracle.full_hamiltonian(oracle.random_heisenberg_mpo(6, 0.5, seed=2022))
    shifted = oracle.full_hamiltonian(oracle.random_heisenberg_mpo(6, 0.5, seed=2022, offset=0.5))
    np.testing.assert_allclose(base - 0.5 * np.eye(2**6), shifted, atol=1e-12)


@pytest.mark.parametrize("n", [6, 8])
@pytest.mark.parametrize("bond_dim", [2, 4, 6])
@pytest.mark.parametrize("phys_dim", [2, 4])
def test_random_mps_shapes(n, bond_dim, phys_dim):
    """tests/test_matrix_product_state.py:13-37, :51-53"""
    mps = oracle.random_mps(n, bond_dim, phys_dim, seed=n + bond_dim)
    chi = [min(phys_dim**i, bond_dim) for i in range(1, n // 2)]
    chi += [int(min(phys_dim ** (n / 2), bond_dim))] + chi[::-1]
    for site, t in enumerate(mps):
        if site == 0:
            assert t.shape == (phys_dim, chi[0])
        elif site == n - 1:
            assert t.shape == (chi[-1], phys_dim)
        else:
            assert t.shape == (chi[site - 1], phys_dim, chi[site])
    np.testing.assert_allclose(oracle.mps_overlap(mps, mps), 1, atol=1e-12)


@pytest.mark.parametrize("site", [2, 3, 4, 6])
def test_split_tensor_invariance(site):
    """tests/test_matrix_product_state.py:83-89"""
    mps = oracle.random_mps(8, 10, 2, seed=site)
    before = np.tensordot(mps[site], mps[site + 1], axes=(2, 0))
    oracle.split_tensor(mps, site, oracle.RIGHTWARD)
    np.testing.assert_allclose(before, np.tensordot(mps[site], mps[site + 1], axes=(2, 0)), atol=1e-12)


def test_fdmrg_energy_matches_ed():
    """tests/test_finite_dmrg.py:22-23 -- XXZ n=10 delta=0.5 chi=32: fDMRG == ED, atol 1e-8;
    SURVEY 8c records E0 = -2.546701755161168 for this model."""
    mpo = oracle.xxz_mpo(10, 0.5)
    e_ed = oracle.exact_ground_energy(mpo)
    np.testing.assert_allclose(e_ed, -2.546701755161168, atol=1e-12)
    for exact in (False, True):
        energies = oracle.FiniteDMRG(mpo, 32, seed=1, exact_local_solver=exact).run(tol=1e-8)
        assert len(energies) >= 2  # first convergence check is against nan (finite_dmrg.py:75, :211)
        np.testing.assert_allclose(energies[-1], e_ed, atol=1e-8)


def test_kernel_level_invariants():
    """No reference test pins the contraction itself; these invariants do."""
    n, chi = 10, 12
    mpo = oracle.thirring_mpo(n, 0.5, 1.0, 100.0, 0)
    mps = oracle.random_mps(n, chi, 2, seed=4)
    env = oracle.Environment(mpo, mps)
    for site in range(n - 1):
        r = env.right[site]
        np.testing.assert_allclose(r[:, -1, :], np.eye(r.shape[0]), atol=1e-12)
    full = oracle.mps_expectation(mps, mpo)
    for site in (0, 4, n - 1):
        h = env.one_site_full_matrix(site)
        np.testing.assert_allclose(h, h.T, atol=1e-10 * np.abs(h).max())
        x = mps[site].reshape(-1)
        np.testing.assert_allclose(env.matvec(site, x).reshape(-1), h.T @ x, atol=1e-10 * np.abs(h).max())
        if site == 0:  # right-canonical state: <psi|H_eff|psi> at the centre equals <H>
            np.testing.assert_allclose(x @ h @ x, full, rtol=1e-12)
    dense = oracle.mps_to_dense(mps)
    np.testing.assert_allclose(dense @ oracle.full_hamiltonian(mpo) @ dense, full, rtol=1e-10)


def test_sweep_keeps_bond_dims_and_norm():
    """bond dimensions never change; the MPS norm after a sweep is |1 + alpha E|."""
    mpo = oracle.xxz_mpo(10, 0.5)
    d = oracle.FiniteDMRG(mpo, 32, seed=0)
    dims0 = [a.shape for a in d.mps]
    energies = d.run(tol=1e-8)
    assert [a.shape for a in d.mps] == dims0
    norm = np.sqrt(oracle.mps_overlap(d.mps, d.mps))
    np.testing.assert_allclose(norm, abs(1 + 1e-5 * energies[-1]), rtol=1e-8)
