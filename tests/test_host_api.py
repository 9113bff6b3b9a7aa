"""Host-side mirror of the reference interface: models / MPO / MPS (bit-exact, no GPU needed)."""
import os
from contextlib import nullcontext as does_not_raise

import numpy as np
import pytest

from oracle import tnpy_oracle as oracle
from tests.test_oracle_golden import H_RH_N2, H_RH_N3, H_XXZ_N2, H_XXZ_N3
from tnpy_b200.matrix_product_state import Direction, MatrixProductState, compressed_bond_dims
from tnpy_b200.model import XXZ, RandomHeisenberg, Thirring, TotalSz, TransverseIsing
from tnpy_b200.operators import FullHamiltonian, MatrixProductOperator, SpinOperators


def test_spin_operators():
    ops = SpinOperators()
    np.testing.assert_array_equal(np.array([[0, 1], [1, 0]]), ops.Sp + ops.Sm)
    np.testing.assert_array_equal(np.array([[0, -1j], [1j, 0]]), -1j * (ops.Sp - ops.Sm))
    Sp, Sm, Sz, I2, O2 = SpinOperators(spin=1)
    np.testing.assert_array_equal(Sz, np.diag([1.0, -1.0]))


@pytest.mark.parametrize(
    "model,expected",
    [
        (lambda: RandomHeisenberg(n=2, h=0), H_RH_N2),
        (lambda: RandomHeisenberg(n=3, h=0), H_RH_N3),
        (lambda: XXZ(n=2, delta=0.5), H_XXZ_N2),
        (lambda: XXZ(n=3, delta=0.5), H_XXZ_N3),
    ],
)
def test_full_hamiltonian_golden(model, expected):
    ham = FullHamiltonian(model().mpo)
    assert ham.n_sites == int(np.log2(expected.shape[0]))
    np.testing.assert_array_equal(ham.matrix, expected)


@pytest.mark.parametrize("penalty", [0, 100.0])
def test_mpo_bit_exact_against_oracle(penalty):
    """MPO construction must be bit-exact (north_star)."""
    pairs = [
        (XXZ(7, 0.5).mpo, oracle.xxz_mpo(7, 0.5)),
        (Thirring(7, 0.5, 1.0, penalty, 0).mpo, oracle.thirring_mpo(7, 0.5, 1.0, penalty, 0)),
        (
            RandomHeisenberg(7, 10.5, penalty=penalty, seed=2022, offset=0.2).mpo,
            oracle.random_heisenberg_mpo(7, 10.5, penalty=penalty, seed=2022, offset=0.2),
        ),
    ]
    for mpo, ref in pairs:
        assert mpo.nsites == len(ref)
        for a, b in zip(mpo.arrays, ref):
            assert a.shape == b.shape and np.array_equal(a, b)
    w = 6 if penalty else 5
    assert Thirring(7, 0.5, 1.0, penalty, 0).mpo[3].shape == (w, w, 2, 2)


@pytest.mark.parametrize("model", [RandomHeisenberg(n=4, h=0), RandomHeisenberg(n=4, h=0.5, seed=1)])
def test_square(model):
    bilayer = model.mpo.square()
    assert [bilayer[i].shape for i in range(4)] == [(25, 2, 2), (25, 25, 2, 2), (25, 25, 2, 2), (25, 2, 2)]
    ham = FullHamiltonian(model.mpo).matrix
    np.testing.assert_allclose(ham @ ham, FullHamiltonian(bilayer).matrix, atol=1e-12)


@pytest.mark.parametrize("n", [2, 4, 6])
@pytest.mark.parametrize("h", [0, 0.5, 1])
def test_multiply_scalar(n, h):
    mpo = RandomHeisenberg(n=n, h=h, seed=0).mpo
    np.testing.assert_array_equal(-1 * FullHamiltonian(mpo).matrix, FullHamiltonian(-1 * mpo).matrix)


def test_mixed_phys_dims_rejected():
    with pytest.raises(ValueError):
        MatrixProductOperator([np.zeros((3, 2, 2)), np.zeros((3, 3, 3, 3)), np.zeros((3, 2, 2))])


def test_other_models_build():
    assert TransverseIsing(5, 1.0, 0.5).mpo[2].shape == (3, 3, 2, 2)
    tz = TotalSz(4)
    np.testing.assert_array_equal(
        FullHamiltonian(tz.mpo).matrix, np.diag([sum(0.5 if not (i >> k) & 1 else -0.5 for k in range(4)) for i in range(16)])
    )
    sub = FullHamiltonian(tz.subsystem_mpo(1)).matrix
    np.testing.assert_array_equal(sub, np.kron(FullHamiltonian(TotalSz(2).mpo).matrix, np.eye(4)))
    with pytest.raises(ValueError):
        tz.subsystem_mpo(9)


@pytest.mark.parametrize("n", [6, 8])
@pytest.mark.parametrize("bond_dim", [2, 4, 6])
@pytest.mark.parametrize("phys_dim", [2, 4])
def test_random_mps(n, bond_dim, phys_dim):
    mps = MatrixProductState.random(n=n, bond_dim=bond_dim, phys_dim=phys_dim, seed=1)
    chi = compressed_bond_dims(n, bond_dim, phys_dim)
    for site, t in enumerate(mps):
        want = (phys_dim, chi[0]) if site == 0 else ((chi[-1], phys_dim) if site == n - 1 else (chi[site - 1], phys_dim, chi[site]))
        assert t.shape == want
    np.testing.assert_allclose(mps @ mps.conj(mangle_inner=True), 1, atol=1e-12)
    assert mps.phys_dim == phys_dim and mps.n_sites == n and mps.bond_dim == max(chi)
    ref = oracle.random_mps(n, bond_dim, phys_dim, seed=1)
    assert all(np.array_equal(a, b) for a, b in zip(mps.arrays, ref))


@pytest.mark.parametrize(
    "filename, expectation",
    [("test.npz", does_not_raise()), ("test.txt", pytest.raises(ValueError))],
)
def test_save_load(filename, expectation, tmp_path):
    mps = MatrixProductState.random(n=12, bond_dim=6, phys_dim=2, seed=2)
    path = str(tmp_path / filename)
    with expectation:
        mps.save(path)
        assert os.path.isfile(path)
        back = MatrixProductState.load(path)
        assert all(np.array_equal(a, b) for a, b in zip(mps.arrays, back.arrays))


def test_quimb_default_layout_accepted():
    mps = MatrixProductState.random(n=6, bond_dim=4, phys_dim=2, seed=0)
    lrp = [mps.arrays[0].T] + [np.transpose(a, (0, 2, 1)) for a in mps.arrays[1:-1]] + [mps.arrays[-1]]
    back = MatrixProductState(lrp, shape="lrp")
    assert all(np.array_equal(a, b) for a, b in zip(mps.arrays, back.arrays))


def test_direction_enum():
    assert Direction.RIGHTWARD.value == 1 and Direction.LEFTWARD.value == -1


def test_bond_dims_of_a_non_uniform_mpo():
    """Bond i is the right bond of site i (site 0 carries no left bond)."""
    rng = np.random.default_rng(0)
    arrays = [rng.standard_normal((2, 2, 2)), rng.standard_normal((2, 3, 2, 2)), rng.standard_normal((3, 4, 2, 2)),
              rng.standard_normal((4, 2, 2))]
    assert MatrixProductOperator(arrays).bond_dims() == [2, 3, 4]
    assert XXZ(n=5, delta=0.5).mpo.bond_dims() == [5, 5, 5, 5]
    assert XXZ(n=5, delta=0.5).mpo.square().bond_dims() == [25, 25, 25, 25]


def test_reference_import_paths_resolve_to_this_implementation():
    """A user of tanlin2013/tnpy keeps their imports (README.md:97-101, scripts/thirring_fdmrg.py:1-2)."""
    import tnpy
    import tnpy_b200.finite_dmrg as impl
    from tnpy.finite_dmrg import FiniteDMRG, Metric, ShiftInvertDMRG
    from tnpy.linalg import eigh, eigshmv, svd  # noqa: F401
    from tnpy.matrix_product_state import Direction as D2, Environment, MatrixProductState as M2  # noqa: F401
    from tnpy.model import RandomHeisenberg as RH2, Thirring as T2, XXZ as X2
    from tnpy.model.thirring import Thirring as T3
    from tnpy.operators import MatrixProductOperator as MPO2

    assert FiniteDMRG is impl.FiniteDMRG and ShiftInvertDMRG is impl.ShiftInvertDMRG and Metric is impl.Metric
    assert X2 is XXZ and T2 is Thirring and T3 is Thirring and RH2 is RandomHeisenberg
    assert M2 is MatrixProductState and D2 is Direction and MPO2 is MatrixProductOperator
    assert tnpy.logger.name == "tnpy"


def test_to_quimb_round_trip():
    """The returned MPS converts to the reference's quimb type when quimb is installed (it is not a dependency)."""
    pytest.importorskip("quimb")
    mps = MatrixProductState.random(n=6, bond_dim=4, phys_dim=2, seed=1)
    q = mps.to_quimb()
    assert q.L == 6 and abs((q.H @ q) - mps.overlap(mps)) < 1e-12
    for i in range(6):
        np.testing.assert_array_equal(np.asarray(q[i].data).reshape(mps[i].shape), mps[i].data)
