"""End-to-end parity of the GPU FiniteDMRG against the CPU oracle and the reference's own anchors."""
import numpy as np
import pytest

from oracle import tnpy_oracle as oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def normalised_overlap(a, b):
    return abs(oracle.mps_overlap(a, b)) / np.sqrt(oracle.mps_overlap(a, a) * oracle.mps_overlap(b, b))


def test_reference_anchor_xxz_n10():
    """tests/test_finite_dmrg.py:22-23 -- fDMRG energy == dense ED, atol 1e-8 (random start)."""
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.model import XXZ

    model = XXZ(n=10, delta=0.5)
    e_ed = oracle.exact_ground_energy(model.mpo.arrays)
    energies = FiniteDMRG(model.mpo, bond_dim=2**5).run(tol=1e-8)
    np.testing.assert_allclose(energies[-1], e_ed, atol=1e-8)


def test_readme_spelling():
    """README.md:97-101 -- FiniteDMRG(mpo=..., chi=...).update(tol=...)."""
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.model import XXZ

    model = XXZ(n=8, delta=0.5)
    fdmrg = FiniteDMRG(mpo=model.mpo, chi=16, seed=0)
    with pytest.raises(RuntimeError):
        fdmrg.measurements
    energies = fdmrg.update(tol=1e-8)
    np.testing.assert_allclose(energies[-1], oracle.exact_ground_energy(model.mpo.arrays), atol=1e-8)
    assert abs(fdmrg.measurements.expectation_value(model.mpo) / fdmrg.measurements.expectation_value() - energies[-1]) < 1e-8


@pytest.mark.parametrize(
    "name,n,chi",
    [("xxz", 10, 32), ("xxz", 14, 20), ("thirring", 12, 16), ("thirring_script", 8, 8), ("random_heisenberg", 12, 16)],
)
def test_parity_with_oracle(name, n, chi):
    """Same model, chi and initial MPS on both sides (north_star): energy 1e-10 relative,
    per-bond singular values 1e-9, |<psi_ref|psi_gpu>| > 1 - 1e-8 on normalised states."""
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import MatrixProductState
    from tnpy_b200 import model as models

    mdl = {
        "xxz": lambda: models.XXZ(n=n, delta=0.5),
        "thirring": lambda: models.Thirring(n=n, delta=0.5, ma=1.0, penalty=1.0, s_target=0),
        # scripts/thirring_fdmrg.py:10 parameters; chi=8 keeps every site on the dense (N < 200) branch
        "thirring_script": lambda: models.Thirring(n=n, delta=0.5, ma=1.0, penalty=100.0, s_target=0),
        "random_heisenberg": lambda: models.RandomHeisenberg(n=n, h=1.0, seed=2022),
    }[name]()
    init = oracle.random_mps(n, chi, 2, seed=11)
    # Same number of sweeps on both sides (SURVEY 7 "eigensolver parity"): the sweep-level stopping
    # rule |dE| < tol is made unreachable so neither side stops a sweep earlier than the other; the
    # local solves run at tol * ||A|| residual on the GPU and exactly (dense) in the oracle.
    tol, sweeps = 1e-13, 6
    ref = oracle.FiniteDMRG(mdl.mpo.arrays, chi, mps=[a.copy() for a in init], exact_local_solver=True)
    e_ref = ref.run(tol=tol, max_sweep=sweeps)
    gpu = FiniteDMRG(mdl.mpo, bond_dim=chi, mps=MatrixProductState([a.copy() for a in init]))
    e_gpu = gpu.run(tol=tol, max_sweep=sweeps)
    assert len(e_gpu) == len(e_ref)
    for a, b in zip(e_gpu, e_ref):
        assert abs(a - b) <= 1e-10 * abs(b), (e_gpu, e_ref)
    assert normalised_overlap(ref.mps, gpu.mps.arrays) > 1 - 1e-8
    # the un-normalised state carries the reference's (1 + alpha E) factors
    assert abs(oracle.mps_overlap(gpu.mps.arrays, gpu.mps.arrays) / oracle.mps_overlap(ref.mps, ref.mps) - 1) < 1e-8
    sv_gpu = gpu.bond_singular_values
    for bond, s_ref in ref.bond_singular_values.items():
        assert np.abs(sv_gpu[bond] - s_ref).max() < 1e-9, bond
    # variance bookkeeping of run() (finite_dmrg.py:248)
    assert abs(gpu._variances[-1] - ref.variances[-1]) < 1e-8 * max(1.0, abs(e_ref[-1])) ** 2


def test_environment_invariants():
    """Kernel-level pins the reference lacks: canonical identity channels, symmetry of
    H_eff, <psi|H_eff|psi> == <H>."""
    from tnpy_b200.matrix_product_state import Environment, MatrixProductState
    from tnpy_b200.model import XXZ

    n, chi = 12, 16
    model = XXZ(n=n, delta=0.5)
    mps = MatrixProductState.random(n, chi, 2, seed=5)
    env = Environment(model.mpo, mps)
    for site in (1, 5, n - 2):
        R = env.right[site].data
        assert np.abs(R[:, -1, :] - np.eye(R.shape[0])).max() < 1e-12  # right-canonical => identity channel
    site = 0
    h = env.one_site_full_matrix(site)
    assert np.abs(h - h.T).max() < 1e-12
    x = mps[site].data.reshape(-1)
    full = oracle.mps_expectation(mps.arrays, model.mpo.arrays)
    assert abs(x @ h @ x - full) < 1e-12
    op = env.one_site_matvec(site)
    y = op.matvec(x)
    assert np.abs(y - h.T @ x).max() < 1e-12
    assert abs(env.expectation() - full) < 1e-12
    assert abs(env.variance() - (oracle.mps_expectation(mps.arrays, oracle.mpo_square(model.mpo.arrays)) - full**2)) < 1e-10


def test_split_tensor_invariance():
    """tests/test_matrix_product_state.py:83-89 -- A[site].A[site+1] is preserved, atol 1e-12."""
    from tnpy_b200.matrix_product_state import Direction, MatrixProductState

    mps = MatrixProductState.random(n=8, bond_dim=10, phys_dim=2, seed=3)
    for site in (2, 3, 4, 6):
        before = np.tensordot(mps.three_leg(site), mps.three_leg(site + 1), axes=(2, 0))
        mps.split_tensor(site, direction=Direction.RIGHTWARD)
        after = np.tensordot(mps.three_leg(site), mps.three_leg(site + 1), axes=(2, 0))
        np.testing.assert_allclose(before, after, atol=1e-12)
        assert mps[site].tags == {f"I{site}"}
    with pytest.raises(KeyError):
        mps.split_tensor(2, direction="sideways")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from tnpy_b200 import _cuda

    monkeypatch.setattr(_cuda, "_lib", None)
    monkeypatch.setattr(_cuda, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cuda.load()


def test_user_mps_can_be_canonicalised_on_device():
    """a non-canonical user MPS (plain Gaussian tensors) is right-canonicalised by device
    SVD splits; the state itself is unchanged and the run reaches the ED energy."""
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import Environment, MatrixProductState
    from tnpy_b200.model import XXZ

    n, chi = 8, 16
    rng = np.random.default_rng(0)
    dims = [1, 2, 4, 8, 16, 8, 4, 2, 1]
    arrays = [rng.standard_normal((dims[i], 2, dims[i + 1])) for i in range(n)]
    arrays[0], arrays[-1] = arrays[0][0], arrays[-1][:, :, 0]
    mps = MatrixProductState(arrays)
    before = mps.to_dense()
    model = XXZ(n=n, delta=0.5)
    env = Environment(model.mpo, mps.copy(), canonicalize=True)
    after = env.mps
    np.testing.assert_allclose(after.to_dense(), before, atol=1e-10 * np.abs(before).max())
    for site in range(1, n):
        a = after.three_leg(site)
        m = a.reshape(a.shape[0], -1)
        np.testing.assert_allclose(m @ m.T, np.eye(a.shape[0]), atol=1e-12)
    energies = FiniteDMRG(model.mpo, bond_dim=chi, mps=mps, canonicalize=True).run(tol=1e-9)
    np.testing.assert_allclose(energies[-1], oracle.exact_ground_energy(model.mpo.arrays), atol=1e-8)


def test_odd_bond_dimension():
    """chi = 15: odd leading dimensions take the generic GEMM / scalar vector paths end to end."""
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.model import XXZ

    model = XXZ(n=8, delta=0.5)
    energies = FiniteDMRG(model.mpo, bond_dim=15, seed=4).run(tol=1e-9)
    e_ed = oracle.exact_ground_energy(model.mpo.arrays)
    assert energies[-1] >= e_ed - 1e-10 and energies[-1] - e_ed < 1e-5


@pytest.mark.parametrize("name,n,chi", [("xxz", 10, 32), ("random_heisenberg", 12, 16)])
def test_qr_split_sweeps_match_svd_split_and_oracle(name, n, chi):
    """The deferred-SVD gauge (Cholesky-QR splits, forced on for every bond >= 4 here) changes nothing
    observable: per-sweep energies match the reference-literal SVD gauge and the oracle to 1e-10 relative,
    bond spectra to 1e-9, the final states overlap to 1e-8 -- and the QR path really ran."""
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import MatrixProductState
    from tnpy_b200 import model as models

    mdl = {"xxz": lambda: models.XXZ(n=n, delta=0.5),
           "random_heisenberg": lambda: models.RandomHeisenberg(n=n, h=1.0, seed=2022)}[name]()
    init = oracle.random_mps(n, chi, 2, seed=11)
    tol, sweeps = 1e-13, 6
    ref = oracle.FiniteDMRG(mdl.mpo.arrays, chi, mps=[a.copy() for a in init], exact_local_solver=True)
    e_ref = ref.run(tol=tol, max_sweep=sweeps)
    runs = {}
    for split in ("qr", "svd"):
        f = FiniteDMRG(mdl.mpo, bond_dim=chi, mps=MatrixProductState([a.copy() for a in init]), split=split)
        f.environment.qr_min_bond = 4
        runs[split] = (f, f.run(tol=tol, max_sweep=sweeps))
    qr, e_qr = runs["qr"]
    svd, e_svd = runs["svd"]
    counts = qr.environment.split_counts
    assert counts["qr"] > 0 and counts["qr"] + counts["qr_shifted"] > counts["svd"]
    assert svd.environment.split_counts["qr"] + svd.environment.split_counts["qr_shifted"] == 0
    for a, b, c in zip(e_qr, e_svd, e_ref):
        assert abs(a - b) <= 1e-10 * abs(b) and abs(a - c) <= 1e-10 * abs(c), (e_qr, e_svd, e_ref)
    assert normalised_overlap(ref.mps, qr.mps.arrays) > 1 - 1e-8
    assert normalised_overlap(svd.mps.arrays, qr.mps.arrays) > 1 - 1e-8
    sv_qr = qr.bond_singular_values
    for bond, s_ref in ref.bond_singular_values.items():
        assert np.abs(sv_qr[bond] - s_ref).max() < 1e-9, bond


@pytest.mark.parametrize("split", ["qr", "svd"])
def test_environment_split_preserves_state_at_every_site(split):
    """Environment.split_tensor leaves the state untouched at every site and in both directions -- ramp sites
    included, where the site tensor is a *square* matrix (l == d r or l d == r) and the order of the two
    factors of the Cholesky-QR split is a choice the caller has to make."""
    from tnpy_b200.matrix_product_state import Direction, Environment, MatrixProductState
    from tnpy_b200.model import XXZ

    n, chi = 10, 16  # bonds 2 4 8 16 16 16 8 4 2
    mps = MatrixProductState.random(n=n, bond_dim=chi, phys_dim=2, seed=4)
    dense = mps.to_dense()
    env = Environment(XXZ(n=n, delta=0.5).mpo, mps, split=split, qr_min_bond=2)
    for direction, sites in ((Direction.RIGHTWARD, range(0, n - 1)), (Direction.LEFTWARD, range(n - 1, 0, -1))):
        for site in sites:
            env.split_tensor(site, direction)
            now = env.mps.to_dense()
            assert np.abs(now - dense).max() < 1e-12 * np.abs(dense).max(), (split, direction, site)
    if split == "qr":
        assert env.split_counts["qr"] + env.split_counts["qr_shifted"] >= 2 * (n - 1) - 4


def test_perturbation_from_the_solver_image_equals_a_fresh_matvec():
    """finite_dmrg.py:116-141: psi += alpha H_eff psi.  Inside a sweep the H_eff psi comes from the eigensolve
    that produced psi (no second matvec); the result must equal the reference-literal route to rounding, and a
    site tensor rewritten in between must not pick up a stale image."""
    import torch

    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import MatrixProductState
    from tnpy_b200.model import XXZ

    n, chi, site = 12, 32, 6
    f = FiniteDMRG(XXZ(n=n, delta=0.5).mpo, bond_dim=chi, mps=MatrixProductState.random(n, chi, 2, seed=2))
    env = f.environment
    f._solve_on_device(site, 1e-9)
    psi = env.device_tensor(site).clone()
    expected = psi + 1e-5 * env.one_site_matvec(site).apply_device(psi)
    assert env._image is not None and env._image[0] == site
    f.perturb_wave_function(site)
    assert env._image is None  # consumed
    assert float((env.device_tensor(site) - expected).abs().max()) <= 1e-14 * float(expected.abs().max())
    # stale image: the tensor is replaced after the solve
    f._solve_on_device(site, 1e-9)
    other = torch.randn_like(psi)
    env.update_mps(site, other)
    assert env._image is None
    f.perturb_wave_function(site)
    expected = other + 1e-5 * env.one_site_matvec(site).apply_device(other)
    assert float((env.device_tensor(site) - expected).abs().max()) <= 1e-14 * float(expected.abs().max())


def test_heff_operator_matmat_columns_do_not_alias():
    """SciPy's default LinearOperator.matmat stacks matvec results; with the reference's solver (primme calls matmat)
    every column must be its own memory.  matvec returns fresh arrays by default (the zero-copy pinned views are an
    explicit opt-in for timing loops), matmat moves the block once."""
    from tnpy_b200.matrix_product_state import Environment, MatrixProductState
    from tnpy_b200.model import XXZ

    n, chi = 10, 16
    env = Environment(XXZ(n=n, delta=0.5).mpo, MatrixProductState(oracle.random_mps(n, chi, 2, seed=2)))
    site = 5
    op = env.one_site_matvec(site)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((op.shape[0], 4))
    want = np.column_stack([
        op.apply_device(torch.from_numpy(np.ascontiguousarray(X[:, i])).cuda()).reshape(-1).cpu().numpy() for i in range(4)
    ])
    np.testing.assert_allclose(op.matmat(X), want, rtol=0, atol=1e-13 * np.abs(want).max())
    np.testing.assert_allclose(op @ X, want, rtol=0, atol=1e-13 * np.abs(want).max())
    cols = [op.matvec(X[:, i]) for i in range(4)]  # kept results stay valid
    np.testing.assert_allclose(np.column_stack(cols), want, rtol=0, atol=1e-13 * np.abs(want).max())
    # against the oracle's H_eff at the same site
    ref_env = oracle.Environment(XXZ(n=n, delta=0.5).mpo.arrays, oracle.random_mps(n, chi, 2, seed=2))
    y = ref_env.matvec(site, X[:, 0])
    np.testing.assert_allclose(cols[0], np.asarray(y).reshape(-1), rtol=0, atol=1e-12 * np.abs(want).max())
    # the operator notices an environment rewritten in place
    env.update_right(site)
    np.testing.assert_allclose(op.matvec(X[:, 1]), want[:, 1], rtol=0, atol=1e-13 * np.abs(want).max())
