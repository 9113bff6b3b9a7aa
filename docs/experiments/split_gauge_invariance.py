"""The claim the default bond split rests on: replacing the SVD inside split_tensor by *any*
orthogonal factorisation psi = Q T changes nothing observable -- same energies sweep by sweep, same bond
spectra, same state -- because the two differ by an orthogonal matrix on the bond.  Checked on the CPU with the
oracle's own sweep loop and numpy.linalg.qr in place of the SVD (no CUDA involved)."""
import numpy as np
import pytest

from oracle import tnpy_oracle as oracle


def qr_split_tensor(mps, site, direction):
    """Drop-in for oracle.split_tensor: orthogonalise with a QR factorisation, return the singular values of
    the small factor (what DeferredSpectrum computes on the device)."""
    a = mps[site]
    n = len(mps)
    d = a.shape[0] if site == 0 else a.shape[1]
    if direction == oracle.RIGHTWARD:
        psi = a if site == 0 else a.reshape(d * a.shape[0], -1)
        q, t = np.linalg.qr(psi)
        mps[site] = q.reshape(a.shape)
        mps[site + 1] = np.tensordot(t, mps[site + 1], axes=(1, 0))
    else:
        psi = a if site == n - 1 else a.reshape(-1, d * a.shape[2])
        qt, tt = np.linalg.qr(psi.T)  # psi = tt.T @ qt.T
        mps[site] = qt.T.reshape(a.shape)
        mps[site - 1] = np.tensordot(mps[site - 1], tt.T, axes=(mps[site - 1].ndim - 1, 0))
        t = tt
    return np.linalg.svd(t, compute_uv=False)


@pytest.mark.parametrize("model,n,chi", [("xxz", 12, 16), ("rh", 10, 12)])
def test_qr_gauge_sweeps_equal_svd_gauge_sweeps(monkeypatch, model, n, chi):
    mpo = oracle.xxz_mpo(n, 0.5) if model == "xxz" else oracle.random_heisenberg_mpo(n, 1.0, seed=2022)
    init = oracle.random_mps(n, chi, 2, seed=4)
    runs = {}
    for gauge in ("svd", "qr"):
        if gauge == "qr":
            monkeypatch.setattr(oracle, "split_tensor", qr_split_tensor)
        f = oracle.FiniteDMRG(mpo, chi, mps=[a.copy() for a in init], exact_local_solver=True)
        energies = f.run(tol=1e-13, max_sweep=5, with_variance=False)
        runs[gauge] = (energies, f.mps, dict(f.bond_singular_values))
    (e_svd, mps_svd, sv_svd), (e_qr, mps_qr, sv_qr) = runs["svd"], runs["qr"]
    assert len(e_svd) == len(e_qr)
    for a, b in zip(e_qr, e_svd):
        assert abs(a - b) <= 1e-11 * abs(b)
    ov = abs(oracle.mps_overlap(mps_svd, mps_qr)) / np.sqrt(oracle.mps_overlap(mps_svd, mps_svd) * oracle.mps_overlap(mps_qr, mps_qr))
    assert ov > 1 - 1e-10
    for bond, s in sv_svd.items():
        assert np.abs(np.sort(sv_qr[bond])[::-1] - s).max() < 1e-10
