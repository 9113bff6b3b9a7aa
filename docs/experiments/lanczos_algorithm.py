"""NumPy restatement of the on-device eigensolver's iteration (csrc/lanczos.cu: thick-restart Lanczos, classical
Gram-Schmidt twice, explicit projected matrix, stop at ||r|| <= tol * max|Ritz|) on genuine first-sweep local
problems from the oracle: the restarted iteration needs within a few percent of the matvecs of unrestarted
Lanczos (the optimum for a Krylov method), and the H psi the solver reports from its own recurrence
(tnpy_eig_lowest_image) is the true H psi.  CPU only; the kernels are tested under -m gpu."""
import numpy as np

from oracle import tnpy_oracle as oracle


def thick_restart_lanczos(H, v0, ncv, keep, tol, max_matvec=5000):
    n = len(v0)
    V = np.zeros((ncv + 1, n))
    V[0] = v0 / np.linalg.norm(v0)
    T = np.zeros((ncv, ncv))
    j = matvecs = 0
    while True:
        w = H @ V[j]
        matvecs += 1
        h = V[:j + 1] @ w
        w = w - V[:j + 1].T @ h
        h2 = V[:j + 1] @ w
        w = w - V[:j + 1].T @ h2
        T[:j + 1, j] = T[j, :j + 1] = h + h2
        beta = np.linalg.norm(w)
        V[j + 1] = w / beta
        m = j + 1
        theta, S = np.linalg.eigh(T[:m, :m])
        resid, anorm = abs(beta * S[m - 1, 0]), np.abs(theta).max()
        if resid <= tol * anorm or matvecs >= max_matvec or m >= n:
            psi = S[:, 0] @ V[:m]
            image = theta[0] * psi + beta * S[m - 1, 0] * V[m]  # H V_m = V_m T + beta v_{m+1} e_m^T
            return matvecs, theta[0], psi, image
        if m == ncv:
            k = min(keep, m)
            V[:k] = S[:, :k].T @ V[:m]
            V[k] = V[m]
            T[:] = 0.0
            T[np.arange(k), np.arange(k)] = theta[:k]
            j = k
        else:
            j += 1


def cold_sweep_problems(n=20, chi=48, sites=(6, 8, 10)):
    mpo = oracle.xxz_mpo(n, 0.5)
    f = oracle.FiniteDMRG(mpo, chi, mps=oracle.random_mps(n, chi, 2, seed=0))
    out = []
    for site in range(max(sites) + 1):
        if site in sites:
            H = f.env.one_site_full_matrix(site)
            out.append((0.5 * (H + H.T), f.mps[site].reshape(-1).copy()))
        e, psi = f.one_site_solver(site, 1e-8)
        f.env.update_mps(site, np.asarray(psi).reshape(f.mps[site].shape))
        f.perturb_wave_function(site)
        f.env.split_tensor(site, oracle.RIGHTWARD)
        f.env.update(site, oracle.RIGHTWARD)
    return out


def test_thick_restart_is_within_a_few_percent_of_unrestarted_lanczos_and_image_is_exact():
    problems = cold_sweep_problems()
    restarted = unrestarted = 0
    for H, v0 in problems:
        mv, theta, psi, image = thick_restart_lanczos(H, v0, ncv=32, keep=10, tol=1e-8)  # pick_sizes() defaults
        mv_full, theta_full, _, _ = thick_restart_lanczos(H, v0, ncv=len(v0) - 1 if len(v0) < 400 else 400, keep=0, tol=1e-8)
        restarted += mv
        unrestarted += mv_full
        anorm = np.abs(np.linalg.eigvalsh(H)).max()
        assert abs(theta - theta_full) <= 1e-12 * anorm
        assert np.linalg.norm(H @ psi - theta * psi) <= 1e-8 * anorm * 1.01
        assert np.abs(image - H @ psi).max() <= 1e-12 * anorm
    assert restarted <= 1.05 * unrestarted, (restarted, unrestarted)
