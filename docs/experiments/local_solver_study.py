"""CPU study of the local eigensolver on genuine cold-sweep H_eff problems from the oracle (NumPy, dense H_eff):

  (1) thick-restart Lanczos, classical Gram-Schmidt with the second pass decided by ||w'|| >= eta ||w|| for several
      eta -- matvec counts, share of steps that needed the second pass, orthogonality of the final basis;
  (2) Davidson (GD+k: restart keeps the Ritz vector and the previous one) with and without the diagonal
      preconditioner (the diagonal of H_eff is cheap: sum_ab L[l,a,l] W[a,b,p,p] R[r,b,r]).

    python docs/experiments/local_solver_study.py [n] [chi]

Results (n=24, chi=64, sites 8..12 of the first sweep, tol 1e-8) are quoted in DESIGN.md section 3.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tnpy_oracle as oracle  # noqa: E402


def trlan(H, v0, ncv, keep, tol, eta, max_matvec=20000):
    n = len(v0)
    V = np.zeros((ncv + 1, n))
    V[0] = v0 / np.linalg.norm(v0)
    T = np.zeros((ncv, ncv))
    j = matvecs = second = 0
    worst_orth = 0.0
    while True:
        w = H @ V[j]
        matvecs += 1
        wn = np.linalg.norm(w)
        h = V[:j + 1] @ w
        w = w - V[:j + 1].T @ h
        col = h.copy()
        if np.linalg.norm(w) < eta * wn:
            h2 = V[:j + 1] @ w
            w = w - V[:j + 1].T @ h2
            col += h2
            second += 1
        T[:j + 1, j] = T[j, :j + 1] = col
        beta = np.linalg.norm(w)
        V[j + 1] = w / beta
        m = j + 1
        worst_orth = max(worst_orth, np.abs(V[:m] @ V[m]).max())
        theta, S = np.linalg.eigh(T[:m, :m])
        resid, anorm = abs(beta * S[m - 1, 0]), np.abs(theta).max()
        if resid <= tol * anorm or matvecs >= max_matvec:
            psi = S[:, 0] @ V[:m]
            psi /= np.linalg.norm(psi)
            return matvecs, psi @ H @ psi, np.linalg.norm(H @ psi - (psi @ H @ psi) * psi) / anorm, second / matvecs, worst_orth
        if m == ncv:
            k = min(keep, m)
            V[:k] = S[:, :k].T @ V[:m]
            V[k] = V[m]
            T[:] = 0.0
            T[np.arange(k), np.arange(k)] = theta[:k]
            j = k
        else:
            j += 1


def trlan_local(H, v0, ncv, keep, tol, max_matvec=20000):
    """What csrc/lanczos.cu does: a local Gram-Schmidt pass against (v_{j-1}, v_j) -- the whole basis in the first
    step after a restart -- then one pass against the whole basis, a third only if that one cancelled (DGKS)."""
    n = len(v0)
    V = np.zeros((ncv + 1, n))
    V[0] = v0 / np.linalg.norm(v0)
    T = np.zeros((ncv, ncv))
    j = matvecs = third = 0
    whole = 0
    worst_orth = 0.0
    while True:
        w = H @ V[j]
        matvecs += 1
        lo = 0 if j == whole else max(j - 1, 0)
        col = np.zeros(j + 1)
        hl = V[lo:j + 1] @ w
        w = w - V[lo:j + 1].T @ hl
        col[lo:] += hl
        wn = np.linalg.norm(w)
        h = V[:j + 1] @ w
        w = w - V[:j + 1].T @ h
        col += h
        if np.linalg.norm(w) < np.linalg.norm(h):
            h2 = V[:j + 1] @ w
            w = w - V[:j + 1].T @ h2
            col += h2
            third += 1
        T[:j + 1, j] = T[j, :j + 1] = col
        beta = np.linalg.norm(w)
        V[j + 1] = w / beta
        m = j + 1
        worst_orth = max(worst_orth, np.abs(V[:m] @ V[m]).max())
        theta, S = np.linalg.eigh(T[:m, :m])
        resid, anorm = abs(beta * S[m - 1, 0]), np.abs(theta).max()
        if resid <= tol * anorm or matvecs >= max_matvec:
            psi = S[:, 0] @ V[:m]
            psi /= np.linalg.norm(psi)
            return matvecs, psi @ H @ psi, np.linalg.norm(H @ psi - (psi @ H @ psi) * psi) / anorm, third / matvecs, worst_orth
        if m == ncv:
            k = min(keep, m)
            V[:k] = S[:, :k].T @ V[:m]
            V[k] = V[m]
            T[:] = 0.0
            T[np.arange(k), np.arange(k)] = theta[:k]
            j = k
            whole = k
        else:
            j += 1


def davidson(H, v0, ncv, tol, diag=None, max_matvec=20000):
    """GD+k with k = 1: basis restarted to [x, x_prev]; correction t = r / (diag - theta) or t = r."""
    n = len(v0)
    V = [v0 / np.linalg.norm(v0)]
    W = [H @ V[0]]
    matvecs = 1
    x_prev = None
    anorm = 0.0
    while True:
        Vm, Wm = np.array(V), np.array(W)
        G = Vm @ Wm.T
        theta, S = np.linalg.eigh(0.5 * (G + G.T))
        anorm = max(anorm, np.abs(theta).max())
        x = S[:, 0] @ Vm
        r = S[:, 0] @ Wm - theta[0] * x
        if np.linalg.norm(r) <= tol * anorm or matvecs >= max_matvec:
            return matvecs, theta[0], np.linalg.norm(r) / anorm
        if len(V) >= ncv:
            keep = [x] + ([x_prev] if x_prev is not None else [])
            Q, _ = np.linalg.qr(np.array(keep).T)
            V = [q for q in Q.T]
            W = [H @ q for q in V]  # (a real implementation recombines W; matvecs not counted here)
        x_prev = x
        t = r if diag is None else r / np.where(np.abs(diag - theta[0]) > 1e-3, diag - theta[0], 1e-3)
        for _ in range(2):
            t = t - np.array(V).T @ (np.array(V) @ t)
        t /= np.linalg.norm(t)
        V.append(t)
        W.append(H @ t)
        matvecs += 1


def problems(n, chi, sites):
    f = oracle.FiniteDMRG(oracle.xxz_mpo(n, 0.5), chi, mps=oracle.random_mps(n, chi, 2, seed=0))
    out = []
    for site in range(max(sites) + 1):
        if site in sites:
            H = f.env.one_site_full_matrix(site)
            out.append((site, 0.5 * (H + H.T), f.mps[site].reshape(-1).copy()))
        e, psi = f.one_site_solver(site, 1e-8)
        f.env.update_mps(site, np.asarray(psi).reshape(f.mps[site].shape))
        f.perturb_wave_function(site)
        f.env.split_tensor(site, oracle.RIGHTWARD)
        f.env.update(site, oracle.RIGHTWARD)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    chi = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    probs = problems(n, chi, sites=(8, 9, 10, 11, 12))
    print("thick-restart Lanczos (32, 10), tol 1e-8: eta -> total matvecs, second-pass share, worst |V^T v_new|, worst resid/|A|")
    for eta in (1.0, 0.7071, 0.25, 0.1, 1e-2, 1e-3, 0.0):
        tot, share, orth, res = 0, [], 0.0, 0.0
        for _, H, v0 in probs:
            mv, _, rr, sh, wo = trlan(H, v0, 32, 10, 1e-8, eta)
            tot += mv; share.append(sh); orth = max(orth, wo); res = max(res, rr)
        print(f"  eta {eta:7.4f}: {tot:6d} matvecs, second pass in {np.mean(share):5.1%} of steps, orth {orth:.1e}, resid {res:.1e}", flush=True)
    for tol in (1e-8, 1e-12):
        tot, share, orth, res = 0, [], 0.0, 0.0
        for _, H, v0 in probs:
            mv, _, rr, sh, wo = trlan_local(H, v0, 32, 10, tol)
            tot += mv; share.append(sh); orth = max(orth, wo); res = max(res, rr)
        print(f"  local + full pass, tol {tol:.0e}: {tot:6d} matvecs, third pass in {np.mean(share):5.1%} of steps, orth {orth:.1e}, resid {res:.1e}", flush=True)
    print("Davidson GD+1, basis 32: preconditioner -> total matvecs")
    for name in ("none", "diagonal"):
        tot = 0
        for _, H, v0 in probs:
            mv, _, _ = davidson(H, v0, 32, 1e-8, diag=None if name == "none" else np.diag(H).copy())
            tot += mv
        print(f"  {name:9s}: {tot:6d} matvecs", flush=True)


if __name__ == "__main__":
    main()
