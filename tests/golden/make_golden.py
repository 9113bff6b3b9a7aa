"""Generate the committed golden fixtures of tests/golden/ from the CPU oracle.

    python tests/golden/make_golden.py

The reference itself cannot be imported in the build container (quimb / primme / tensornetwork are not
installable), so these vectors come from ``oracle/tnpy_oracle.py`` -- the NumPy restatement
that ``tests/test_oracle_golden.py`` pins to the literals of the reference's own tests.  They freeze the
oracle (a later edit that changes its arithmetic fails ``tests/test_golden_fixtures.py`` on CPU) and give the
GPU parity tests inputs / outputs that do not depend on recomputing the oracle at test time.

Fixtures (float64, NumPy .npz):
  heff_cases.npz      H_eff matvec y = heff_apply(L, W, R, x) for bulk and edge shapes, three models' W
  env_cases.npz       update_left / update_right outputs for the same operands
  dmrg_xxz_n10_chi16.npz, dmrg_thirring_n10_chi12.npz, dmrg_rh_n10_chi16.npz
                      initial MPS, per-sweep energies (6 sweeps, exact local solves), final bond spectra, ED energy
  config1_xxz_n100_chi60.npz
                      BASELINE.json configs[0] (the reference's README example) run to convergence by the oracle at
                      tol 1e-8 from random_mps(seed=0): per-sweep energies (the initial state is regenerated from the
                      seed by the test, it is not stored)
  config2_thirring_n100_chi256.npz
                      BASELINE.json configs[1] (scripts/thirring_fdmrg.py's model -- delta 0.5, ma 1.0, penalty 100,
                      s_target 0 -- at n=100, chi=256): four sweeps from random_mps(seed=0) with every local solve
                      converged to 1e-12 ||A|| (with the penalty term ||A|| ~ 2e3, so the script's own tol of 1e-8
                      leaves 1e-5 of slack per local solve and two solvers' trajectories cannot be compared):
                      per-sweep energies, every bond spectrum of the last two sweeps, the <Sz_i> profile and the
                      squared norm of the final state.  About an hour of host BLAS; generate it alone with
                      ``python tests/golden/make_golden.py config2``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tnpy_oracle as oracle  # noqa: E402


def heff_and_env_cases():
    rng = np.random.default_rng(20261017)
    mpos = {
        "xxz": oracle.xxz_mpo(6, 0.5),
        "thirring": oracle.thirring_mpo(6, 0.5, 1.0, 100.0, 0),
        "rh": oracle.random_heisenberg_mpo(6, 1.0, seed=2022),
    }
    heff, env = {}, {}
    shapes = [(8, 12), (16, 16), (33, 17), (1, 4), (4, 1)]  # (l, r); l == 1 / r == 1 are the chain ends
    for name, mpo in mpos.items():
        for l, r in shapes:
            site = 0 if l == 1 else (len(mpo) - 1 if r == 1 else 2)
            w4 = oracle._w4(mpo[site], site, len(mpo))
            wl, wr, d = w4.shape[0], w4.shape[1], w4.shape[2]
            L = rng.standard_normal((l, wl, l))
            R = rng.standard_normal((r, wr, r))
            x = rng.standard_normal((l, d, r))
            key = f"{name}_{l}x{r}"
            heff[key + "_L"], heff[key + "_W"], heff[key + "_R"], heff[key + "_x"] = L, w4, R, x
            heff[key + "_y"] = oracle.heff_apply(L, w4, R, x)
            env[key + "_left"] = oracle.env_update_left(L, x, w4)
            env[key + "_right"] = oracle.env_update_right(R, x, w4)
    return heff, env


def dmrg_case(mpo, n, chi, seed):
    init = oracle.random_mps(n, chi, 2, seed=seed)
    f = oracle.FiniteDMRG(mpo, chi, mps=[a.copy() for a in init], exact_local_solver=True)
    energies = f.run(tol=1e-13, max_sweep=6)
    out = {"energies": np.array(energies), "ed_energy": np.array(oracle.exact_ground_energy(mpo)), "n": np.array(n), "chi": np.array(chi)}
    for i, a in enumerate(init):
        out[f"init_{i}"] = a
    for bond, s in f.bond_singular_values.items():
        out[f"spectrum_{bond}"] = np.asarray(s)
    for i, w in enumerate(mpo):
        out[f"mpo_{i}"] = np.asarray(w)
    return out


def config2_case(n=100, chi=256, tol=1e-12, sweeps=4, seed=0):
    mpo = oracle.thirring_mpo(n, 0.5, 1.0, 100.0, 0)
    f = oracle.FiniteDMRG(mpo, chi, mps=oracle.random_mps(n, chi, 2, seed=seed))
    energies, matvecs = [], []
    for k in range(sweeps):
        m0 = f.n_matvec
        energies.append(f.sweep(oracle.RIGHTWARD if k % 2 == 0 else oracle.LEFTWARD, tol=tol))
        matvecs.append(f.n_matvec - m0)
        print("config2 sweep", k + 1, energies[-1], matvecs[-1], flush=True)
    prof, norm2 = oracle.mps_sz_profile(f.mps)
    out = {"energies": np.array(energies), "matvecs": np.array(matvecs), "n": np.array(n), "chi": np.array(chi),
           "tol": np.array(tol), "seed": np.array(seed), "delta": np.array(0.5), "ma": np.array(1.0),
           "penalty": np.array(100.0), "s_target": np.array(0), "sz_profile": prof, "norm2": np.array(norm2)}
    for bond, s in f.bond_singular_values.items():
        out[f"spectrum_{bond}"] = np.asarray(s)
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "config2":
        np.savez_compressed(os.path.join(HERE, "config2_thirring_n100_chi256.npz"), **config2_case())
        return
    heff, env = heff_and_env_cases()
    np.savez_compressed(os.path.join(HERE, "heff_cases.npz"), **heff)
    np.savez_compressed(os.path.join(HERE, "env_cases.npz"), **env)
    np.savez_compressed(os.path.join(HERE, "dmrg_xxz_n10_chi16.npz"), **dmrg_case(oracle.xxz_mpo(10, 0.5), 10, 16, 11))
    np.savez_compressed(os.path.join(HERE, "dmrg_thirring_n10_chi12.npz"),
                        **dmrg_case(oracle.thirring_mpo(10, 0.5, 1.0, 1.0, 0), 10, 12, 5))
    np.savez_compressed(os.path.join(HERE, "dmrg_rh_n10_chi16.npz"),
                        **dmrg_case(oracle.random_heisenberg_mpo(10, 1.0, seed=2022), 10, 16, 7))
    f = oracle.FiniteDMRG(oracle.xxz_mpo(100, 0.5), 60, mps=oracle.random_mps(100, 60, 2, seed=0))
    energies = f.run(tol=1e-8, with_variance=False)
    np.savez_compressed(os.path.join(HERE, "config1_xxz_n100_chi60.npz"), energies=np.array(energies), n=np.array(100),
                        chi=np.array(60), delta=np.array(0.5), seed=np.array(0), tol=np.array(1e-8))
    for name in sorted(os.listdir(HERE)):
        if name.endswith(".npz"):
            print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


if __name__ == "__main__":
    main()
