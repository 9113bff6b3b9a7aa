"""Finite-size single-site DMRG driver (mirrors tnpy/finite_dmrg.py:22-263).

Same public surface as the reference -- ``FiniteDMRG(mpo, bond_dim, block_size=1, mps=None,
exact_solver_dim=200)``, ``run(tol, max_sweep, metric, **kwargs) -> List[float]``, ``sweep``,
``one_site_solver``, ``perturb_wave_function``, ``variance``, the ``bond_dim / mps / n_sites /
phys_dim / measurements`` properties -- plus the README spelling (``chi=`` and ``update(tol=)``,
README.md:97-101).  The sweep keeps every tensor on the GPU: per site it issues the on-device
eigensolve, the perturbation matvec, the SVD split and the environment update through the C ABI and
reads back one status record per Lanczos step (never a vector).
"""
from __future__ import annotations

import time
from datetime import timedelta
from enum import Enum
from functools import partial
from itertools import cycle
from typing import Dict, List, Optional, Tuple

import numpy as np

from tnpy_b200 import _cuda, logger
from tnpy_b200.linalg import eigh, eigshmv
from tnpy_b200.matrix_product_state import (
    Direction,
    Environment,
    MatrixProductState,
    MatrixProductStateMeasurements,
)
from tnpy_b200.operators import MatrixProductOperator


class Metric(Enum):
    ENERGY = 1
    VARIANCE = 2


class FiniteDMRG:
    def __init__(
        self,
        mpo: MatrixProductOperator,
        bond_dim: Optional[int] = None,
        block_size: int = 1,
        mps=None,
        exact_solver_dim: int = 200,
        *,
        chi: Optional[int] = None,
        compute_variance: bool = True,
        seed: Optional[int] = None,
        canonicalize: bool = False,
        split: str = "qr",
    ):
        if bond_dim is None:
            bond_dim = chi
        if bond_dim is None:
            raise TypeError("FiniteDMRG needs bond_dim (or its README alias chi)")
        self._n_sites = mpo.n_sites
        self._bond_dim = bond_dim
        self._phys_dim = mpo.phys_dim
        self._block_size = block_size
        self._exact_solver_dim = exact_solver_dim
        self._compute_variance = compute_variance
        if mps is None:
            mps = MatrixProductState.random(n=self.n_sites, bond_dim=self.bond_dim, phys_dim=self.phys_dim, seed=seed)
        # canonicalize=True right-canonicalises a user-supplied MPS on the device first; the reference
        # (and the default here) trusts the caller, as MatrixProductState.random is right-canonical
        # split="qr" (default): bonds >= 16 are orthogonalised by the verified Cholesky-QR split and the bond's
        # small SVD is deferred until `bond_singular_values` is read; split="svd": one Jacobi SVD per split,
        # the reference's literal gauge (matrix_product_state.py:187-225).  Energies, the state and the bond
        # spectra are the same either way (the two differ by an orthogonal gauge on each bond).
        self._env = Environment(mpo=mpo, mps=mps, canonicalize=canonicalize, split=split)
        self._energies: List[float] = [np.nan]
        self._variances: List[float] = [np.nan]
        self.solver_stats: List[Dict] = []  # one record per local solve of the last sweep
        #: set to a dict to collect synchronised per-phase wall seconds of the sweeps (benchmarking aid)
        self.phase_seconds: Optional[Dict[str, float]] = None

    # -- properties ---------------------------------------------------------------------------------
    bond_dim = property(lambda self: self._bond_dim)
    n_sites = property(lambda self: self._n_sites)
    phys_dim = property(lambda self: self._phys_dim)

    @property
    def mps(self) -> MatrixProductState:
        return self._env.mps

    @property
    def environment(self) -> Environment:
        return self._env

    @property
    def bond_singular_values(self) -> Dict[int, np.ndarray]:
        """Singular values of the most recent split on every bond (host copies).  Bonds split by the
        Cholesky-QR path run their deferred small SVD here, once."""
        return {b: s.cpu().numpy() for b, s in self._env.bond_singular_values.items()}

    def variance(self) -> float:
        return self._env.variance()

    # -- a4 / a2 / a5 ---------------------------------------------------------------------------------
    def _solve_on_device(self, site: int, tol: float, **kwargs) -> float:
        """Solve the local problem and leave the eigenvector in the environment's site tensor."""
        env = self._env
        psi = env.device_tensor(site)
        if psi.numel() < self._exact_solver_dim:
            energy, vec = eigh(env.one_site_full_matrix_device(site))
            psi.copy_(vec.reshape(psi.shape))
            env._dirty.add(site)
            self.solver_stats.append({"site": site, "dense": True, "n_matvec": 0})
            return float(energy.item())
        op = env.one_site_matvec(site)
        import torch

        image = torch.empty_like(psi)
        energy, vec = eigshmv(op, v0=psi, tol=tol, image=image, **kwargs)
        psi.copy_(vec.reshape(psi.shape))
        env._dirty.add(site)
        # H_eff psi of the vector just written, as the Lanczos relation gives it: perturb_wave_function(site)
        # uses it instead of a second matvec as long as nothing else touches the site tensor in between
        env.remember_image(site, image)
        self.solver_stats.append({"site": site, "dense": False, **op.last_stats})
        return float(energy)

    def one_site_solver(self, site: int, tol: float = 1e-8, **kwargs) -> Tuple[float, np.ndarray]:
        """Reference-shaped return: ``(energy, psi)`` with psi on the host, (N,) from the dense
        branch and (N, 1) from the iterative branch (finite_dmrg.py:97-111)."""
        env = self._env
        v0 = env.device_tensor(site)
        if v0.numel() < self._exact_solver_dim:
            energy, vec = eigh(env.one_site_full_matrix_device(site))
            return float(energy.item()), vec.cpu().numpy()
        energy, vec = eigshmv(env.one_site_matvec(site), v0=v0, tol=tol, **kwargs)
        return float(energy), vec.cpu().numpy()

    def two_site_solver(self, site: int, tol: float = 1e-8, **kwargs):
        return NotImplemented

    # -- a3 -------------------------------------------------------------------------------------------
    def perturb_wave_function(self, site: int, alpha: float = 1e-5):
        """psi <- psi + alpha * H_eff psi, in place, no renormalisation (finite_dmrg.py:116-141)."""
        env = self._env
        psi = env.device_tensor(site)
        hpsi = env.take_image(site)
        if hpsi is None:
            hpsi = env.one_site_matvec(site).apply_device(psi)
        _cuda.axpy(alpha, hpsi, psi)
        env._dirty.add(site)

    # -- a11 ------------------------------------------------------------------------------------------
    def sweep(self, direction: Direction = Direction.RIGHTWARD, tol: float = 1e-8, **kwargs) -> Optional[float]:
        sites = range(self.n_sites - 1) if direction == Direction.RIGHTWARD else range(self.n_sites - 1, 0, -1)
        energy = None
        self.solver_stats = []
        tick = self._phase_timer()
        for site in sites:
            energy = self._solve_on_device(site, tol, **kwargs)
            tick("eigensolve")
            logger.info(f"Sweeping to site [{site + 1}/{self.n_sites}], E0 = {energy}")
            self.perturb_wave_function(site)
            tick("perturb")
            self._env.split_tensor(site, direction=direction)
            tick("split")
            self._env.update(site, direction=direction)
            tick("env_update")
        return energy

    def _phase_timer(self):
        if self.phase_seconds is None:
            return lambda name: None
        import torch

        torch.cuda.synchronize()
        last = [time.perf_counter()]

        def tick(name: str):
            torch.cuda.synchronize()
            now = time.perf_counter()
            self.phase_seconds[name] = self.phase_seconds.get(name, 0.0) + now - last[0]
            last[0] = now

        return tick

    def _converged(self, n_sweep: int, tol: float, max_sweep: int, metric: Metric) -> bool:
        series = self._variances if metric == Metric.VARIANCE else self._energies
        gradient = np.diff(series[-2:])[0]
        logger.info(f"Metric {metric.name} is lowered by {gradient:e} in this sweep.")
        if abs(gradient) < tol:
            logger.info(f"Reaching set tolerance {tol}, stop sweeping.")
            return True
        if n_sweep == max_sweep:
            logger.warning(
                f"Maximum number of sweeps {max_sweep} is reached, yet {metric.name} gradient = {gradient:e} "
                f"is still greater than tol = {tol}."
            )
        elif abs(gradient) > tol and gradient < 0:
            logger.warning(
                f"Might be trapped in local minimum in this sweep, got {metric.name} gradient = {gradient:e}, "
                f"skip and proceed."
            )
        return False

    def run(self, tol: float = 1e-8, max_sweep: int = 100, metric: Metric = Metric.ENERGY, **kwargs) -> List[float]:
        clock = [time.perf_counter()]
        converged = partial(self._converged, tol=tol, max_sweep=max_sweep, metric=metric)
        logger.info(f"Set tolerance = {tol} to metric {metric.name}, up to maximally {max_sweep} sweeps.")
        n_sweep = 0
        for n_sweep, direction in zip(range(1, max_sweep + 1), cycle([Direction.RIGHTWARD, Direction.LEFTWARD])):
            logger.info(f"<==== In sweep epoch [{n_sweep}/{max_sweep}] ====>")
            energy = self.sweep(direction, tol=tol, **kwargs)
            clock.append(time.perf_counter())
            self._energies.append(energy)
            need_var = self._compute_variance or metric == Metric.VARIANCE
            self._variances.append(self._env.variance() if need_var else np.nan)
            logger.info(f"Last sweep took {timedelta(seconds=np.diff(clock[-2:])[0])}.")
            if converged(n_sweep):
                break
        elapsed = np.mean(np.sort(np.diff(clock))[:3])
        logger.info(f"Summary - {n_sweep} sweeps, best of {min(3, n_sweep)} - {timedelta(seconds=elapsed)} per sweep.")
        return self._energies[1:]

    #: README.md:101 spells the entry point ``update``
    update = run

    @property
    def measurements(self) -> MatrixProductStateMeasurements:
        if len(self._energies) == 1:
            raise RuntimeError("FiniteDMRG is probably not executed yet.")
        return MatrixProductStateMeasurements(self.mps)


class ShiftInvertDMRG(FiniteDMRG):
    """DMRG on the shift-invert spectrum (mirrors tnpy/finite_dmrg.py:266-407):

        (H - eps) phi = 1 / (E - eps) * (H - eps)^2 phi,        psi = (H - eps) phi.

    ``mpo`` is the MPO of ``H - eps``; a second environment over ``mpo.square()`` supplies the right-hand
    side of the generalised local problem, which ``tnpy_geig_lowest`` solves on the device (the
    reference: ``primme.eigsh(A, M=M)`` in the bulk, ``scipy.linalg.eigh(a, b)`` for tiny sites -- here
    one on-device generalised Davidson serves both; for N <= 40 its basis spans the whole space)."""

    def __init__(self, mpo, bond_dim: Optional[int] = None, offset: float = 0, block_size: int = 1, mps=None,
                 exact_solver_dim: int = 200, *, chi: Optional[int] = None, seed: Optional[int] = None,
                 split: str = "qr"):
        super().__init__(mpo, bond_dim=bond_dim, block_size=block_size, mps=mps, exact_solver_dim=exact_solver_dim,
                         chi=chi, seed=seed, split=split)
        self._env2 = Environment(mpo=mpo.square(), mps=self.mps, share_state_with=self._env)
        self._offset = offset
        self._restored_mps = None

    @property
    def restored_mps(self) -> Optional[MatrixProductState]:
        return self._restored_mps

    def _restore_mps(self):
        """|psi> = (H - eps)|phi> as an MPS of bond chi * w (finite_dmrg.py:313-339): the fused bonds are
        (MPS bond, MPO bond) with the MPS bond slow."""
        mps, mpo = self.mps, self._env.mpo
        arrays = []
        for site in range(self.n_sites):
            a, w = mps.three_leg(site), mpo.as_four_leg(site)  # (l, b, r), (wl, wr, k, b)
            t = np.einsum("lbr,xykb->lxkry", a, w)
            l, wl, k, r, wr = t.shape
            arrays.append(t.reshape(l * wl, k, r * wr))
        arrays[0] = arrays[0][0]
        arrays[-1] = arrays[-1][:, :, 0]
        self._restored_mps = MatrixProductState(arrays)

    #: The projected H^2 is too ill-conditioned (1e9 and up) for inverse-free Krylov iterations on the pencil itself
    #: (csrc/geig.cu), so the local pencil is formed densely on the device and reduced to a standard problem:
    #: up to ``dense_pencil_dim`` unknowns through a Jacobi SVD of the right-hand matrix (tolerates a numerically
    #: singular one; the reference's own test size -- n=10, chi=64: 1024 unknowns -- is in this range), up to
    #: ``cholesky_pencil_dim`` through a Cholesky factor and the on-device Lanczos solver (tnpy_geig_chol_lowest:
    #: O(N^3) tensor-pipe work, seconds at N = 16384; needs 6 N^2 doubles).  Larger sites fall to the iterative
    #: solver, which raises RuntimeError when it does not converge.
    dense_pencil_dim = 2048
    cholesky_pencil_dim = 32768

    def _solve_on_device(self, site: int, tol: float, **kwargs) -> float:
        env, env2 = self._env, self._env2
        psi = env.device_tensor(site)
        if psi.numel() <= max(self.dense_pencil_dim, self.cholesky_pencil_dim):
            a = env.one_site_full_matrix_device(site)
            b = env2.one_site_full_matrix_device(site)
            if psi.numel() <= self.dense_pencil_dim:
                theta, x = _cuda.geig_dense_lowest(a, b)
                stats = {"n_matvec": 0}
            else:
                theta, x, stats = _cuda.geig_chol_lowest(a, b, tol=min(tol, 1e-10))
                if not stats["converged"]:
                    raise RuntimeError(
                        f"ShiftInvertDMRG: the Lanczos solve of the reduced pencil at site {site} ({psi.numel()} unknowns) "
                        f"stopped at residual {stats['resid']:.3e} after {stats['n_matvec']} matvecs; the site tensor was "
                        "left unchanged."
                    )
            psi.copy_(x.reshape(psi.shape))
            env._dirty.add(site)
            self.solver_stats.append({"site": site, "dense": True, **stats})
            return float(theta.item())
        la, wa, ra = env.operands(site)
        lm, wm, rm = env2.operands(site)
        opts = {}
        if "ncv" in kwargs:
            opts["ncv"] = int(kwargs["ncv"])
        if "maxiter" in kwargs:
            opts["max_iter"] = int(kwargs["maxiter"])
        saved = psi.clone()
        stats = _cuda.geig_lowest(la, wa, ra, lm, wm, rm, psi, tol=tol, flags_a=env.gauge_flags(site), **opts)
        if not stats["converged"]:
            # never hand back an unconverged vector as if it were the local ground state: the projected
            # (H - eps)^2 is too ill-conditioned for the inverse-free iteration beyond small sites
            psi.copy_(saved)
            raise RuntimeError(
                f"ShiftInvertDMRG: the iterative pencil solve at site {site} ({psi.numel()} unknowns) did not converge "
                f"(residual {stats['resid']:.3e} after {stats['n_iter']} iterations); the site tensor was left unchanged. "
                f"Sites up to cholesky_pencil_dim = {self.cholesky_pencil_dim} unknowns are solved through dense "
                "factorisations -- raise it (memory: 6 N^2 doubles) or lower the bond dimension."
            )
        env._dirty.add(site)
        self.solver_stats.append({"site": site, "dense": False, "n_matvec": 2 * stats["n_iter"], **stats})
        return float(stats["theta"])

    def one_site_solver(self, site: int, tol: float = 1e-8, **kwargs) -> Tuple[float, np.ndarray]:
        saved = self._env.device_tensor(site).clone()
        energy = self._solve_on_device(site, tol, **kwargs)
        vec = self._env.device_tensor(site).reshape(-1, 1).cpu().numpy()
        self._env.device_tensor(site).copy_(saved)
        return energy, vec

    def sweep(self, direction: Direction = Direction.RIGHTWARD, tol: float = 1e-8, **kwargs) -> Optional[float]:
        sites = range(self.n_sites - 1) if direction == Direction.RIGHTWARD else range(self.n_sites - 1, 0, -1)
        energy = None
        self.solver_stats = []
        for site in sites:
            energy = self._solve_on_device(site, tol, **kwargs)
            logger.info(f"Sweeping to site [{site + 1}/{self.n_sites}], E0 = {1 / energy + self._offset}")
            self.perturb_wave_function(site)
            self._env.split_tensor(site, direction=direction)
            self._env.update(site, direction=direction)
            self._env2.update(site, direction=direction)  # site tensors are shared with the first environment
        return energy

    def run(self, tol: float = 1e-7, max_sweep: int = 100, metric: Metric = Metric.ENERGY, **kwargs) -> List[float]:
        energies = super().run(tol, max_sweep, metric, **kwargs)
        self._restore_mps()
        return (np.reciprocal(energies) + self._offset).tolist()

    update = run

    @property
    def measurements(self) -> MatrixProductStateMeasurements:
        if len(self._energies) == 1:
            raise RuntimeError("FiniteDMRG is probably not executed yet.")
        return MatrixProductStateMeasurements(self.restored_mps)

    def variance(self) -> float:
        """finite_dmrg.py:401-407, verbatim arithmetic on the restored state."""
        meas = MatrixProductStateMeasurements(self.restored_mps)
        return meas.expectation_value(self._env.mpo.square()) - self._energies[-1] ** 2 - self._offset**2
