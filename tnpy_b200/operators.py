"""Spin operators, the MPO container and the dense-Hamiltonian helper (host side, bit-exact).

Mirrors tnpy/operators.py (SpinOperators :10-63, MatrixProductOperator :66-116, FullHamiltonian
:119-166).  The reference subclasses quimb's MPO; quimb is not a dependency here, so the container
keeps just what the DMRG path and the reference's tests touch: per-site arrays in 'lrud' layout
(edge tensors without the outer bond), ``nsites``/``n_sites``, ``phys_dim``, indexing, scalar
multiplication and ``square()``.  MPO construction never touches the GPU.
"""
from __future__ import annotations

from dataclasses import InitVar, astuple, dataclass, field
from typing import List, Sequence

import numpy as np


@dataclass
class SpinOperators:
    """``Sp, Sm, Sz, I2, O2 = SpinOperators(spin)`` -- unpacking order matters (operators.py:44-46)."""

    spin: InitVar[float] = field(default=0.5)
    Sp: np.ndarray = field(init=False)
    Sm: np.ndarray = field(init=False)
    Sz: np.ndarray = field(init=False)
    I2: np.ndarray = field(init=False)
    O2: np.ndarray = field(init=False)

    def __post_init__(self, spin: float):
        raising = np.zeros((2, 2), dtype=float)
        raising[0, 1] = 2.0
        self.Sp = spin * raising
        self.Sm = spin * raising.T.copy()
        self.Sz = spin * np.diag([1.0, -1.0])
        self.I2 = np.eye(2, dtype=float)
        self.O2 = np.zeros((2, 2), dtype=float)

    def __iter__(self):
        return iter(astuple(self))


class _SiteOperator:
    """Minimal stand-in for the quimb tensor a site lookup returns: ``.data``, ``.shape``."""

    __slots__ = ("data",)

    def __init__(self, data: np.ndarray):
        self.data = data

    @property
    def shape(self):
        return self.data.shape


class MatrixProductOperator:
    """Per-site arrays: site 0 ``(w_r, d, d)``, bulk ``(w_l, w_r, d, d)``, last ``(w_l, d, d)``."""

    def __init__(self, arrays: Sequence[np.ndarray]):
        self._arrays: List[np.ndarray] = [np.ascontiguousarray(a, dtype=float) for a in arrays]
        if len(self._arrays) < 2:
            raise ValueError("An MPO needs at least two sites.")
        dims = {a.shape[-1] for a in self._arrays} | {a.shape[-2] for a in self._arrays}
        if len(dims) != 1:  # operators.py:76-79
            raise ValueError("All MPO tensors are assumed to have same physical dims.")
        self._phys_dim = dims.pop()

    # -- quimb-flavoured accessors the reference code relies on
    @property
    def nsites(self) -> int:
        return len(self._arrays)

    @property
    def n_sites(self) -> int:
        return len(self._arrays)

    @property
    def phys_dim(self) -> int:
        return self._phys_dim

    @property
    def arrays(self) -> List[np.ndarray]:
        return self._arrays

    def __len__(self) -> int:
        return len(self._arrays)

    def __getitem__(self, site: int) -> _SiteOperator:
        return _SiteOperator(self._arrays[site])

    def __iter__(self):
        return (_SiteOperator(a) for a in self._arrays)

    def __mul__(self, scalar: float) -> "MatrixProductOperator":
        # quimb spreads a scalar over the sites as |x|**(1/n) with the sign on one tensor; dense H
        # of (-1 * mpo) must equal -H exactly (tests/test_operators.py:30-36), which holds for -1.
        n = self.nsites
        mag = abs(scalar) ** (1.0 / n)
        arrays = [a * mag for a in self._arrays]
        if scalar < 0:
            arrays[0] = -arrays[0]
        return MatrixProductOperator(arrays)

    __rmul__ = __mul__

    def bond_dims(self) -> List[int]:
        """Bond i joins sites i and i + 1: the right bond of every site but the last (site 0 is (w_r, d, d))."""
        return [a.shape[0] if i == 0 else a.shape[1] for i, a in enumerate(self._arrays[:-1])]

    def as_four_leg(self, site: int) -> np.ndarray:
        """Site tensor with explicit unit bonds at the chain ends: always (w_l, w_r, d, d)."""
        a = self._arrays[site]
        if a.ndim == 4:
            return a
        return a[None] if site == 0 else a[:, None]

    def square(self) -> "MatrixProductOperator":
        """Two stacked layers fused into one MPO (operators.py:91-116); fused bond = (upper, lower)
        with the first layer slow."""
        out = []
        for site in range(self.nsites):
            w = self.as_four_leg(site)
            wl, wr, d, _ = w.shape
            two = np.einsum("acpx,bdxq->abcdpq", w, w).reshape(wl * wl, wr * wr, d, d)
            if site == 0:
                two = two[0]
            elif site == self.nsites - 1:
                two = two[:, 0]
            out.append(two)
        return MatrixProductOperator(out)


class FullHamiltonian:
    """Dense matrix of an MPO, rows = ket indices, columns = bra indices (operators.py:119-166)."""

    def __init__(self, mpo: MatrixProductOperator):
        self._n_sites = mpo.n_sites
        self._phys_dim = mpo.phys_dim
        if self.phys_dim**self.n_sites > 2**12:  # operators.py:140-143
            raise ResourceWarning(f"Requesting more than {self.n_sites} sites with physical dim {self.phys_dim}.")
        acc = np.ones((1, 1, 1))  # (ket, bra, mpo bond)
        for site in range(mpo.n_sites):
            w = mpo.as_four_leg(site)
            acc = np.einsum("KBa,abpq->KpBqb", acc, w)
            acc = acc.reshape(acc.shape[0] * acc.shape[1], acc.shape[2] * acc.shape[3], acc.shape[4])
        self._matrix = acc[:, :, 0]

    @property
    def n_sites(self) -> int:
        return self._n_sites

    @property
    def phys_dim(self) -> int:
        return self._phys_dim

    @property
    def matrix(self) -> np.ndarray:
        return self._matrix
