"""Built-in spin-chain MPOs.  Host NumPy, bit-exact with the reference's arithmetic:

    XXZ              tnpy/model/xxz.py:19-30                 w = 5
    Thirring         tnpy/model/thirring.py:40-65            w = 6 (5 if penalty == 0)
    RandomHeisenberg tnpy/model/random_heisenberg.py:84-109  w = 6 (5 if penalty == 0), seeded fields
    DimerXXZ         tnpy/model/dimer_xxz.py:42-63           spin-1 prefactors, unseeded draws per call
    TransverseIsing  tnpy/model/transverse_ising.py:32-40    w = 3
    TotalSz          tnpy/model/total_sz.py:19-55            w = 2

Each scalar coefficient is formed with the same operations in the same order as the reference, so
the float64 values agree bit for bit (pinned by tests against tests/test_operators.py:46-88).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from tnpy_b200.model.model_1d import Model1D, drop_channel_if, operator_matrix
from tnpy_b200.operators import MatrixProductOperator, SpinOperators


class XXZ(Model1D):
    def __init__(self, n: int, delta: float):
        super().__init__(n)
        self.delta = delta

    def _bulk_elem(self, site: int) -> np.ndarray:
        Sp, Sm, Sz, I2, O2 = SpinOperators()
        hop_p, hop_m, zz = -0.5 * Sp, -0.5 * Sm, -self.delta * Sz
        return operator_matrix(
            [
                [I2, hop_p, hop_m, zz, O2],
                [O2, O2, O2, O2, Sm],
                [O2, O2, O2, O2, Sp],
                [O2, O2, O2, O2, Sz],
                [O2, O2, O2, O2, I2],
            ]
        )


def _six_channel(first_row: list, closing: list, I2: np.ndarray, O2: np.ndarray) -> np.ndarray:
    """Upper-triangular 6-channel MPO with a penalty channel (index 3) that carries an identity."""
    zero_row = [O2] * 5
    rows = [first_row]
    for k in range(1, 5):
        row = list(zero_row) + [closing[k - 1]]
        if k == 3:
            row[3] = I2
        rows.append(row)
    rows.append(list(zero_row) + [I2])
    return operator_matrix(rows)


class Thirring(Model1D):
    def __init__(self, n: int, delta: float, ma: float, penalty: float, s_target: int) -> None:
        super().__init__(n)
        self.delta = delta
        self.ma = ma
        self.penalty = penalty
        self.s_target = s_target

    def _bulk_elem(self, site: int) -> np.ndarray:
        Sp, Sm, Sz, I2, O2 = SpinOperators()
        beta = self.delta + ((-1.0) ** site * self.ma) - 2.0 * self.penalty * self.s_target
        gamma = self.penalty * (0.25 + self.s_target**2 / self.n) + 0.25 * self.delta
        first = [I2, -0.5 * Sp, -0.5 * Sm, 2.0 * np.sqrt(self.penalty) * Sz, self.delta * Sz, gamma * I2 + beta * Sz]
        closing = [Sm, Sp, np.sqrt(self.penalty) * Sz, Sz]
        return drop_channel_if(self.penalty == 0, _six_channel(first, closing, I2, O2), 3, 3)


class RandomHeisenberg(Model1D):
    def __init__(
        self,
        n: int,
        h: float,
        penalty: float = 0,
        s_target: int = 0,
        offset: float = 0,
        trial_id: Optional[str] = None,
        seed: Optional[int] = None,
    ):
        super().__init__(n)
        self._h = h
        self._penalty = penalty
        self._s_target = s_target
        self._offset = offset
        self._trial_id = trial_id
        self.seed = seed  # draws the fields (random_heisenberg.py:57-58)

    h = property(lambda self: self._h)
    penalty = property(lambda self: self._penalty)
    s_target = property(lambda self: self._s_target)
    trial_id = property(lambda self: self._trial_id)

    @property
    def offset(self) -> float:
        return self._offset

    @offset.setter
    def offset(self, offset: float):
        self._offset = offset

    @property
    def seed(self) -> Optional[int]:
        return self._seed

    @seed.setter
    def seed(self, seed: Optional[int]) -> None:
        self._seed = seed
        self._random_sequence = np.random.RandomState(seed).uniform(-self.h, self.h, size=self.n)

    def _bulk_elem(self, site: int) -> np.ndarray:
        Sp, Sm, Sz, I2, O2 = SpinOperators()
        alpha = self.penalty * (0.25 + self.s_target**2 / self.n) - self.offset / self.n
        beta = self._random_sequence[site] - 2.0 * self.penalty * self.s_target
        first = [I2, 0.5 * Sp, 0.5 * Sm, 2.0 * self.penalty * Sz, Sz, alpha * I2 + beta * Sz]
        return drop_channel_if(self.penalty == 0, _six_channel(first, [Sm, Sp, Sz, Sz], I2, O2), 3, 3)


class DimerXXZ(Model1D):
    def __init__(self, n: int, J: float, delta: float, h: float, penalty: float = 0, s_target: int = 0,
                 trial_id: Optional[str] = None):
        super().__init__(n)
        self.J = J
        self.delta = delta
        self.h = h
        self.penalty = penalty
        self.s_target = s_target
        self.trial_id = trial_id

    def _bulk_elem(self, site: int) -> np.ndarray:
        Sp, Sm, Sz, I2, O2 = SpinOperators(spin=1)
        rand_J = (1 + self.delta * (-1) ** site) * np.random.uniform() ** self.J
        alpha = self.penalty * (0.25 + self.s_target**2 / self.n)
        beta = np.random.uniform(-self.h, self.h) - 2.0 * self.penalty * self.s_target
        first = [I2, 0.5 * rand_J * Sp, 0.5 * rand_J * Sm, 2.0 * self.penalty * Sz, Sz, alpha * I2 + beta * Sz]
        return drop_channel_if(self.penalty == 0, _six_channel(first, [Sm, Sp, Sz, Sz], I2, O2), 3, 3)


class TransverseIsing(Model1D):
    def __init__(self, n, j, h):
        super().__init__(n)
        self._j = j
        self._h = h

    j = property(lambda self: self._j)
    h = property(lambda self: self._h)

    def _bulk_elem(self, site: int) -> np.ndarray:
        Sp, Sm, Sz, I2, O2 = SpinOperators()
        Sx = Sp + Sm
        return operator_matrix([[I2, -self.j * Sz, -self.j * self.h * Sx], [O2, O2, Sz], [O2, O2, I2]])


class TotalSz(Model1D):
    def _bulk_elem(self, site: int) -> np.ndarray:
        _, _, Sz, I2, O2 = SpinOperators()
        return operator_matrix([[I2, Sz], [O2, I2]])

    def _identity_elem(self, site: int) -> np.ndarray:
        _, _, _, I2, O2 = SpinOperators()
        full = operator_matrix([[I2, O2], [O2, I2]])
        return full[0] if site == 0 else (full[:, -1] if site == self.n - 1 else full)

    def subsystem_mpo(self, partition_site: int) -> MatrixProductOperator:
        if not 0 <= partition_site < self.n:
            raise ValueError("Partition site must be in between 0 and the system size n.")
        return MatrixProductOperator(
            [self._elem(s) if s <= partition_site else self._identity_elem(s) for s in range(self.n)]
        )
