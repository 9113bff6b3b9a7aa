"""Base class of the built-in 1-D models (mirrors tnpy/model/model_1d.py:8-33)."""
from __future__ import annotations

import abc
from typing import List

import numpy as np

from tnpy_b200.operators import MatrixProductOperator


class Model1D(abc.ABC):
    #: MPO row kept at the left end / column kept at the right end (model/utils.py:7-32)
    boundary_row = 0
    boundary_col = -1

    def __init__(self, n: int):
        self._n = n

    @property
    def n(self) -> int:
        return self._n

    @abc.abstractmethod
    def _bulk_elem(self, site: int) -> np.ndarray:
        """Full (w, w, d, d) operator-valued matrix of ``site`` before boundary selection."""

    def _elem(self, site: int) -> np.ndarray:
        full = self._bulk_elem(site)
        if site == 0:
            return full[self.boundary_row]
        if site == self.n - 1:
            return full[:, self.boundary_col]
        return full

    @property
    def mpo(self) -> MatrixProductOperator:
        return MatrixProductOperator([self._elem(site) for site in range(self.n)])


def operator_matrix(rows: List[list]) -> np.ndarray:
    """Stack a nested list of 2x2 blocks into a (w, w, d, d) float array."""
    return np.array(rows, dtype=float)


def drop_channel_if(flag: bool, mat: np.ndarray, row: int, col: int) -> np.ndarray:
    """model/utils.py:35-64 -- remove the penalty row/column when the penalty is switched off."""
    return np.delete(np.delete(mat, row, axis=0), col, axis=1) if flag else mat
