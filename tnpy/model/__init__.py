"""Alias of :mod:`tnpy_b200.model` under the reference's module path (tnpy/model/)."""
from tnpy_b200.model import *  # noqa: F401,F403
from tnpy_b200.model import DimerXXZ, Model1D, RandomHeisenberg, Thirring, TotalSz, TransverseIsing, XXZ  # noqa: F401
