"""Alias under the reference's module path (tnpy/model/transverse_ising.py)."""
from tnpy_b200.model import TransverseIsing  # noqa: F401
