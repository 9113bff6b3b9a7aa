"""Alias under the reference's module path (tnpy/model/dimer_xxz.py)."""
from tnpy_b200.model import DimerXXZ  # noqa: F401
