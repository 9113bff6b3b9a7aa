"""Alias under the reference's module path (tnpy/model/xxz.py)."""
from tnpy_b200.model import XXZ  # noqa: F401
