"""Alias under the reference's module path (tnpy/model/total_sz.py)."""
from tnpy_b200.model import TotalSz  # noqa: F401
