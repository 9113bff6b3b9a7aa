"""Alias under the reference's module path (tnpy/model/thirring.py)."""
from tnpy_b200.model import Thirring  # noqa: F401
