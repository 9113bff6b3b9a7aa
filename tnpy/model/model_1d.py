"""Alias under the reference's module path (tnpy/model/model_1d.py)."""
from tnpy_b200.model import Model1D  # noqa: F401
