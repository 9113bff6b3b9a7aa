"""Alias under the reference's module path (tnpy/model/random_heisenberg.py)."""
from tnpy_b200.model import RandomHeisenberg  # noqa: F401
