"""Alias of :mod:`tnpy_b200.linalg` under the reference's module path (tnpy/linalg.py)."""
from tnpy_b200.linalg import *  # noqa: F401,F403
from tnpy_b200.linalg import __dict__ as _d

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
del _d
