"""Alias of :mod:`tnpy_b200.operators` under the reference's module path (tnpy/operators.py)."""
from tnpy_b200.operators import *  # noqa: F401,F403
from tnpy_b200.operators import __dict__ as _d

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
del _d
