"""Alias of :mod:`tnpy_b200.finite_dmrg` under the reference's module path (tnpy/finite_dmrg.py)."""
from tnpy_b200.finite_dmrg import *  # noqa: F401,F403
from tnpy_b200.finite_dmrg import __dict__ as _d

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
del _d
