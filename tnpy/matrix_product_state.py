"""Alias of :mod:`tnpy_b200.matrix_product_state` under the reference's module path (tnpy/matrix_product_state.py)."""
from tnpy_b200.matrix_product_state import *  # noqa: F401,F403
from tnpy_b200.matrix_product_state import __dict__ as _d

globals().update({k: v for k, v in _d.items() if not k.startswith("__")})
del _d
