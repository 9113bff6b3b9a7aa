"""``tnpy`` -- the reference's import name, served by the B200-native implementation.

A user of tanlin2013/tnpy keeps their imports: ``from tnpy.finite_dmrg import FiniteDMRG``,
``from tnpy.model import XXZ, Thirring, RandomHeisenberg``, ``from tnpy.matrix_product_state import
MatrixProductState`` ... all resolve to ``tnpy_b200`` (same classes, not copies).  Only the modules of the
finite-DMRG path exist here (tSDRG, TDVP and exact diagonalisation are out of scope).
"""
from tnpy_b200 import __version__, logger  # noqa: F401
