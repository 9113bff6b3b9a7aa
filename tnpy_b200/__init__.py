"""tnpy_b200 -- B200-native finite-DMRG local-update path behind tnpy's Python API.

Mirrors the public surface of tanlin2013/tnpy for this path (``FiniteDMRG``, the ``model`` MPOs,
``MatrixProductState``); the numerical work runs in hand-written sm_100a CUDA kernels reached through
the C ABI in ``include/tnpy_cuda.h``.  Logging follows the reference (tnpy/__init__.py:10-20): logger
"tnpy", INFO, one stream handler.
"""
import logging

__version__ = "0.1.0"

logger = logging.getLogger("tnpy")
if not logger.handlers:
    _handler = logging.StreamHandler()
    _handler.setLevel(logging.INFO)
    _handler.setFormatter(
        logging.Formatter("%(asctime)s [%(filename)s] %(levelname)s: %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    )
    logger.addHandler(_handler)
logger.setLevel(logging.INFO)
