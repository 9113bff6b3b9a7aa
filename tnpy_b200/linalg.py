"""Factorisation and eigensolver seams (mirrors tnpy/linalg.py), executed on the GPU.

    svd(matrix, cutoff)                 linalg.py:9-23    -> tnpy_svd        (one-sided Jacobi)
    eigh(matrix, k=1)                   linalg.py:42-61   -> tnpy_eigh_lowest
    eigshmv(linear_operator, v0, tol)   linalg.py:64-87   -> tnpy_eig_lowest (on-device Lanczos)

NumPy arrays in, NumPy arrays out (CUDA tensors in, CUDA tensors out); same return shapes as the
reference: ``eigh`` gives ``(eval, evec (N,))``, ``eigshmv`` gives ``(eval, evec (N, 1))``.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from tnpy_b200 import _cuda, logger

# primme.eigsh keyword -> our solver keyword; anything else is accepted and ignored (with a debug line),
# as the reference forwards **kwargs verbatim to primme (finite_dmrg.py:97-111).
_PRIMME_KWARGS = {"ncv": "ncv", "maxBasisSize": "ncv", "maxiter": "max_matvec", "maxMatvecs": "max_matvec"}


def _to_device(x):
    import torch

    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=torch.float64).contiguous(), True
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda(), False


def svd(matrix, cutoff: int):
    """Thin SVD, singular values descending, keep the first ``cutoff``."""
    a, was_tensor = _to_device(matrix)
    u, s, vt = _cuda.svd(a.clone())
    u, s, vt = u[:, :cutoff], s[:cutoff], vt[:cutoff, :]
    if was_tensor:
        return u, s, vt
    return u.cpu().numpy(), s.cpu().numpy(), vt.cpu().numpy()


def eigh(matrix, k: int = 1, backend: str = "numpy", **kwargs):
    """Lowest eigenpair of a dense real-symmetric matrix: ``(evals[0], evecs[:, 0])``.
    Only ``k == 1`` is on the DMRG path (finite_dmrg.py:110); ``backend`` is accepted for
    signature compatibility."""
    if k != 1:
        raise NotImplementedError("tnpy_b200.linalg.eigh implements the k=1 case used by FiniteDMRG")
    a, was_tensor = _to_device(matrix)
    ev, vec = _cuda.eigh_lowest(a.clone())
    if was_tensor:
        return ev, vec
    return float(ev.item()), vec.cpu().numpy()


def eigshmv(linear_operator, v0, k: int = 1, which: str = "SA", tol: float = 0, **kwargs) -> Tuple[float, np.ndarray]:
    """Lowest eigenpair of an ``Environment.one_site_matvec`` operator, solved on the device.
    ``image=<device tensor>`` (not a primme option) additionally receives H_eff psi for the returned psi."""
    from tnpy_b200.matrix_product_state import HeffOperator

    if not isinstance(linear_operator, HeffOperator):
        raise TypeError("eigshmv runs on the GPU and needs the HeffOperator returned by Environment.one_site_matvec")
    if k != 1 or which != "SA":
        raise NotImplementedError("only k=1, which='SA' (the FiniteDMRG call) is implemented")
    opts = {}
    image = kwargs.pop("image", None)
    for key, value in kwargs.items():
        if key in _PRIMME_KWARGS:
            opts[_PRIMME_KWARGS[key]] = int(value)
        else:
            logger.debug(f"eigshmv: ignoring primme option {key}={value!r}")
    psi, was_tensor = _to_device(v0)
    psi = psi.reshape(linear_operator.site_shape).clone()
    L, W, R = linear_operator.env.operands(linear_operator.site)
    flags = linear_operator.env.gauge_flags(linear_operator.site)
    stats = _cuda.eig_lowest(L, W, R, psi, tol=tol, flags=flags, image=image, **opts)
    if not stats["converged"]:
        logger.warning(f"eigshmv: not converged after {stats['n_matvec']} matvecs, residual {stats['resid']:.3e}")
    linear_operator.last_stats = stats
    evec = psi.reshape(-1, 1)
    return stats["theta"], (evec if was_tensor else evec.cpu().numpy())
