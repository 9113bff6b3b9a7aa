"""ctypes binding of ``libtnpy_cuda.so`` -- the C ABI declared in ``include/tnpy_cuda.h``.

PyTorch is used for device memory and streams only; every numerical operation of the hot path goes
through this library.  There is no CPU fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path
from typing import Optional

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libtnpy_cuda.so"

GEMM_AUTO, GEMM_GENERIC, GEMM_DMMA, GEMM_OZAKI, GEMM_FP64 = 0, 1, 2, 3, 4
LEFT_IDENTITY, RIGHT_IDENTITY = 1, 2
ENOCONV = -4

_PD = c_void_p  # device double*

# name -> (restype, argtypes); mirrors include/tnpy_cuda.h one to one
SIGNATURES = {
    "tnpy_version": (c_int, []),
    "tnpy_last_error": (c_char_p, []),
    "tnpy_launch_count": (c_int64, []),
    "tnpy_set_gemm_algo": (c_int, [c_int]),
    "tnpy_set_gemm_tile": (c_int, [c_int]),
    "tnpy_probe_fp64": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_double), c_void_p]),
    "tnpy_gemm_tn": (c_int, [_PD, c_int64, _PD, c_int64, _PD, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tnpy_ozaki_workspace_bytes": (c_size_t, [c_int] * 4),
    "tnpy_set_ozaki_slices": (c_int, [c_int]),
    "tnpy_set_fused_steps": (c_int, [c_int]),
    "tnpy_set_inexact_slices": (c_int, [c_int]),
    "tnpy_steps_trace": (c_int, [c_void_p]),
    "tnpy_ozaki_gemm_tn": (
        c_int,
        [_PD, c_int64, _PD, c_int64, _PD, c_int64] + [c_int] * 6 + [c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_ozaki_error_bound": (c_int, [c_int] * 4 + [c_void_p, c_size_t, _PD, c_void_p]),
    "tnpy_heff_workspace_bytes": (c_size_t, [c_int] * 5),
    "tnpy_heff_plan_bytes": (c_size_t, [c_int] * 5),
    "tnpy_heff_plan_create": (
        c_int,
        [ctypes.POINTER(c_void_p), _PD, _PD, _PD, ctypes.POINTER(c_double)] + [c_int] * 7 + [c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_heff_plan_create_rows": (
        c_int,
        [ctypes.POINTER(c_void_p), _PD, _PD, _PD, ctypes.POINTER(c_double)] + [c_int] * 9 + [c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_heff_plan_mode": (c_int, [c_void_p]),
    "tnpy_heff_plan_apply": (c_int, [c_void_p, _PD, _PD, c_int, c_void_p, c_size_t, c_void_p]),
    "tnpy_heff_plan_error_bound": (c_int, [c_void_p, _PD, c_void_p]),
    "tnpy_heff_plan_destroy": (c_int, [c_void_p]),
    "tnpy_heff_apply": (c_int, [_PD, _PD, _PD, _PD, _PD] + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "tnpy_identity_defect": (c_int, [_PD, c_int, c_int, c_int, _PD, c_void_p, c_size_t, c_void_p]),
    "tnpy_heff_apply_rows": (c_int, [_PD, _PD, _PD, _PD, _PD] + [c_int] * 8 + [c_void_p, c_size_t, c_void_p]),
    "tnpy_env_workspace_bytes": (c_size_t, [c_int] * 5),
    "tnpy_env_update_left": (c_int, [_PD, _PD, _PD, _PD] + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "tnpy_env_update_right": (c_int, [_PD, _PD, _PD, _PD] + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "tnpy_env_update_left_rows": (c_int, [_PD, _PD, _PD, _PD] + [c_int] * 8 + [c_void_p, c_size_t, c_void_p]),
    "tnpy_heff_dense_workspace_bytes": (c_size_t, [c_int] * 5),
    "tnpy_heff_dense": (c_int, [_PD, _PD, _PD, _PD] + [c_int] * 5 + [c_void_p, c_size_t, c_void_p]),
    "tnpy_dot": (c_int, [_PD, _PD, c_int64, _PD, c_void_p]),
    "tnpy_nrm2": (c_int, [_PD, c_int64, _PD, c_void_p]),
    "tnpy_axpy": (c_int, [c_double, _PD, _PD, c_int64, c_void_p]),
    "tnpy_axpy_dev": (c_int, [_PD, c_double, _PD, _PD, c_int64, c_void_p]),
    "tnpy_scal": (c_int, [c_double, _PD, c_int64, c_void_p]),
    "tnpy_multi_dot": (c_int, [_PD, c_int64, c_int, _PD, c_int64, _PD, c_void_p]),
    "tnpy_multi_axpy": (c_int, [_PD, c_int64, c_int, _PD, _PD, c_int64, c_void_p]),
    "tnpy_eig_workspace_bytes": (c_size_t, [c_int] * 6),
    "tnpy_eig_lowest": (
        c_int,
        [_PD, _PD, _PD, _PD] + [c_int] * 6 + [c_double, c_int, c_int, ctypes.POINTER(c_double), c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_last_eig_counters": (c_int, [ctypes.POINTER(c_int64), c_int]),
    "tnpy_eig_lowest_image": (
        c_int,
        [_PD, _PD, _PD, _PD, _PD] + [c_int] * 6 + [c_double, c_int, c_int, ctypes.POINTER(c_double), c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_comm_unique_id": (c_int, [ctypes.c_char_p]),
    "tnpy_comm_create": (c_int, [ctypes.POINTER(c_void_p), ctypes.c_char_p, c_int, c_int]),
    "tnpy_comm_destroy": (c_int, [c_void_p]),
    "tnpy_comm_world": (c_int, [c_void_p]),
    "tnpy_comm_rank": (c_int, [c_void_p]),
    "tnpy_comm_allgather": (c_int, [c_void_p, _PD, _PD, c_int64, c_void_p]),
    "tnpy_comm_allreduce_sum": (c_int, [c_void_p, _PD, c_int64, c_void_p]),
    "tnpy_eig_rows_workspace_bytes": (c_size_t, [c_int] * 7),
    "tnpy_eig_lowest_rows": (
        c_int,
        [c_void_p, _PD, _PD, _PD, _PD, _PD] + [c_int] * 8 + [c_double, c_int, c_int, ctypes.POINTER(c_double), c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_geig_workspace_bytes": (c_size_t, [c_int] * 8),
    "tnpy_geig_lowest": (
        c_int,
        [_PD] * 7 + [c_int] * 8 + [c_double, c_int, c_int, ctypes.POINTER(c_double), c_void_p, c_size_t, c_void_p],
    ),
    "tnpy_geig_dense_workspace_bytes": (c_size_t, [c_int]),
    "tnpy_geig_dense_lowest": (c_int, [_PD, _PD, c_int, _PD, _PD, c_void_p, c_size_t, c_void_p]),
    "tnpy_geig_chol_workspace_bytes": (c_size_t, [c_int]),
    "tnpy_geig_chol_lowest": (c_int, [_PD, _PD, c_int, c_double, c_int, _PD, _PD, ctypes.POINTER(c_double), c_void_p,
                              c_size_t, c_void_p]),
    "tnpy_eigh_workspace_bytes": (c_size_t, [c_int]),
    "tnpy_eigh_lowest": (c_int, [_PD, c_int, _PD, _PD, c_void_p, c_size_t, c_void_p]),
    "tnpy_svd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tnpy_last_svd_sweeps": (c_int, []),
    "tnpy_last_svd_trace": (c_int, [ctypes.POINTER(ctypes.c_uint), c_int]),
    "tnpy_svd": (c_int, [_PD, c_int, c_int, _PD, _PD, _PD, c_void_p, c_size_t, c_void_p]),
    "tnpy_qr_split_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tnpy_qr_split": (c_int, [_PD, c_int, c_int, _PD, _PD, _PD, c_int, c_void_p, c_size_t, c_void_p]),
    "tnpy_absorb_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "tnpy_absorb_right": (c_int, [_PD, _PD, c_int, c_int, _PD, c_int, _PD, c_void_p, c_size_t, c_void_p]),
    "tnpy_absorb_left": (c_int, [_PD, _PD, c_int, c_int, _PD, c_int, _PD, c_void_p, c_size_t, c_void_p]),
    "tnpy_mirror_lpr": (c_int, [_PD, _PD, c_int, c_int, c_int, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach the prototypes.  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library has not been built. "
            "Run `python -m tnpy_b200._cuda.build` (needs nvcc); there is no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => header and library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    import os

    choice = os.environ.get("TNPY_GEMM_ALGO", "").lower()
    if choice:
        table = {"auto": GEMM_AUTO, "generic": GEMM_GENERIC, "dmma": GEMM_DMMA, "ozaki": GEMM_OZAKI, "fp64": GEMM_FP64}
        if choice not in table:
            raise RuntimeError(f"TNPY_GEMM_ALGO={choice!r}: expected one of {sorted(table)}")
        lib.tnpy_set_gemm_algo(table[choice])
    return lib


def last_error() -> str:
    return load().tnpy_last_error().decode()


def check(rc: int, what: str, allow_noconv: bool = False) -> int:
    if rc == 0 or (allow_noconv and rc == ENOCONV):
        return rc
    raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")


def launch_count() -> int:
    return int(load().tnpy_launch_count())


def set_gemm_algo(algo: int) -> None:
    """Process-wide GEMM selection of the contraction chains: GEMM_AUTO (default: large products on the tcgen05
    int8 path, FP64-accurate; the rest DMMA / generic), GEMM_FP64 (native FP64 arithmetic only), GEMM_DMMA /
    GEMM_GENERIC (force one kernel).  Also settable through the environment: TNPY_GEMM_ALGO=fp64|dmma|generic|auto."""
    check(load().tnpy_set_gemm_algo(algo), "tnpy_set_gemm_algo")


def set_ozaki_slices(slices: int) -> None:
    check(load().tnpy_set_ozaki_slices(int(slices)), "tnpy_set_ozaki_slices")


def set_fused_steps(on: bool) -> bool:
    """Small sites run whole Lanczos steps in one cooperative launch (csrc/lanczos_steps.cu); False keeps every
    site on the general multi-kernel eigensolver.  Returns the previous setting."""
    return bool(load().tnpy_set_fused_steps(1 if on else 0))


def set_inexact_slices(on: bool) -> bool:
    """The eigensolver's inexact-Krylov slice schedule on the tcgen05 path (fewer int8 slices for the later steps of a
    solve, accepted only on the true residual); False keeps every product at the solve's base slice count.  Returns the
    previous setting."""
    return bool(load().tnpy_set_inexact_slices(1 if on else 0))


HEFF_FP64_CHAIN, HEFF_OZ_CHAIN, HEFF_OZ_DIRECT = 0, 1, 2


# ------------------------------------------------------------------------------------------------
# torch-facing helpers (device memory + stream plumbing only)
# ------------------------------------------------------------------------------------------------
def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors) -> None:
    import torch

    for t in tensors:
        if t is None:
            continue
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise RuntimeError(
                "tnpy_b200 kernels need contiguous float64 CUDA tensors "
                f"(got {type(t).__name__} {getattr(t, 'dtype', None)} on {getattr(t, 'device', None)}); no CPU fallback exists"
            )


class Scratch:
    """Grow-only device workspace owned by the Python side and lent to the library per call: one buffer per
    (device, stream), so that calls enqueued on different streams never share scratch."""

    def __init__(self):
        self._bufs = {}

    def get(self, nbytes: int):
        import torch

        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self._bufs.pop(key, None)  # the caching allocator keeps the old block alive for work already enqueued on this stream
            buf = None
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
            self._bufs[key] = buf
        return buf

    def release(self):
        """Drop every workspace (they are re-created on demand)."""
        self._bufs.clear()


_scratch = Scratch()


def gemm_tn(a, b, out=None, accumulate: bool = False, algo: int = GEMM_AUTO):
    """out[m, n] (+)= sum_k a[k, m] * b[k, n]; a: (K, M), b: (K, N)."""
    import torch

    _need_cuda(a, b, out)
    k, m = a.shape
    k2, n = b.shape
    assert k == k2
    if out is None:
        out = torch.empty((m, n), dtype=torch.float64, device=a.device)
    rc = load().tnpy_gemm_tn(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0), m, n, k,
                             int(accumulate), algo, _stream())
    check(rc, "tnpy_gemm_tn")
    return out


def ozaki_gemm_tn(a, b, out=None, slices: int = 8, accumulate: bool = False, phase: int = 0):
    """EXPERIMENT: out[m, n] (+)= sum_k a[k, m] b[k, n] in FP64 accuracy on the int8 tensor cores."""
    import torch

    _need_cuda(a, b, out)
    k, m = a.shape
    n = b.shape[1]
    if out is None:
        out = torch.empty((m, n), dtype=torch.float64, device=a.device)
    lib = load()
    nbytes = lib.tnpy_ozaki_workspace_bytes(m, n, k, slices)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_ozaki_gemm_tn(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0), m, n, k, int(slices),
                                int(accumulate), int(phase), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_ozaki_gemm_tn")
    return out


def ozaki_error_bound(m: int, n: int, k: int, slices: int = 8) -> float:
    """Rigorous bound on ||C - A^T B||_F for the operands the last ``ozaki_gemm_tn`` call sliced (same m, n, k)."""
    import torch

    lib = load()
    nbytes = lib.tnpy_ozaki_workspace_bytes(m, n, k, slices)
    ws = _scratch.get(nbytes)
    out = torch.empty((), dtype=torch.float64, device="cuda")
    check(lib.tnpy_ozaki_error_bound(m, n, k, int(slices), _ptr(ws), nbytes, _ptr(out), _stream()), "tnpy_ozaki_error_bound")
    return float(out.item())


class HeffPlan:
    """Prepared H_eff of one site (``tnpy_heff_plan_*``): everything that depends only on (L, W, R) -- on the tcgen05
    path the int8 slices of the environments -- is computed once into a torch-owned buffer; ``apply`` then runs one
    matvec.  The plan keeps L, W, R alive and must not outlive changes to them."""

    def __init__(self, L, W, R, l: int, r: int, flags: int = 0, algo: int = GEMM_AUTO, w_host=None, l_rows=None,
                 row0: int = 0):
        """``l_rows`` / ``row0``: L holds only the bra rows row0 .. row0 + l_rows - 1, (l, wl, l_rows) contiguous -- one
        rank's block of the chi-sharded matvec; ``apply`` then maps the full x (l, d, r) to y_rows (l_rows, d, r)."""
        import numpy as np
        import torch

        _need_cuda(L, W, R)
        self._lib = load()
        self._keep = (L, W, R)
        wl, wr, d = W.shape[0], W.shape[1], W.shape[2]
        self.dims = (l, r, wl, wr, d)
        self.l_rows = l if l_rows is None else int(l_rows)
        nbytes = self._lib.tnpy_heff_plan_bytes(l, r, wl, wr, d)
        self._memory = torch.empty(int(nbytes), dtype=torch.uint8, device=W.device)
        self._ws_bytes = self._lib.tnpy_heff_workspace_bytes(l, r, wl, wr, d)
        handle = c_void_p()
        wh = None
        if w_host is not None:
            self._w_host = np.ascontiguousarray(w_host, dtype=np.float64)
            wh = self._w_host.ctypes.data_as(ctypes.POINTER(c_double))
        if l_rows is not None:
            rc = self._lib.tnpy_heff_plan_create_rows(ctypes.byref(handle), _ptr(L), _ptr(W), _ptr(R), wh, l, int(row0),
                                                      self.l_rows, r, wl, wr, d, int(flags), int(algo),
                                                      _ptr(self._memory), nbytes, _stream())
        else:
            rc = self._lib.tnpy_heff_plan_create(ctypes.byref(handle), _ptr(L), _ptr(W), _ptr(R), wh, l, r, wl, wr, d,
                                                 int(flags), int(algo), _ptr(self._memory), nbytes, _stream())
        check(rc, "tnpy_heff_plan_create")
        self._handle = handle

    @property
    def mode(self) -> int:
        return int(self._lib.tnpy_heff_plan_mode(self._handle))

    def apply(self, x, out=None, slices: int = 0):
        import torch

        _need_cuda(x, out)
        l, r, wl, wr, d = self.dims
        if x.numel() != l * d * r:
            raise ValueError(f"HeffPlan.apply: x has {x.numel()} elements, the site has {l * d * r}")
        if out is None:
            out = torch.empty((self.l_rows, d, r), dtype=torch.float64, device=x.device)
        ws = _scratch.get(self._ws_bytes)
        rc = self._lib.tnpy_heff_plan_apply(self._handle, _ptr(x), _ptr(out), int(slices), _ptr(ws), self._ws_bytes, _stream())
        check(rc, "tnpy_heff_plan_apply")
        return out

    def error_bound(self) -> float:
        import torch

        out = torch.empty((), dtype=torch.float64, device="cuda")
        check(self._lib.tnpy_heff_plan_error_bound(self._handle, _ptr(out), _stream()), "tnpy_heff_plan_error_bound")
        return float(out.item())

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._lib.tnpy_heff_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _dims(x_shape, w_shape):
    l, d, r = x_shape
    wl, wr = w_shape[0], w_shape[1]
    return l, r, wl, wr, d


def heff_apply(L, W, R, x, out=None, flags: int = 0):
    """y = H_eff x with x (l, d, r), W (wl, wr, d, d), L (l, wl, l) or None, R (r, wr, r) or None.
    ``flags``: LEFT_IDENTITY / RIGHT_IDENTITY canonical-gauge shortcuts (the caller vouches for them)."""
    import torch

    _need_cuda(L, W, R, x, out)
    l, r, wl, wr, d = _dims(x.shape, W.shape)
    if out is None:
        out = torch.empty_like(x)
    lib = load()
    nbytes = lib.tnpy_heff_workspace_bytes(l, r, wl, wr, d)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_heff_apply(_ptr(L), _ptr(W), _ptr(R), _ptr(x), _ptr(out), l, r, wl, wr, d, int(flags), _ptr(ws), nbytes,
                             _stream())
    check(rc, "tnpy_heff_apply")
    return out


def heff_apply_rows(L_rows, W, R, x, out=None, row0: int = 0, flags: int = 0):
    """Row block of H_eff x: L_rows = L[:, :, row0 : row0 + l_rows] contiguous as (l, wl, l_rows), x (l, d, r) full
    -> (l_rows, d, r).  ``flags`` as for :func:`heff_apply`."""
    import torch

    _need_cuda(L_rows, W, R, x, out)
    l, r, wl, wr, d = _dims(x.shape, W.shape)
    lo = L_rows.shape[2]
    if out is None:
        out = torch.empty((lo, d, r), dtype=torch.float64, device=x.device)
    lib = load()
    nbytes = lib.tnpy_heff_workspace_bytes(l, r, wl, wr, d)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_heff_apply_rows(_ptr(L_rows), _ptr(W), _ptr(R), _ptr(x), _ptr(out), l, int(row0), lo, r, wl, wr, d,
                                  int(flags), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_heff_apply_rows")
    return out


def identity_defect(E, channel: int) -> float:
    """max |E[:, channel, :] - I| of an environment tensor (dim, w, dim); one scalar read-back."""
    import torch

    _need_cuda(E)
    dim, w = E.shape[0], E.shape[1]
    out = torch.empty((), dtype=torch.float64, device=E.device)
    ws = _scratch.get(16384)
    check(load().tnpy_identity_defect(_ptr(E), dim, w, int(channel), _ptr(out), _ptr(ws), 16384, _stream()),
          "tnpy_identity_defect")
    return float(out.item())


def env_update_left(L, A, W, out=None, flags: int = 0):
    import torch

    _need_cuda(L, A, W, out)
    l, r, wl, wr, d = _dims(A.shape, W.shape)
    if out is None:
        out = torch.empty((r, wr, r), dtype=torch.float64, device=A.device)
    lib = load()
    nbytes = lib.tnpy_env_workspace_bytes(l, r, wl, wr, d)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_env_update_left(_ptr(L), _ptr(A), _ptr(W), _ptr(out), l, r, wl, wr, d, int(flags), _ptr(ws), nbytes,
                                  _stream())
    check(rc, "tnpy_env_update_left")
    return out


def env_update_left_rows(L_rows, A, W, row0: int, out=None, flags: int = 0):
    """One row block's contribution to update_left: L_rows (l, wl, l_rows) = L[:, :, row0 : row0 + l_rows], A full."""
    import torch

    _need_cuda(L_rows, A, W, out)
    l, r, wl, wr, d = _dims(A.shape, W.shape)
    lo = 1 if L_rows is None else L_rows.shape[2]
    if out is None:
        out = torch.empty((r, wr, r), dtype=torch.float64, device=A.device)
    lib = load()
    nbytes = lib.tnpy_env_workspace_bytes(l, r, wl, wr, d)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_env_update_left_rows(_ptr(L_rows), _ptr(A), _ptr(W), _ptr(out), l, int(row0), lo, r, wl, wr, d, int(flags),
                                       _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_env_update_left_rows")
    return out


def env_update_right(R, A, W, out=None, flags: int = 0):
    import torch

    _need_cuda(R, A, W, out)
    l, r, wl, wr, d = _dims(A.shape, W.shape)
    if out is None:
        out = torch.empty((l, wl, l), dtype=torch.float64, device=A.device)
    lib = load()
    nbytes = lib.tnpy_env_workspace_bytes(l, r, wl, wr, d)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_env_update_right(_ptr(R), _ptr(A), _ptr(W), _ptr(out), l, r, wl, wr, d, int(flags), _ptr(ws), nbytes,
                                   _stream())
    check(rc, "tnpy_env_update_right")
    return out


def heff_dense(L, W, R, l, r):
    import torch

    _need_cuda(L, W, R)
    wl, wr, d = W.shape[0], W.shape[1], W.shape[2]
    n = l * d * r
    out = torch.empty((n, n), dtype=torch.float64, device=W.device)
    nbytes = load().tnpy_heff_dense_workspace_bytes(l, r, wl, wr, d)
    ws = _scratch.get(nbytes)
    rc = load().tnpy_heff_dense(_ptr(L), _ptr(W), _ptr(R), _ptr(out), l, r, wl, wr, d, _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_heff_dense")
    return out


def eig_lowest(L, W, R, psi, tol: float = 1e-8, max_matvec: int = 1000, ncv: int = 0, flags: int = 0, image=None):
    """In-place: psi (l, d, r) holds v0 on entry and the eigenvector on return.
    Returns dict(theta, resid, n_matvec, n_restart, converged, anorm).  ``image``: an (l, d, r) tensor that
    receives H_eff psi for the returned psi (tnpy_eig_lowest_image: from the Lanczos relation, no extra matvec)."""
    _need_cuda(L, W, R, psi)
    l, r, wl, wr, d = _dims(psi.shape, W.shape)
    lib = load()
    nbytes = lib.tnpy_eig_workspace_bytes(l, r, wl, wr, d, ncv)
    ws = _scratch.get(nbytes)
    stats = (c_double * 8)()
    if image is None:
        rc = lib.tnpy_eig_lowest(_ptr(L), _ptr(W), _ptr(R), _ptr(psi), l, r, wl, wr, d, int(flags), float(tol),
                                 int(max_matvec), int(ncv), stats, _ptr(ws), nbytes, _stream())
    else:
        _need_cuda(image)
        if tuple(image.shape) != tuple(psi.shape) or not image.is_contiguous():
            raise ValueError("eig_lowest: image must be a contiguous tensor of psi's shape")
        rc = lib.tnpy_eig_lowest_image(_ptr(L), _ptr(W), _ptr(R), _ptr(psi), _ptr(image), l, r, wl, wr, d, int(flags),
                                       float(tol), int(max_matvec), int(ncv), stats, _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_eig_lowest", allow_noconv=True)
    counters = last_eig_counters()
    return {
        "theta": stats[0], "resid": stats[1], "n_matvec": int(stats[2]), "n_restart": int(stats[3]),
        "converged": bool(stats[4]), "anorm": stats[5], "int8_error_bound": stats[6],
        "heff_mode": int(stats[7]) // 10, "slices": int(stats[7]) % 10,
        "looks": counters["looks"], "extra_gs_passes": counters["extra_gs_passes"],
        "reduced_slice_matvecs": counters["reduced_slice_matvecs"],
        "failed_residual_checks": counters["failed_residual_checks"],
        "five_slice_matvecs": counters["five_slice_matvecs"],
    }


COMM_ID_BYTES = 128


class Comm:
    """NCCL communicator owned by the library (``tnpy_comm_*``), one per process.  ``Comm.from_torch_distributed()``
    creates it on the current CUDA device from an initialised ``torch.distributed`` group of any backend (the
    128-byte id travels by broadcast)."""

    def __init__(self, unique_id: bytes, world: int, rank: int):
        self._lib = load()
        handle = c_void_p()
        buf = ctypes.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        check(self._lib.tnpy_comm_create(ctypes.byref(handle), buf, int(world), int(rank)), "tnpy_comm_create")
        self._handle, self.world, self.rank = handle, int(world), int(rank)

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(COMM_ID_BYTES)
        check(load().tnpy_comm_unique_id(buf), "tnpy_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, group=None) -> "Comm":
        import torch
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        torch.cuda.synchronize()
        return cls(box[0], world, rank)

    @property
    def handle(self):
        return self._handle

    def allgather(self, send, recv):
        _need_cuda(send, recv)
        check(self._lib.tnpy_comm_allgather(self._handle, _ptr(send), _ptr(recv), send.numel(), _stream()), "tnpy_comm_allgather")
        return recv

    def allreduce_sum(self, buf):
        _need_cuda(buf)
        check(self._lib.tnpy_comm_allreduce_sum(self._handle, _ptr(buf), buf.numel(), _stream()), "tnpy_comm_allreduce_sum")
        return buf

    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._lib.tnpy_comm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def eig_lowest_rows(comm: Comm, L_rows, W, R, psi_rows, l: int, row0: int, tol: float = 1e-8, max_matvec: int = 1000,
                    ncv: int = 0, flags: int = 0, image_rows=None):
    """Row-sharded ``eig_lowest`` (collective over ``comm``): in place on this rank's rows of psi, (l_rows, d, r).
    Returns the same dict as :func:`eig_lowest` (identical on every rank)."""
    _need_cuda(L_rows, W, R, psi_rows, image_rows)
    lo, d, r = psi_rows.shape
    wl, wr = W.shape[0], W.shape[1]
    lib = load()
    nbytes = lib.tnpy_eig_rows_workspace_bytes(l, lo, r, wl, wr, d, ncv)
    ws = _scratch.get(nbytes)
    stats = (c_double * 8)()
    rc = lib.tnpy_eig_lowest_rows(comm.handle, _ptr(L_rows), _ptr(W), _ptr(R), _ptr(psi_rows), _ptr(image_rows), l, int(row0),
                                  lo, r, wl, wr, d, int(flags), float(tol), int(max_matvec), int(ncv), stats, _ptr(ws),
                                  nbytes, _stream())
    check(rc, "tnpy_eig_lowest_rows", allow_noconv=True)
    return {
        "theta": stats[0], "resid": stats[1], "n_matvec": int(stats[2]), "n_restart": int(stats[3]),
        "converged": bool(stats[4]), "anorm": stats[5], "int8_error_bound": stats[6],
        "heff_mode": int(stats[7]) // 10, "slices": int(stats[7]) % 10,
    }


def last_eig_counters() -> dict:
    """Diagnostics of this thread's last on-device eigensolve: matvecs, looks, extra Gram-Schmidt passes, restarts."""
    buf = (c_int64 * 7)()
    load().tnpy_last_eig_counters(buf, 7)
    return {"n_matvec": buf[0], "looks": buf[1], "extra_gs_passes": buf[2], "restarts": buf[3],
            "reduced_slice_matvecs": buf[4], "failed_residual_checks": buf[5], "five_slice_matvecs": buf[6]}


def geig_lowest(LA, WA, RA, LM, WM, RM, psi, tol: float = 1e-8, max_iter: int = 2000, ncv: int = 0, flags_a: int = 0):
    """Lowest eigenpair of A x = lambda M x (A, M = H_eff of two environments at the same site).
    In-place on psi (l, d, r).  Returns dict(theta, resid, n_iter, n_restart, converged)."""
    _need_cuda(LA, WA, RA, LM, WM, RM, psi)
    l, d, r = psi.shape
    wla, wra, wlm, wrm = WA.shape[0], WA.shape[1], WM.shape[0], WM.shape[1]
    lib = load()
    nbytes = lib.tnpy_geig_workspace_bytes(l, r, wla, wra, wlm, wrm, d, ncv)
    ws = _scratch.get(nbytes)
    stats = (c_double * 8)()
    rc = lib.tnpy_geig_lowest(_ptr(LA), _ptr(WA), _ptr(RA), _ptr(LM), _ptr(WM), _ptr(RM), _ptr(psi), l, r, wla, wra,
                              wlm, wrm, d, int(flags_a), float(tol), int(max_iter), int(ncv), stats, _ptr(ws), nbytes,
                              _stream())
    check(rc, "tnpy_geig_lowest", allow_noconv=True)
    return {"theta": stats[0], "resid": stats[1], "n_iter": int(stats[2]), "n_restart": int(stats[3]),
            "converged": bool(stats[4])}


def geig_dense_lowest(a, b):
    """Lowest eigenpair of the dense pencil a x = lambda b x (both destroyed).  Returns (theta 0-d, x)."""
    import torch

    _need_cuda(a, b)
    n = a.shape[0]
    lib = load()
    nbytes = lib.tnpy_geig_dense_workspace_bytes(n)
    ws = _scratch.get(nbytes)
    theta = torch.empty((), dtype=torch.float64, device=a.device)
    x = torch.empty(n, dtype=torch.float64, device=a.device)
    check(lib.tnpy_geig_dense_lowest(_ptr(a), _ptr(b), n, _ptr(theta), _ptr(x), _ptr(ws), nbytes, _stream()),
          "tnpy_geig_dense_lowest")
    return theta, x


def geig_chol_lowest(a, b, tol: float = 1e-12, max_matvec: int = 2000):
    """Lowest eigenpair of the dense pencil a x = lambda b x through a Cholesky factor of b and the on-device
    Lanczos solver (a, b left intact; thousands to tens of thousands of unknowns).  Returns (theta 0-d, x, stats)
    with x^T b x = 1; raises RuntimeError when b is not positive definite to working precision."""
    import torch

    _need_cuda(a, b)
    n = a.shape[0]
    lib = load()
    nbytes = lib.tnpy_geig_chol_workspace_bytes(n)
    ws = _scratch.get(nbytes)
    theta = torch.empty((), dtype=torch.float64, device=a.device)
    x = torch.empty(n, dtype=torch.float64, device=a.device)
    stats = (c_double * 8)()
    rc = lib.tnpy_geig_chol_lowest(_ptr(a), _ptr(b), n, float(tol), int(max_matvec), _ptr(theta), _ptr(x), stats,
                                   _ptr(ws), nbytes, _stream())
    if rc == ENOCONV and "positive definite" in last_error():
        raise RuntimeError(f"tnpy_geig_chol_lowest failed (code {rc}): {last_error()}")
    check(rc, "tnpy_geig_chol_lowest", allow_noconv=True)
    return theta, x, {"theta": stats[0], "resid": stats[1], "n_matvec": int(stats[2]), "n_restart": int(stats[3]),
                      "converged": bool(stats[4]), "anorm": stats[5]}


def eigh_lowest(H):
    """Lowest eigenpair of a dense symmetric matrix (destroyed).  Returns (eval 0-d tensor, evec)."""
    import torch

    _need_cuda(H)
    n = H.shape[0]
    lib = load()
    nbytes = lib.tnpy_eigh_workspace_bytes(n)
    ws = _scratch.get(nbytes)
    ev = torch.empty((), dtype=torch.float64, device=H.device)
    vec = torch.empty(n, dtype=torch.float64, device=H.device)
    rc = lib.tnpy_eigh_lowest(_ptr(H), n, _ptr(ev), _ptr(vec), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_eigh_lowest")
    return ev, vec


def svd(A):
    """Thin SVD of a 2-D tensor (destroyed).  Returns (U, s, Vt), s descending."""
    import torch

    _need_cuda(A)
    rows, cols = A.shape
    k = min(rows, cols)
    U = torch.empty((rows, k), dtype=torch.float64, device=A.device)
    s = torch.empty(k, dtype=torch.float64, device=A.device)
    Vt = torch.empty((k, cols), dtype=torch.float64, device=A.device)
    lib = load()
    nbytes = lib.tnpy_svd_workspace_bytes(rows, cols)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_svd(_ptr(A), rows, cols, _ptr(U), _ptr(s), _ptr(Vt), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_svd")
    return U, s, Vt


def qr_split(A, shifted: bool = False, t_first: bool = False):
    """Orthogonal split of a 2-D tensor (not modified): ``A = Q @ T`` (rows >= cols, Q with orthonormal
    columns) or ``A = T @ Q`` (rows < cols, Q with orthonormal rows).  Returns (Q, T, defect) with
    defect = max|Q^T Q - I| measured on the device (inf when the Cholesky factorisation broke down); the
    caller decides whether to accept the split.  ``shifted``: the three-pass variant for ill-conditioned
    vectors (TNPY_QR_SHIFTED).  ``t_first``: a *square* A is factorised as ``T @ Q`` instead of ``Q @ T``
    (TNPY_QR_T_FIRST; the leftward split of a square site tensor) -- ignored otherwise."""
    import torch

    _need_cuda(A)
    rows, cols = A.shape
    k = min(rows, cols)
    Q = torch.empty((rows, cols), dtype=torch.float64, device=A.device)
    T = torch.empty((k, k), dtype=torch.float64, device=A.device)
    defect = torch.empty(1, dtype=torch.float64, device=A.device)
    lib = load()
    nbytes = lib.tnpy_qr_split_workspace_bytes(rows, cols)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_qr_split(_ptr(A), rows, cols, _ptr(Q), _ptr(T), _ptr(defect), (1 if shifted else 0) | (2 if t_first else 0), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_qr_split")
    return Q, T, float(defect.item())


def absorb_right(s, Vt, nb):
    """diag(s) Vt @ nb  with nb (n, cols)."""
    import torch

    _need_cuda(s, Vt, nb)
    k, n = Vt.shape
    cols = nb.shape[1]
    out = torch.empty((k, cols), dtype=torch.float64, device=nb.device)
    lib = load()
    nbytes = lib.tnpy_absorb_workspace_bytes(k, n, cols)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_absorb_right(_ptr(s), _ptr(Vt), k, n, _ptr(nb), cols, _ptr(out), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_absorb_right")
    return out


def absorb_left(U, s, nb):
    """nb @ U diag(s)  with nb (rows, n)."""
    import torch

    _need_cuda(U, s, nb)
    n, k = U.shape
    rows = nb.shape[0]
    out = torch.empty((rows, k), dtype=torch.float64, device=nb.device)
    lib = load()
    nbytes = lib.tnpy_absorb_workspace_bytes(k, n, rows)
    ws = _scratch.get(nbytes)
    rc = lib.tnpy_absorb_left(_ptr(U), _ptr(s), n, k, _ptr(nb), rows, _ptr(out), _ptr(ws), nbytes, _stream())
    check(rc, "tnpy_absorb_left")
    return out


def mirror_lpr(a):
    import torch

    _need_cuda(a)
    l, d, r = a.shape
    out = torch.empty((r, d, l), dtype=torch.float64, device=a.device)
    check(load().tnpy_mirror_lpr(_ptr(a), _ptr(out), l, d, r, _stream()), "tnpy_mirror_lpr")
    return out


def dot(x, y):
    import torch

    _need_cuda(x, y)
    out = torch.empty((), dtype=torch.float64, device=x.device)
    check(load().tnpy_dot(_ptr(x), _ptr(y), x.numel(), _ptr(out), _stream()), "tnpy_dot")
    return out


def nrm2(x):
    import torch

    _need_cuda(x)
    out = torch.empty((), dtype=torch.float64, device=x.device)
    check(load().tnpy_nrm2(_ptr(x), x.numel(), _ptr(out), _stream()), "tnpy_nrm2")
    return out


def axpy(alpha: float, x, y):
    _need_cuda(x, y)
    check(load().tnpy_axpy(float(alpha), _ptr(x), _ptr(y), x.numel(), _stream()), "tnpy_axpy")
    return y


def scal(alpha: float, x):
    _need_cuda(x)
    check(load().tnpy_scal(float(alpha), _ptr(x), x.numel(), _stream()), "tnpy_scal")
    return x


def multi_dot(V, w, m: Optional[int] = None):
    import torch

    _need_cuda(V, w)
    m = V.shape[0] if m is None else m
    out = torch.empty(m, dtype=torch.float64, device=w.device)
    check(load().tnpy_multi_dot(_ptr(V), V.stride(0), m, _ptr(w), w.numel(), _ptr(out), _stream()), "tnpy_multi_dot")
    return out


def multi_axpy(V, h, w, m: Optional[int] = None):
    _need_cuda(V, h, w)
    m = V.shape[0] if m is None else m
    check(load().tnpy_multi_axpy(_ptr(V), V.stride(0), m, _ptr(h), _ptr(w), w.numel(), _stream()), "tnpy_multi_axpy")
    return w
