"""MPS container (host) and the device-resident environment cache of the fDMRG local update.

Mirrors tnpy/matrix_product_state.py: ``Direction`` (:19-25), ``MatrixProductState`` (:28-228, arrays in
'lpr' order), ``Environment`` (:231-440) and ``MatrixProductStateMeasurements`` (:443-453).

What lives where
  * ``MatrixProductState`` holds host float64 arrays and is what ``FiniteDMRG.mps`` hands back
    (``to_quimb()`` converts when quimb is importable; quimb is not a dependency of this package).
  * ``Environment`` keeps the site tensors, the MPO tensors and the L / R stacks as float64 CUDA
    tensors for its whole life and does every contraction through the C ABI (``tnpy_b200._cuda``):
    ``update_left/right`` -> ``tnpy_env_update_*``, ``one_site_matvec`` -> ``tnpy_heff_apply``,
    ``split_tensor`` -> ``tnpy_svd`` + ``tnpy_absorb_*``.  Real data: the bra copy the reference
    keeps (``_conj_mps``, :243) is the ket itself.
"""
from __future__ import annotations

from enum import Enum
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
from scipy.sparse.linalg import LinearOperator

from tnpy_b200 import _cuda, logger
from tnpy_b200.operators import MatrixProductOperator


class Direction(Enum):
    RIGHTWARD = 1
    LEFTWARD = -1


def compressed_bond_dims(n: int, bond_dim: int, phys_dim: int) -> List[int]:
    """Bond dimensions after quimb's ``compress()``: min(d^i, chi, d^(n-i)) (pinned by the
    reference's tests/test_matrix_product_state.py:13-37)."""
    return [int(min(phys_dim ** (i + 1), bond_dim, phys_dim ** (n - 1 - i))) for i in range(n - 1)]


class SiteTensor:
    """The slice of quimb's Tensor API the reference touches on a site: data/shape/size/inds/tags/modify."""

    def __init__(self, data: np.ndarray, inds: Sequence[str], tags: Sequence[str]):
        self.data = data
        self.inds = tuple(inds)
        self.tags = set(tags)

    shape = property(lambda self: self.data.shape)
    size = property(lambda self: self.data.size)

    def modify(self, data: np.ndarray):
        self.data = np.asarray(data, dtype=float)


class MatrixProductState:
    """Open-boundary MPS, per-site arrays in 'lpr' order: (d, r) / (l, d, r) / (l, d)."""

    def __init__(self, arrays: Sequence[np.ndarray], shape: str = "lpr"):
        arrays = [np.array(a, dtype=float) for a in arrays]
        n = len(arrays)
        if n < 2:
            raise ValueError("An MPS needs at least two sites.")
        if shape != "lpr":  # accept quimb's default 'lrp' too
            arrays = [self._to_lpr(a, shape, i, n) for i, a in enumerate(arrays)]
        self._tensors: List[SiteTensor] = []
        for i, a in enumerate(arrays):
            inds = ([f"_bond{i - 1}"] if i > 0 else []) + [f"k{i}"] + ([f"_bond{i}"] if i < n - 1 else [])
            self._tensors.append(SiteTensor(a, inds, [f"I{i}"]))

    @staticmethod
    def _to_lpr(a: np.ndarray, shape: str, site: int, n: int) -> np.ndarray:
        have = [s for s in shape if not ((s == "l" and site == 0) or (s == "r" and site == n - 1))]
        want = [s for s in "lpr" if s in have]
        return np.transpose(a, [have.index(s) for s in want])

    # -- bookkeeping ------------------------------------------------------------------------------
    @property
    def n_sites(self) -> int:
        return len(self._tensors)

    nsites = n_sites
    L = n_sites

    @property
    def phys_dim(self) -> int:
        return self._tensors[0].shape[0]

    @property
    def bond_dim(self) -> int:
        return self.max_bond()

    def max_bond(self) -> int:
        return max(t.shape[-1] for t in self._tensors[:-1])

    def bond_dims(self) -> List[int]:
        return [t.shape[-1] for t in self._tensors[:-1]]

    def site_tag(self, site: int) -> str:
        return f"I{site}"

    def site_ind(self, site: int) -> str:
        return f"k{site}"

    def __len__(self) -> int:
        return self.n_sites

    def __getitem__(self, site: int) -> SiteTensor:
        return self._tensors[site]

    def __iter__(self):
        return iter(self._tensors)

    @property
    def arrays(self) -> List[np.ndarray]:
        return [t.data for t in self._tensors]

    def three_leg(self, site: int) -> np.ndarray:
        a = self._tensors[site].data
        if a.ndim == 3:
            return a
        return a[None] if site == 0 else a[:, :, None]

    def copy(self) -> "MatrixProductState":
        return MatrixProductState([a.copy() for a in self.arrays])

    def conj(self, mangle_inner: bool = False, mangle_outer: bool = False) -> "MatrixProductState":
        """Real tensors: a copy (the reference renames indices only, :93-116)."""
        out = self.copy()
        if mangle_outer:
            for i, t in enumerate(out._tensors):
                t.inds = tuple(f"b{i}" if ind == f"k{i}" else ind for ind in t.inds)
        return out

    # -- constructors -------------------------------------------------------------------------------
    @classmethod
    def random(cls, n: int, bond_dim: int, phys_dim: int, seed: Optional[int] = None, **kwargs) -> "MatrixProductState":
        """Random normalised right-canonical MPS at the compressed bond dimensions (:170-185).

        The reference draws from quimb's RNG and calls ``compress()``; here ``default_rng(seed)``
        standard normals are right-canonicalised by QR from the last site down, site 0 normalised
        (the same state class; SURVEY 8d fixes this generator for both sides of every parity test).
        """
        rng = np.random.default_rng(seed)
        dims = [1] + compressed_bond_dims(n, bond_dim, phys_dim) + [1]
        arrays = [rng.standard_normal((dims[i], phys_dim, dims[i + 1])) for i in range(n)]
        for site in range(n - 1, 0, -1):
            l, d, r = arrays[site].shape
            q, rr = np.linalg.qr(arrays[site].reshape(l, d * r).T)
            arrays[site] = q.T.reshape(l, d, r)
            arrays[site - 1] = np.einsum("lpr,sr->lps", arrays[site - 1], rr)
        arrays[0] /= np.linalg.norm(arrays[0])
        arrays[0] = arrays[0].reshape(phys_dim, dims[1])
        arrays[-1] = arrays[-1].reshape(dims[n - 1], phys_dim)
        return cls(arrays, **kwargs)

    def save(self, filename: str):
        """'lpr' arrays keyed by site tag, ``.npz`` or ``.hdf5`` (:118-143)."""
        datasets = {self.site_tag(i): t.data for i, t in enumerate(self._tensors)}
        suffix = Path(filename).suffix
        if suffix == ".npz":
            np.savez(str(filename), **datasets)
        elif suffix == ".hdf5":
            import h5py  # optional, as in the reference

            with h5py.File(str(filename), "w") as f:
                for tag, array in datasets.items():
                    f.create_dataset(tag, data=array)
        else:
            raise ValueError(f"File extension {suffix} is not supported.")

    @classmethod
    def load(cls, filename: str) -> "MatrixProductState":
        """Inverse of :meth:`save`; sites are ordered by the integer in their tag (fixes the
        alphabetical-order defect the reference flags at :130/:158)."""
        suffix = Path(filename).suffix
        if suffix == ".npz":
            with np.load(str(filename)) as f:
                items = {k: f[k] for k in f.files}
        elif suffix == ".hdf5":
            import h5py

            with h5py.File(str(filename), "r") as f:
                items = {k: v[()] for k, v in f.items()}
        else:
            raise ValueError(f"File extension {suffix} is not supported.")
        return cls([items[k] for k in sorted(items, key=lambda tag: int(tag[1:]))])

    # -- host-side measurements ---------------------------------------------------------------------
    def overlap(self, other: "MatrixProductState") -> float:
        e = np.ones((1, 1))
        for site in range(self.n_sites):
            e = np.einsum("lm,lpr,mps->rs", e, self.three_leg(site), other.three_leg(site), optimize=True)
        return float(e[0, 0])

    def norm(self) -> float:
        return float(np.sqrt(self.overlap(self)))

    def __matmul__(self, other: "MatrixProductState") -> float:
        return self.overlap(other)

    def to_dense(self) -> np.ndarray:
        acc = self.three_leg(0)[0]
        for site in range(1, self.n_sites):
            acc = np.tensordot(acc, self.three_leg(site), axes=(acc.ndim - 1, 0)).reshape(-1, self.three_leg(site).shape[2])
        return acc.reshape(-1)

    def to_quimb(self):
        import quimb.tensor as qtn  # only if the user has it

        return qtn.MatrixProductState(self.arrays, shape="lpr")

    # -- a9: split_tensor ---------------------------------------------------------------------------
    def split_tensor(self, site: int, direction: Direction):
        """SVD-split ``site`` and push ``s Vt`` / ``U s`` into the neighbour (:187-225); runs on the
        GPU through the same kernels as :meth:`Environment.split_tensor`."""
        import torch

        if direction not in (Direction.RIGHTWARD, Direction.LEFTWARD):
            raise KeyError("MatrixProductState only supplies left or right direction.")
        nb_site = site + 1 if direction == Direction.RIGHTWARD else site - 1
        a = torch.from_numpy(np.ascontiguousarray(self.three_leg(site))).cuda()
        nb = torch.from_numpy(np.ascontiguousarray(self.three_leg(nb_site))).cuda()
        a, nb, _ = _split_on_device(a, nb, direction)
        self[site].modify(data=a.cpu().numpy().reshape(self[site].shape))
        self[nb_site].modify(data=nb.cpu().numpy().reshape(self[nb_site].shape))

    def enlarge_bond_dim(self, new_bond_dim: int, method: str):
        return NotImplemented


#: the verified orthogonality max|Q^T Q - I| a Cholesky-QR split must reach to be accepted (else: Jacobi SVD)
QR_DEFECT_TOL = 1e-13


class DeferredSpectrum:
    """Bond spectrum of a QR split: the square factor T (device) whose singular values are the bond's;
    the small SVD runs the first time the values are asked for."""

    def __init__(self, t, shifted: bool = False):
        self._t, self._s = t, None
        self.shifted = shifted  # the split needed the shifted three-pass factorisation

    def values(self):
        if self._s is None:
            self._s = _cuda.svd(self._t.clone())[1]
            self._t = None
        return self._s

    def cpu(self):
        return self.values().cpu()


def _ones(k: int, like):
    import torch

    return torch.ones(k, dtype=torch.float64, device=like.device)


def _verified_qr(mat, t_first: bool):
    """Cholesky-QR split accepted only on the orthogonality the device measured: two passes first, the
    shifted three-pass variant when those were not enough (ill-conditioned cold-sweep tensors); (None, None,
    False) sends the caller to the Jacobi SVD.  ``t_first`` fixes the order of the factors for a square
    matrix (mat = T Q for a leftward split, Q T for a rightward one)."""
    q, t, defect = _cuda.qr_split(mat, t_first=t_first)
    if defect <= QR_DEFECT_TOL:
        return q, t, False
    q, t, defect3 = _cuda.qr_split(mat, shifted=True, t_first=t_first)
    if defect3 <= QR_DEFECT_TOL:
        return q, t, True
    logger.info(f"split_tensor: Cholesky-QR defects {defect:.1e} / {defect3:.1e} (shifted) on a "
                f"{mat.shape[0]}x{mat.shape[1]} split, using the SVD")
    return None, None, False


def _split_on_device(a, nb, direction: Direction, mode: str = "svd", qr_min_bond: int = 16):
    """Device tensors in, device tensors out: (new site tensor, new neighbour, bond spectrum).

    ``mode="svd"``: the reference's split (linalg.py:9-23 with cutoff = current bond), spectrum = s.
    ``mode="qr"``: the same orthogonalisation as A = Q T / T Q by ``tnpy_qr_split`` for bonds >=
    ``qr_min_bond`` -- the state and every later local problem are unchanged (a bond gauge); the
    spectrum is a :class:`DeferredSpectrum` over T.  The split is accepted only when the orthogonality
    measured on the device is at rounding level, otherwise this call falls back to the SVD."""
    l, d, r = a.shape
    if direction == Direction.RIGHTWARD:
        if l * d < r:
            raise ValueError(f"split_tensor: site tensor {tuple(a.shape)} cannot keep a right bond of {r}")
        r2 = nb.shape[2]
        if mode == "qr" and r >= qr_min_bond:
            q, t, shifted = _verified_qr(a.reshape(l * d, r), t_first=False)
            if q is not None:
                new_nb = _cuda.absorb_right(_ones(r, a), t, nb.reshape(r, d * r2)).reshape(r, d, r2)
                return q.reshape(l, d, r), new_nb, DeferredSpectrum(t, shifted)
        u, s, vt = _cuda.svd(a.reshape(l * d, r).clone())
        new_nb = _cuda.absorb_right(s, vt, nb.reshape(r, d * r2)).reshape(r, d, r2)
        return u.reshape(l, d, r), new_nb, s
    if d * r < l:
        raise ValueError(f"split_tensor: site tensor {tuple(a.shape)} cannot keep a left bond of {l}")
    l0 = nb.shape[0]
    if mode == "qr" and l >= qr_min_bond:
        q, t, shifted = _verified_qr(a.reshape(l, d * r), t_first=True)
        if q is not None:
            new_nb = _cuda.absorb_left(t, _ones(l, a), nb.reshape(l0 * d, l)).reshape(l0, d, l)
            return q.reshape(l, d, r), new_nb, DeferredSpectrum(t, shifted)
    u, s, vt = _cuda.svd(a.reshape(l, d * r).clone())
    new_nb = _cuda.absorb_left(u, s, nb.reshape(l0 * d, l)).reshape(l0, d, l)
    return vt.reshape(l, d, r), new_nb, s


class HeffOperator(LinearOperator):
    """``Environment.one_site_matvec(site)``: a SciPy LinearOperator whose matvec runs on the GPU.

    Host seam (the reference's contract, :411-440): ``matvec(x)`` takes a host vector of length N
    (shape (N,) or (N,1)), copies it to the device, applies H_eff and copies the result back as (N,1) / (N,);
    ``matmat(X)`` moves the whole block once.  ``linalg.eigshmv`` recognises this class and keeps
    the whole eigensolve on the device instead of calling back per matvec.

    The operator is a *prepared* H_eff (``tnpy_heff_plan``): what depends only on (L, W, R) -- on the tcgen05
    path the int8 slices of the environments -- is computed at the first application and reused by every later
    one, as long as the environment at this site is not updated (then it is prepared again).

    ``zero_copy=True`` makes ``matvec`` return a view of one of two alternating pinned buffers (valid until the
    call after next) instead of a fresh array: for timing loops only, never for block solvers.
    """

    def __init__(self, env: "Environment", site: int, zero_copy: bool = False):
        self.env = env
        self.site = site
        self.zero_copy = zero_copy
        self.site_shape = tuple(env.device_tensor(site).shape)
        self._plan = None
        self._plan_version = None
        n = int(np.prod(self.site_shape))
        super().__init__(dtype=np.dtype("float64"), shape=(n, n))

    def plan(self) -> "_cuda.HeffPlan":
        flags = self.env.gauge_flags(self.site)
        version = (self.env.operand_version(self.site), flags)
        if self._plan is None or self._plan_version != version:
            if self._plan is not None:
                self._plan.close()
            L, W, R = self.env.operands(self.site)
            l, _, r = self.site_shape
            self._plan = _cuda.HeffPlan(L, W, R, l, r, flags=flags, w_host=self.env.mpo.as_four_leg(self.site))
            self._plan_version = version
        return self._plan

    def apply_device(self, x, out=None):
        return self.plan().apply(x.reshape(self.site_shape), out)

    def _staging(self):
        import torch

        if getattr(self, "_stage", None) is None:
            n = self.shape[0]
            self._stage = {
                "xd": torch.empty(self.site_shape, dtype=torch.float64, device="cuda"),
                "yd": torch.empty(self.site_shape, dtype=torch.float64, device="cuda"),
                "pin_in": torch.empty(n, dtype=torch.float64).pin_memory(),
                "pin_out": [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)],
                "flip": 0,
            }
        return self._stage

    def _matvec(self, x: np.ndarray) -> np.ndarray:
        """Host vector in, host vector out; both copies go through pinned staging buffers."""
        import torch

        st = self._staging()
        src = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1))
        if not src.is_pinned():
            st["pin_in"].copy_(src)
            src = st["pin_in"]
        st["xd"].reshape(-1).copy_(src, non_blocking=True)
        self.apply_device(st["xd"], st["yd"])
        st["flip"] ^= 1
        out = st["pin_out"][st["flip"]]
        out.copy_(st["yd"].reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out.numpy() if self.zero_copy else out.numpy().copy()

    def _matmat(self, X: np.ndarray) -> np.ndarray:
        """Block application (what primme / lobpcg call): one host-to-device copy of the block, one matvec per
        column on the device, one copy back.  Every column of the result is its own memory."""
        import torch

        cols = np.ascontiguousarray(np.asarray(X, dtype=np.float64).T)  # (k, N)
        xd = torch.from_numpy(cols).cuda()
        yd = torch.empty_like(xd)
        for i in range(xd.shape[0]):
            self.apply_device(xd[i].reshape(self.site_shape), yd[i].reshape(self.site_shape))
        return np.ascontiguousarray(yd.cpu().numpy().T)

    def _adjoint(self):
        return self  # H_eff is real symmetric


class _DeviceDictView:
    """Read-only mapping over an environment stack: ``view[site].data`` is a host copy in the
    reference's (ket, mpo, bra) order, ``view.device(site)`` the CUDA tensor itself."""

    class _Item:
        def __init__(self, t):
            self._t = t
            self.inds = ("ket", "mpo", "bra")

        @property
        def data(self) -> np.ndarray:
            return self._t.cpu().numpy()

        @property
        def shape(self):
            return tuple(self._t.shape)

    def __init__(self, store: Dict[int, object]):
        self._store = store

    def __getitem__(self, site: int):
        return self._Item(self._store[site])

    def __contains__(self, site: int) -> bool:
        return site in self._store

    def __len__(self) -> int:
        return len(self._store)

    def keys(self):
        return self._store.keys()

    def device(self, site: int):
        return self._store[site]


class Environment:
    #: an environment channel counts as the identity when max|E[:, c, :] - I| is below this
    IDENTITY_TOL = 1e-12

    def __init__(self, mpo: MatrixProductOperator, mps, build_left: bool = True, use_identity_channels: bool = True,
                 share_state_with: Optional["Environment"] = None, canonicalize: bool = False,
                 split: str = "qr", qr_min_bond: int = 16):
        """``mps`` is a :class:`MatrixProductState` (host, as in the reference) or a list of
        three-leg (l, d, r) float64 CUDA tensors (device-born synthetic states for benchmarks).
        ``share_state_with``: a second environment over the *same* MPS (ShiftInvertDMRG's H^2
        environment) borrows the first one's device site tensors instead of copying them."""
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("tnpy_b200.Environment needs a CUDA device; there is no CPU fallback.")
        _cuda.load()
        if split not in ("qr", "svd"):
            raise ValueError(f"split must be 'qr' or 'svd', got {split!r}")
        #: how split_tensor orthogonalises: "qr" (Cholesky-QR for bonds >= qr_min_bond, small SVD deferred
        #: until the bond spectrum is read, SVD fallback when the verified orthogonality is not at rounding
        #: level) or "svd" (the reference's literal gauge, one Jacobi SVD per split)
        self.split_mode, self.qr_min_bond = split, qr_min_bond
        self.split_counts = {"qr": 0, "qr_shifted": 0, "svd": 0}
        self._mpo = mpo
        self._n_sites = mpo.nsites
        if len(mps) != self._n_sites:
            raise ValueError("MPO and MPS have different lengths.")
        if share_state_with is not None:
            other = share_state_with
            self._A, self._shapes = other._A, other._shapes  # same list objects: updates are seen by both
            self._host_mps, self._dirty = other._host_mps, other._dirty
        elif isinstance(mps, MatrixProductState):
            self._shapes = [tuple(t.shape) for t in mps]
            self._A = [torch.from_numpy(np.ascontiguousarray(mps.three_leg(i))).cuda() for i in range(self._n_sites)]
            self._host_mps = mps
            self._dirty = set()
        else:
            self._A = [t.to(device="cuda", dtype=torch.float64).contiguous() for t in mps]
            n = self._n_sites
            self._shapes = [
                tuple(t.shape[1:]) if i == 0 else (tuple(t.shape[:2]) if i == n - 1 else tuple(t.shape))
                for i, t in enumerate(self._A)
            ]
            self._host_mps = None
            self._dirty = set(range(n))
        self._W = [torch.from_numpy(np.ascontiguousarray(mpo.as_four_leg(i))).cuda() for i in range(self._n_sites)]
        self._W2 = None
        self._left: Dict[int, object] = {}
        self._right: Dict[int, object] = {}
        # canonical-gauge shortcuts: is L[site][:, 0, :] / R[site][:, -1, :] the identity?  Measured on the
        # device after every environment update (one scalar read-back), never assumed.
        self._use_identity = use_identity_channels
        #: runtime switch over the measured shortcuts (benchmarks time the general chain by clearing it)
        self.use_identity_channels = use_identity_channels
        self._left_identity: Dict[int, bool] = {}
        self._right_identity: Dict[int, bool] = {}
        self.bond_singular_values: Dict[int, object] = {}
        self._image = None  # (site, H_eff psi) left by the last on-device eigensolve, valid until the site changes
        self._versions: Dict[tuple, int] = {}  # ("L" | "R", site) -> number of times that environment was rewritten
        if canonicalize and share_state_with is None:
            self.right_canonicalize()
        if build_left:  # the reference builds both stacks up front (:247-250)
            for site in range(1, self.n_sites):
                self.update_left(site)
        for site in range(self.n_sites - 2, -1, -1):
            self.update_right(site)

    def right_canonicalize(self):
        """Bring a user-supplied MPS into the right-canonical form the solver assumes (the reference
        relies on ``MatrixProductState.random`` being right-canonical and does not canonicalise a
        user ``mps=``, finite_dmrg.py:67-74).  Device SVD splits from the last site down."""
        for site in range(self.n_sites - 1, 0, -1):
            a, nb, _ = _split_on_device(self._A[site], self._A[site - 1], Direction.LEFTWARD)
            self._A[site], self._A[site - 1] = a.contiguous(), nb.contiguous()
            self._dirty.update((site, site - 1))

    def close(self):
        logger.info("Deleting left and right environments.")
        self._left, self._right = {}, {}

    # -- accessors ----------------------------------------------------------------------------------
    @property
    def mpo(self) -> MatrixProductOperator:
        return self._mpo

    @property
    def n_sites(self) -> int:
        return self._n_sites

    @property
    def mps(self) -> MatrixProductState:
        """Host MPS, refreshed from the device for every site changed since the last access."""
        if self._host_mps is None:
            self._host_mps = MatrixProductState(
                [a.cpu().numpy().reshape(shape) for a, shape in zip(self._A, self._shapes)]
            )
            self._dirty.clear()
        for site in sorted(self._dirty):
            self._host_mps[site].modify(data=self._A[site].cpu().numpy().reshape(self._shapes[site]))
        self._dirty.clear()
        return self._host_mps

    @property
    def left(self) -> _DeviceDictView:
        return _DeviceDictView(self._left)

    @property
    def right(self) -> _DeviceDictView:
        return _DeviceDictView(self._right)

    def device_tensor(self, site: int):
        return self._A[site]

    def site_shape(self, site: int):
        return self._shapes[site]

    def operands(self, site: int):
        """(L, W, R) device operands of H_eff at ``site`` (None = unit boundary)."""
        return self._left.get(site) if site > 0 else None, self._W[site], (
            self._right.get(site) if site < self.n_sites - 1 else None
        )

    def operand_version(self, site: int):
        """Changes whenever L[site] or R[site] is rewritten (environment tensors are updated in place, so a prepared
        operator must notice)."""
        return self._versions.get(("L", site), 0), self._versions.get(("R", site), 0)

    def gauge_flags(self, site: int) -> int:
        """TNPY_LEFT_IDENTITY / TNPY_RIGHT_IDENTITY bits valid for H_eff at ``site`` (0 while
        ``use_identity_channels`` is switched off: the general chain, every channel multiplied out)."""
        flags = 0
        if not self.use_identity_channels:
            return 0
        if self._left_identity.get(site, False):
            flags |= _cuda.LEFT_IDENTITY
        if self._right_identity.get(site, False):
            flags |= _cuda.RIGHT_IDENTITY
        return flags

    # -- a7 -------------------------------------------------------------------------------------------
    def update_left(self, site: int):
        prev = None if site == 1 else self._left[site - 1]
        flags = _cuda.LEFT_IDENTITY if self._left_identity.get(site - 1, False) else 0
        self._left[site] = _cuda.env_update_left(prev, self._A[site - 1], self._W[site - 1], self._left.get(site), flags)
        self._versions[("L", site)] = self._versions.get(("L", site), 0) + 1
        self._left_identity[site] = (
            self._use_identity and _cuda.identity_defect(self._left[site], 0) <= self.IDENTITY_TOL
        )

    def update_right(self, site: int):
        prev = None if site == self.n_sites - 2 else self._right[site + 1]
        flags = _cuda.RIGHT_IDENTITY if self._right_identity.get(site + 1, False) else 0
        self._right[site] = _cuda.env_update_right(prev, self._A[site + 1], self._W[site + 1], self._right.get(site), flags)
        self._versions[("R", site)] = self._versions.get(("R", site), 0) + 1
        w = self._right[site].shape[1]
        self._right_identity[site] = (
            self._use_identity and _cuda.identity_defect(self._right[site], w - 1) <= self.IDENTITY_TOL
        )

    def update(self, site: int, direction: Direction):
        if direction == Direction.RIGHTWARD:
            self.update_left(site + 1)
        elif direction == Direction.LEFTWARD:
            self.update_right(site - 1)

    def remember_image(self, site: int, image):
        """H_eff . A[site] for the tensor currently stored at ``site`` (set by the eigensolve that wrote it)."""
        self._image = (site, image)

    def take_image(self, site: int):
        """The remembered H_eff . A[site], once; None if the site tensor was rewritten since."""
        held, self._image = self._image, None
        return held[1] if held is not None and held[0] == site else None

    # -- a10 ------------------------------------------------------------------------------------------
    def update_mps(self, site: int, data):
        import torch

        self._image = None
        if isinstance(data, torch.Tensor):
            src = data.to(device="cuda", dtype=torch.float64)
        else:
            src = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).cuda()
        self._A[site].copy_(src.reshape(self._A[site].shape))
        self._dirty.add(site)

    # -- a9 -------------------------------------------------------------------------------------------
    def split_tensor(self, site: int, direction: Direction):
        if direction not in (Direction.RIGHTWARD, Direction.LEFTWARD):
            raise KeyError("MatrixProductState only supplies left or right direction.")
        nb_site = site + 1 if direction == Direction.RIGHTWARD else site - 1
        self._image = None
        a, nb, s = _split_on_device(self._A[site], self._A[nb_site], direction, self.split_mode, self.qr_min_bond)
        kind = "svd" if not isinstance(s, DeferredSpectrum) else ("qr_shifted" if s.shifted else "qr")
        self.split_counts[kind] += 1
        self._A[site], self._A[nb_site] = a.contiguous(), nb.contiguous()
        self._dirty.update((site, nb_site))
        self.bond_singular_values[min(site, nb_site)] = s
        return s

    # -- a13 ------------------------------------------------------------------------------------------
    def _expectation(self, w_tensors) -> float:
        e = None
        for site in range(self.n_sites):
            e = _cuda.env_update_left(e, self._A[site], w_tensors[site])
        return float(e.reshape(-1)[0].item())

    def variance(self) -> float:
        """<H^2> - <H>^2 on the current (un-normalised) state (:367-370), as two transfer-matrix
        passes of the environment-update chain (w^2 channels for the squared MPO)."""
        import torch

        if self._W2 is None:
            sq = self._mpo.square()
            self._W2 = [torch.from_numpy(np.ascontiguousarray(sq.as_four_leg(i))).cuda() for i in range(self.n_sites)]
        return self._expectation(self._W2) - self._expectation(self._W) ** 2

    def expectation(self) -> float:
        return self._expectation(self._W)

    # -- a6 / a1 --------------------------------------------------------------------------------------
    def one_site_full_matrix_device(self, site: int):
        L, W, R = self.operands(site)
        l, _, r = self._A[site].shape
        return _cuda.heff_dense(L, W, R, l, r)

    def one_site_full_matrix(self, site: int) -> np.ndarray:
        return self.one_site_full_matrix_device(site).cpu().numpy()

    def one_site_matvec(self, site: int, zero_copy: bool = False) -> HeffOperator:
        return HeffOperator(self, site, zero_copy=zero_copy)


class MatrixProductStateMeasurements:
    def __init__(self, mps: MatrixProductState):
        self._mps = mps

    def expectation_value(self, mpo: Optional[MatrixProductOperator] = None) -> float:
        """<psi|psi> or <psi|O|psi> (:447-453); host transfer matrices on the returned MPS."""
        if mpo is None:
            return self._mps.overlap(self._mps)
        e = np.ones((1, 1, 1))
        for site in range(self._mps.n_sites):
            a, w = self._mps.three_leg(site), mpo.as_four_leg(site)
            e = np.einsum("lam,lpr,abpq,mqs->rbs", e, a, w, a, optimize=True)
        return float(e[0, 0, 0])
