"""Host-side logic of the bond split (tnpy_b200.matrix_product_state._split_on_device) with the C-ABI calls
replaced by torch-CPU stand-ins: which entry point is tried in which order, the order of the two factors for
square site tensors, the fallback chain two-pass -> shifted -> SVD, and that the two-site product survives.
No CUDA code runs here; the kernels themselves are covered by the -m gpu tests."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tnpy_b200 import matrix_product_state as mps_mod  # noqa: E402
from tnpy_b200.matrix_product_state import DeferredSpectrum, Direction, _split_on_device  # noqa: E402


class FakeCuda:
    """torch-CPU stand-ins with the contracts of include/tnpy_cuda.h (tnpy_qr_split, tnpy_svd, tnpy_absorb_*)."""

    def __init__(self, defects=(0.0, 0.0)):
        self.defects = defects  # what the two-pass / shifted calls report
        self.calls = []

    def qr_split(self, a, shifted=False, t_first=False):
        self.calls.append(("qr", shifted, t_first))
        rows, cols = a.shape
        tall = rows > cols or (rows == cols and not t_first)
        if tall:
            q, t = torch.linalg.qr(a)
        else:
            qt, tt = torch.linalg.qr(a.t())
            q, t = qt.t().contiguous(), tt.t().contiguous()
        return q, t, self.defects[1 if shifted else 0]

    def svd(self, a):
        self.calls.append(("svd",))
        u, s, vt = torch.linalg.svd(a, full_matrices=False)
        return u, s, vt

    def absorb_right(self, s, vt, nb):
        return (s[:, None] * vt) @ nb

    def absorb_left(self, u, s, nb):
        return nb @ (u * s[None, :])


@pytest.fixture
def fake(monkeypatch):
    def install(defects=(0.0, 0.0)):
        f = FakeCuda(defects)
        monkeypatch.setattr(mps_mod, "_cuda", f)
        return f

    return install


SHAPES = [(8, 2, 12), (8, 2, 16), (6, 2, 4), (16, 2, 8), (16, 2, 4), (1, 2, 2), (2, 2, 1)]


@pytest.mark.parametrize("l,d,r", SHAPES)
@pytest.mark.parametrize("mode", ["qr", "svd"])
def test_split_preserves_two_site_product(fake, l, d, r, mode):
    f = fake()
    g = torch.Generator().manual_seed(l * 100 + r)
    a = torch.randn((l, d, r), generator=g, dtype=torch.float64)
    if l * d >= r:  # rightward: keeps the right bond
        nb = torch.randn((r, d, 5), generator=g, dtype=torch.float64)
        theta = torch.einsum("lpr,rqs->lpqs", a, nb)
        q, new_nb, s = _split_on_device(a, nb, Direction.RIGHTWARD, mode, 1)
        assert torch.allclose(torch.einsum("lpr,rqs->lpqs", q, new_nb), theta, atol=1e-12)
        iso = q.reshape(l * d, r)
        assert torch.allclose(iso.t() @ iso, torch.eye(r, dtype=torch.float64), atol=1e-12)
        if mode == "qr":
            assert f.calls[0] == ("qr", False, False)
    if d * r >= l:  # leftward: keeps the left bond
        f.calls.clear()
        nb = torch.randn((3, d, l), generator=g, dtype=torch.float64)
        theta = torch.einsum("lpr,rqs->lpqs", nb, a)
        q, new_nb, s = _split_on_device(a, nb, Direction.LEFTWARD, mode, 1)
        assert torch.allclose(torch.einsum("lpr,rqs->lpqs", new_nb, q), theta, atol=1e-12)
        iso = q.reshape(l, d * r)
        assert torch.allclose(iso @ iso.t(), torch.eye(l, dtype=torch.float64), atol=1e-12)
        if mode == "qr":
            assert f.calls[0] == ("qr", False, True)  # T first: the square case depends on it
    assert isinstance(s, DeferredSpectrum) == (mode == "qr")


def test_deferred_spectrum_is_the_bond_spectrum(fake):
    fake()
    g = torch.Generator().manual_seed(1)
    a = torch.randn((8, 2, 12), generator=g, dtype=torch.float64)
    nb = torch.randn((12, 2, 5), generator=g, dtype=torch.float64)
    _, _, s = _split_on_device(a, nb, Direction.RIGHTWARD, "qr", 1)
    want = np.linalg.svd(a.reshape(16, 12).numpy(), compute_uv=False)
    np.testing.assert_allclose(np.sort(s.cpu().numpy())[::-1], want, atol=1e-13)
    assert s.values() is s.values()  # the small SVD runs once


def test_fallback_chain_and_small_bonds(fake):
    g = torch.Generator().manual_seed(2)
    a = torch.randn((8, 2, 12), generator=g, dtype=torch.float64)
    nb = torch.randn((12, 2, 5), generator=g, dtype=torch.float64)
    f = fake(defects=(1e-9, 0.0))  # two passes rejected, shifted accepted
    _, _, s = _split_on_device(a, nb, Direction.RIGHTWARD, "qr", 1)
    assert f.calls == [("qr", False, False), ("qr", True, False)] and isinstance(s, DeferredSpectrum) and s.shifted
    f = fake(defects=(float("inf"), 1e-6))  # both rejected -> SVD
    q, new_nb, s = _split_on_device(a, nb, Direction.RIGHTWARD, "qr", 1)
    assert f.calls == [("qr", False, False), ("qr", True, False), ("svd",)] and not isinstance(s, DeferredSpectrum)
    assert torch.allclose(torch.einsum("lpr,rqs->lpqs", q, new_nb), torch.einsum("lpr,rqs->lpqs", a, nb), atol=1e-12)
    f = fake()  # bond below qr_min_bond: straight to the SVD
    _split_on_device(a, nb, Direction.RIGHTWARD, "qr", 64)
    assert f.calls == [("svd",)]


def test_split_rejects_impossible_bonds(fake):
    fake()
    a = torch.zeros((2, 2, 8), dtype=torch.float64)
    with pytest.raises(ValueError):
        _split_on_device(a, torch.zeros((8, 2, 3), dtype=torch.float64), Direction.RIGHTWARD, "qr", 1)
