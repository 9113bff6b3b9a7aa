"""Tiny ncu target: one tcgen05 Ozaki GEMM launch per chi=2048 matvec shape (variant from OZ_VARIANTS)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

_cuda.load()
_cuda.set_ozaki_variant(int(os.environ.get("OZ_VARIANTS", "2").split(",")[0]))
g = torch.Generator(device="cuda").manual_seed(0)
for (m, n, k) in ((4096, 10240, 2048), (4096, 2048, 10240)):
    a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    c = torch.empty((m, n), dtype=torch.float64, device="cuda")
    _cuda.ozaki_gemm_tn(a, b, out=c, slices=8, phase=1)
    for _ in range(2):
        _cuda.ozaki_gemm_tn(a, b, out=c, slices=8, phase=2)
    torch.cuda.synchronize()
