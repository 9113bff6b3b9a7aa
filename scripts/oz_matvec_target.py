"""ncu / timing target: chi=2048 H_eff matvec on the tcgen05 path, L and R declared constant (as in tnpy_eig_lowest)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

_cuda.load()
_cuda.set_gemm_algo(_cuda.GEMM_OZAKI)
chi, w, d = int(os.environ.get("CHI", "2048")), 5, 2
g = torch.Generator(device="cuda").manual_seed(0)
L = torch.randn((chi, w, chi), generator=g, dtype=torch.float64, device="cuda")
R = torch.randn((chi, w, chi), generator=g, dtype=torch.float64, device="cuda")
W = torch.randn((w, w, d, d), generator=g, dtype=torch.float64, device="cuda")
x = torch.randn((chi, d, chi), generator=g, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
flags = int(os.environ.get("FLAGS", "0"))
if flags:
    L[:, 0, :] = torch.eye(chi, dtype=torch.float64, device="cuda")
    R[:, w - 1, :] = torch.eye(chi, dtype=torch.float64, device="cuda")
_cuda.ozaki_const_scope(True)
for _ in range(int(os.environ.get("REPS", "3"))):
    _cuda.heff_apply(L, W, R, x, y, flags=flags)
torch.cuda.synchronize()
_cuda.ozaki_const_scope(False)
