"""One verified Cholesky-QR split at the chi=2048 bench shape, for `ncu` launch lists (run under gpurun)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402

rows, cols = 4096, 2048
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((rows, cols), generator=g, dtype=torch.float64, device="cuda")
a = (a * torch.logspace(0, -12, cols, dtype=torch.float64, device="cuda")[None, :]).contiguous()
q, t, defect = _cuda.qr_split(a)
torch.cuda.synchronize()
print("defect", defect)
