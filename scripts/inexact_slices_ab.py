import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from tnpy_b200 import _cuda as cu
# five-slice GEMM against FP64
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((2048, 1024), generator=g, dtype=torch.float64, device="cuda")
b = torch.randn((2048, 512), generator=g, dtype=torch.float64, device="cuda")
ref = a.t() @ b
errs = {}
for s in (5, 6, 7, 8):
    c = cu.ozaki_gemm_tn(a, b, slices=s)
    errs[s] = float((c - ref).abs().max() / ref.abs().max())
out = {"gemm_rel_err_by_slices": errs}
out["sweeps"] = bench.measure_sweeps(40, 2048, 1e-8, 3)
print(json.dumps(out))
