import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
from tnpy_b200 import _cuda as cu
g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 4096, 2048, 8192
a = torch.randn((K, M), generator=g, dtype=torch.float64, device="cuda")
b = torch.randn((K, N), generator=g, dtype=torch.float64, device="cuda")
out = torch.empty((M, N), dtype=torch.float64, device="cuda")
res = {}
for s in (8, 7, 6, 5):
    cu.ozaki_gemm_tn(a, b, out=out, slices=s, phase=1)
    for _ in range(3): cu.ozaki_gemm_tn(a, b, out=out, slices=s, phase=2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): cu.ozaki_gemm_tn(a, b, out=out, slices=s, phase=2)
    e1.record(); torch.cuda.synchronize()
    res[s] = e0.elapsed_time(e1) / 10
print(json.dumps({"mma_only_ms_by_slices_4096x2048x8192": res}))
