import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda
_cuda.load(); _cuda.set_ozaki_variant(2)
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
g = torch.Generator(device="cuda").manual_seed(0)
for (m, n, k) in ((4096, 10240, 2048), (4096, 2048, 10240), (4096, 10240, 8192)):
    a = torch.randn((k, m), generator=g, dtype=torch.float64, device="cuda")
    b = torch.randn((k, n), generator=g, dtype=torch.float64, device="cuda")
    c = torch.empty((m, n), dtype=torch.float64, device="cuda")
    _cuda.ozaki_gemm_tn(a, b, out=c, slices=8, phase=1)
    row = {"shape": [m, n, k], "ideal_ms_at_4.48POPs": 2.0*m*n*k*36/4.48e15*1e3}
    for flags in (0,):
        row[f"ms_flags{flags}"] = timed(lambda: _cuda.ozaki_gemm_tn(a, b, out=c, slices=8, phase=2, accumulate=flags))
    print(json.dumps(row), flush=True)
