"""Profiling target (run under ncu via gpurun): the chi=2048 matvec on every path plus a short on-device eigensolve.

    python scripts/ncu_target.py [direct|chain|fp64|eig|vectors ...]   (default: all)

Operands are synthetic (random L / R with the identity channels of the mixed-canonical gauge planted, the XXZ MPO
tensor); every section is preceded by an NVTX-free marker launch-free pause so that launch lists read in order."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnpy_b200 import _cuda  # noqa: E402
from tnpy_b200.model import XXZ  # noqa: E402


def main():
    what = set(sys.argv[1:]) or {"direct", "chain", "fp64", "eig", "vectors"}
    chi, w, d = int(os.environ.get("CHI", "2048")), 5, 2
    _cuda.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(s, generator=g, dtype=torch.float64, device="cuda")  # noqa: E731
    L, R = rnd(chi, w, chi) / chi**0.5, rnd(chi, w, chi) / chi**0.5
    L = 0.5 * (L + L.permute(2, 1, 0))
    R = 0.5 * (R + R.permute(2, 1, 0))
    L[:, 0, :] = torch.eye(chi, dtype=torch.float64, device="cuda")
    R[:, w - 1, :] = torch.eye(chi, dtype=torch.float64, device="cuda")
    L, R = L.contiguous(), R.contiguous()
    Wh = np.ascontiguousarray(XXZ(n=4, delta=0.5).mpo.as_four_leg(1))
    W = torch.from_numpy(Wh).cuda()
    x = rnd(chi, d, chi)
    x /= x.norm()
    y = torch.empty_like(x)
    reps = int(os.environ.get("REPS", "3"))
    if "direct" in what:
        plan = _cuda.HeffPlan(L, W, R, chi, chi, flags=3, w_host=Wh)
        assert plan.mode == _cuda.HEFF_OZ_DIRECT
        for s in (8, 7):
            for _ in range(reps):
                plan.apply(x, y, slices=s)
        torch.cuda.synchronize()
        plan.close()
    if "chain" in what:
        plan = _cuda.HeffPlan(L, W, R, chi, chi, flags=0, w_host=Wh)
        for _ in range(reps):
            plan.apply(x, y)
        torch.cuda.synchronize()
        plan.close()
    if "fp64" in what:
        plan = _cuda.HeffPlan(L, W, R, chi, chi, flags=3, algo=_cuda.GEMM_FP64)
        for _ in range(reps):
            plan.apply(x, y)
        torch.cuda.synchronize()
        plan.close()
    if "eig" in what:
        psi = x.clone()
        stats = _cuda.eig_lowest(L, W, R, psi, tol=1e-8, max_matvec=int(os.environ.get("MATVECS", "40")), flags=3)
        print("eig", stats, flush=True)
    if "vectors" in what:
        n = chi * chi * d
        V = rnd(33, n)
        wv = rnd(n)
        for m in (8, 16, 32):
            h = _cuda.multi_dot(V, wv, m)
            _cuda.multi_axpy(V, h * 1e-3, wv, m)
        _cuda.nrm2(wv); _cuda.dot(wv, V[0]); _cuda.axpy(0.5, V[1], wv); _cuda.scal(0.99, wv)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
