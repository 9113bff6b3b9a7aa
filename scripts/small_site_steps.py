"""Time the fused small-site Lanczos steps (csrc/lanczos_steps.cu) at one site shape: seconds per step from two solves
that differ only in the number of steps, the cost of a look, and the same through the general multi-kernel solver.

    python scripts/small_site_steps.py [--chi 60] [--w 5]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chi", type=int, default=60)
    ap.add_argument("--w", type=int, default=5)
    ap.add_argument("--reps", type=int, default=50)
    args = ap.parse_args()

    import torch

    from tnpy_b200 import _cuda as cu

    chi, w, d = args.chi, args.w, 2
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *shape: torch.randn(shape, generator=g, dtype=torch.float64, device="cuda")
    L = rnd(chi, w, chi)
    L = L + L.permute(2, 1, 0)
    R = rnd(chi, w, chi)
    R = R + R.permute(2, 1, 0)
    W = rnd(w, w, d, d)
    W = W + W.permute(0, 1, 3, 2)
    W = W * (torch.rand(w, w, 1, 1, generator=g, dtype=torch.float64, device="cuda") < 0.3)
    start = rnd(chi, d, chi)

    def solve_seconds(max_matvec, fused):
        previous = cu.set_fused_steps(fused)
        try:
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    psi = start.clone()
                    stats = cu.eig_lowest(L, W, R, psi, tol=1e-30, ncv=32, max_matvec=max_matvec)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / args.reps
                best = dt if best is None else min(best, dt)
            return best, stats
        finally:
            cu.set_fused_steps(previous)

    out = {"shape": [chi, d, chi], "w": w}
    for fused in (True, False):
        t10, s10 = solve_seconds(10, fused)
        t30, s30 = solve_seconds(30, fused)
        key = "fused" if fused else "general"
        out[key] = {"solve_10_steps_us": 1e6 * t10, "solve_30_steps_us": 1e6 * t30, "us_per_step": 1e6 * (t30 - t10) / 20,
                    "looks_10": s10["looks"], "looks_30": s30["looks"], "n_matvec_30": s30["n_matvec"]}
    if not cu.load().tnpy_set_fused_steps(1) or chi * d * chi > 32768:
        print(json.dumps(out))  # beyond the fused path's range: the general solver's numbers only
        return
    # phase durations inside one 30-step launch (ns, CTA 0): P1, barrier 1, P2, P3, P4, P5 (each up to the point where
    # the CTA arrives at the following barrier, i.e. including the wait at the preceding one), barrier 5, P6
    trace = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
    cu.load().tnpy_steps_trace(trace.data_ptr())
    psi = start.clone()
    cu.eig_lowest(L, W, R, psi, tol=1e-30, ncv=32, max_matvec=30)
    torch.cuda.synchronize()
    cu.load().tnpy_steps_trace(None)
    t = trace.cpu().numpy().reshape(64, 16)
    names = ["P1", "barrier1", "P2", "barrier2+P3", "barrier3+P4", "barrier4+P5", "barrier5"]
    rows = {}
    for step in (1, 5, 15, 28):
        cta0 = t[step, :8]
        watcher = t[step, 8:]
        rows[str(step)] = {"cta0_ns": {n: int(cta0[i + 1] - cta0[i]) for i, n in enumerate(names)},
                           "step_ns": int(t[step + 1, 0] - t[step, 0]),
                           "watcher_ns": {n: int(watcher[i + 1] - watcher[i]) for i, n in enumerate(names)}}
    out["trace"] = rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()
