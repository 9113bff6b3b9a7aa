#!/usr/bin/env python
"""bench.py -- the fDMRG local-update hot path on B200: H_eff matvec FP64 TFLOP/s (+ measured sweep seconds).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one H_eff.psi contraction at a full-chi bulk site
(the call primme makes ~10-30x per site in the reference, matrix_product_state.py:411-440).

  N = 1   workload = BASELINE.json configs[2]: XXZ n=100 delta=0.5 chi=2048 (w=5, d=2), mid-chain site of a
          synthetic random MPS brought into the mixed-canonical form it has inside a sweep (left half
          left-canonical by the sweep's own QR splits, right half right-canonical), genuine L / R from the
          update recursions.
  N > 1   workload = configs[4]: XXZ n=200 chi=8192, the matvec sharded over chi-row blocks of L
          (SURVEY 8e.1): every rank holds L[:, :, rows_g], the full R and x; per step the ranks
          all-gather x over NCCL and run the local chain; no reduction is needed.

`value` is whole-job algorithmic TFLOP/s (F_mv = 4 w d chi^3 + 2 w^2 d^2 chi^2 per matvec) of the
GENERAL chain -- every MPO channel of both environments multiplied out, executed flops = algorithmic flops -- on
the library's default path (large GEMMs FP64-accurately on the tcgen05 int8 tensor cores), with operands resident
in HBM, through the prepared operator the eigensolver uses.  `e2e` is the same through the reference-facing seam
(``Environment.one_site_matvec(site).matvec(x_host)``): pinned host vector in, host vector out, both copies inside
the timed region.  `canonical_gauge` is the matvec as every sweep step actually runs it (identity channels of the
mixed-canonical gauge measured and skipped: the direct path, executed flops 0.8 F_mv), `native_fp64` the same two
on the FP64 tensor pipe (DMMA).  `roofline` is for the dominant kernel of `value`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm (--impl reference, rank 0 only) has to
# use all host cores, and OpenBLAS sizes its pool when NumPy is first imported.
if "reference" in sys.argv and os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def f_mv(l, r, wl, wr, d):
    """Algorithmic flops of one matvec."""
    return 2.0 * wl * d * l * l * r + 2.0 * wl * wr * d * d * l * r + 2.0 * wr * d * l * r * r


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=200):
        self.index = index
        self.period_ms = period_ms
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms), "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, flag in zip(names, parts[4:8]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        return out


def cuda_time(fn, steps, stream_sync):
    """Time `steps` calls of fn with CUDA events on the current stream; returns seconds."""
    import torch

    stream_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    stream_sync()
    return e0.elapsed_time(e1) * 1e-3


def random_right_canonical_device(n, chi, d, seed):
    """Synthetic initial MPS born on the device (input generation, not the hot path): Gaussian
    tensors with orthonormal rows (right-canonical), bond dims min(d^i, chi, d^(n-i))."""
    import torch

    from tnpy_b200.matrix_product_state import compressed_bond_dims

    g = torch.Generator(device="cuda").manual_seed(seed)
    dims = [1] + compressed_bond_dims(n, chi, d) + [1]
    out = []
    for i in range(n):
        l, r = dims[i], dims[i + 1]
        a = torch.randn((d * r, l), generator=g, dtype=torch.float64, device="cuda")
        q, _ = torch.linalg.qr(a)  # (d r, l) orthonormal columns
        out.append(q.t().contiguous().reshape(l, d, r))
    out[0] = out[0] / out[0].norm()
    return out


# ---------------------------------------------------------------------------------------------------
def all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core it can."""
    try:
        from threadpoolctl import threadpool_limits

        return threadpool_limits(limits=os.cpu_count())
    except Exception:
        from contextlib import nullcontext

        return nullcontext()


def cpu_workload(chi, n_gpus):
    """Operands of the CPU arm: the workload's own chi.  chi = 2048: the full matvec (~1.2 s of host BLAS).
    chi = 8192 (the sharded workload): a 128-row block of it -- y[rows] = H_eff(L[:, :, rows], W, R) x, the unit of
    work one rank of the GPU arm does -- because one full matvec is 2.2e13 flop, minutes per step on host cores."""
    from oracle import tnpy_oracle as oracle

    w, d = 5, 2
    rows = chi if chi <= 2048 else 128
    rng = np.random.default_rng(0)
    L = rng.standard_normal((chi, w, rows))
    R = rng.standard_normal((chi, w, chi))
    W = np.ascontiguousarray(oracle.xxz_mpo(4, 0.5)[1])
    x = rng.standard_normal((chi, d, chi))
    flops = f_mv(chi, chi, w, w, d) * rows / chi
    what = (f"NumPy tensordot-chain matvecs (oracle/tnpy_oracle.heff_apply) at chi={chi}, w=5, d=2"
            + ("" if rows == chi else f", rows [0, {rows}) of the output (1/{chi // rows} of one matvec per step)"))
    return (lambda: oracle.heff_apply(L, W, R, x)), flops, what


def reference_install():
    """The real reference (tnpy + quimb + primme) when it is importable on this box -- an install under
    baseline/_ref or site-packages.  It is not in the build container (no wheels for quimb / primme /
    tensornetwork); the probe is here so that a box that has it uses it."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import primme  # noqa: F401
        import quimb  # noqa: F401
        import tnpy as ref_tnpy

        if os.path.realpath(os.path.dirname(ref_tnpy.__file__)).startswith(os.path.realpath(ROOT) + os.sep + "tnpy"):
            return None  # that is this repo's alias package, not the reference
        return ref_tnpy
    except Exception:
        return None


def reference_matvec_step(ref_tnpy, chi):
    """One H_eff matvec of the real reference at a full-chi site of XXZ n=24 (same bulk shapes as n=100)."""
    from tnpy.matrix_product_state import Environment, MatrixProductState  # the reference's
    from tnpy.model import XXZ

    n = max(24, 2 * int(np.ceil(np.log2(chi))) + 2)
    env = Environment(XXZ(n=n, delta=0.5).mpo, MatrixProductState.random(n=n, bond_dim=chi, phys_dim=2))
    op = env.one_site_matvec(n // 2)
    x = np.random.default_rng(0).standard_normal(op.shape[0])
    return (lambda: op.matvec(x)), f_mv(chi, chi, 5, 5, 2), f"tnpy/quimb Environment.one_site_matvec(site).matvec at chi={chi}"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on the box's host cores -- the real
    tnpy/quimb/primme when importable, else its NumPy restatement in oracle/ -- at the workload's own chi."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with all_host_threads():
        return _run_reference(args)


def _run_reference(args):
    chi = args.chi or (2048 if args.gpus == 1 else 8192)
    ref_tnpy = reference_install() if chi <= 2048 else None
    kind = "reference" if ref_tnpy is not None else "port"
    step, flops, what = reference_matvec_step(ref_tnpy, chi) if ref_tnpy is not None else cpu_workload(chi, args.gpus)
    for _ in range(max(args.warmup, 1) if chi <= 2048 else 1):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = flops * args.steps / dt / 1e12
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "heff_matvec_fp64_tflops", "value": value, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, chi), "chi": chi, "cpu_sample_chi": chi, "same_config": True},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": f"{args.steps} {what}"},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(n_gpus, chi):
    if n_gpus == 1:
        return f"XXZ n=100 delta=0.5 chi={chi}: H_eff matvec at the mid-chain site (BASELINE configs[2])"
    return f"XXZ n=200 delta=0.5 chi={chi}: H_eff matvec sharded over chi-row blocks of L on {n_gpus} GPUs (BASELINE configs[4])"


def cpu_baseline(chi, budget_s=20.0):
    """Oracle matvec on the host cores at the workload's chi, bounded sample."""
    with all_host_threads():
        step, flops, what = cpu_workload(chi, 1)
        step()
        t0 = time.perf_counter()
        n = 0
        while True:
            step()
            n += 1
            if time.perf_counter() - t0 > budget_s * 0.6 or n >= 100:
                break
        dt = (time.perf_counter() - t0) / n
    return {"value": flops / dt / 1e12, "unit": "TFLOP/s", "cores": os.cpu_count(), "kind": "port",
            "ms_per_matvec": dt * 1e3, "sample": f"{n} {what}"}


# ---------------------------------------------------------------------------------------------------
INT8_PEAK_POPS = 4.48  # ncu sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained at 1965 MHz (profiles/r01_oz2_mma_kernel_ncu_full_raw.csv)


def mixed_canonicalize(dmrg, site):
    """Left-canonicalise sites 0 .. site-1 with the sweep's own split + environment update (no local solves):
    the state at `site` is then in the mixed-canonical form every sweep step sees, and the identity channels of
    both environments are *measured* by tnpy_identity_defect, not planted."""
    from tnpy_b200.matrix_product_state import Direction

    env = dmrg.environment
    for s in range(site):
        env.split_tensor(s, Direction.RIGHTWARD)
        env.update(s, Direction.RIGHTWARD)


def timed_plan(plan, x, y, steps, warmup, sync, slices=0):
    for _ in range(warmup):
        plan.apply(x, y, slices=slices)
    return cuda_time(lambda: plan.apply(x, y, slices=slices), steps, sync) / steps


def fp64_peak_cublas(m8, sync):
    """FP64 GEMM yardstick measured in this process: cuBLAS dgemm m8^3, best of 5 (burst)."""
    import torch

    a = torch.randn(m8, m8, dtype=torch.float64, device="cuda")
    b = torch.randn(m8, m8, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = min(cuda_time(lambda: torch.matmul(a, b), 1, sync) for _ in range(5))
    return 2.0 * m8**3 / best / 1e12


def int8_peak_cublaslt(sync):
    """int8 tensor-core yardstick measured in this process the way MEASURED_PEAKS.json measures bf16: a library GEMM
    (cuBLASLt through torch._int_mm, int8 x int8 -> int32, 8192^3), best of 5 (burst) and back to back for 1.5 s
    (sustained, under the power cap).  None when the library call is not available."""
    import torch

    try:
        n = 8192
        a = torch.randint(-64, 65, (n, n), dtype=torch.int8, device="cuda")
        b = torch.randint(-64, 65, (n, n), dtype=torch.int8, device="cuda")
        torch._int_mm(a, b)
        sync()
        best = min(cuda_time(lambda: torch._int_mm(a, b), 1, sync) for _ in range(5))
        t0, cnt = time.perf_counter(), 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < 1.5:
            for _ in range(20):
                torch._int_mm(a, b)
            cnt += 20
            sync()
        e1.record()
        sync()
        return {"burst_pops": 2.0 * n**3 / best / 1e15, "sustained_pops": 2.0 * n**3 * cnt / (e0.elapsed_time(e1) * 1e-3) / 1e15,
                "how": "torch._int_mm 8192^3 (cuBLASLt int8 -> int32): best of 5, and back to back for 1.5 s"}
    except Exception as exc:  # pragma: no cover - depends on the installed cuBLASLt
        return {"burst_pops": None, "sustained_pops": None, "how": f"torch._int_mm unavailable: {exc}"}


def run_single(args):
    import torch

    from tnpy_b200 import _cuda
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.model import XXZ

    torch.cuda.set_device(0)
    lib = _cuda.load()
    chi = args.chi or 2048
    n = args.n or 100
    d = 2
    mpo = XXZ(n=n, delta=0.5).mpo
    sync = torch.cuda.synchronize
    m8 = 8192 if chi >= 1024 else 4096
    fp64_peak = fp64_peak_cublas(m8, sync)
    int8_lib = int8_peak_cublaslt(sync)

    t_setup = time.perf_counter()
    tensors = random_right_canonical_device(n, chi, d, seed=0)
    dmrg = FiniteDMRG(mpo, bond_dim=chi, mps=tensors, compute_variance=False)
    env = dmrg.environment
    del tensors
    site = n // 2
    mixed_canonicalize(dmrg, site)
    sync()
    t_setup = time.perf_counter() - t_setup
    L, W, R = env.operands(site)
    W_host = mpo.as_four_leg(site)
    x = env.device_tensor(site).clone()
    y = torch.empty_like(x)
    l, _, r = x.shape
    wl, wr = W.shape[0], W.shape[1]
    flops = f_mv(l, r, wl, wr, d)
    gauge = env.gauge_flags(site)
    tf = lambda sec: flops / sec / 1e12  # noqa: E731

    # ---- the step: general chain on the default path, through the prepared operator ----------------
    plan_g = _cuda.HeffPlan(L, W, R, l, r, flags=0, w_host=W_host)
    for _ in range(args.warmup):
        plan_g.apply(x, y)
    sync()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = _cuda.launch_count()
    dt = cuda_time(lambda: plan_g.apply(x, y), args.steps, sync)
    launches = _cuda.launch_count() - launches0
    y_general = y.clone()
    bound_g = plan_g.error_bound()
    mode_g = plan_g.mode

    # ---- the same operands on the native FP64 chain (DMMA) ---------------------------------------
    plan_f = _cuda.HeffPlan(L, W, R, l, r, flags=0, algo=_cuda.GEMM_FP64)
    t_f = timed_plan(plan_f, x, y, args.steps, args.warmup, sync)
    y_fp64 = y.clone()
    diff_general = float((y_general - y_fp64).abs().max() / y_fp64.abs().max())
    plan_f.close()

    # ---- as a sweep runs it: measured identity channels (the direct path when both hold) ---------
    plan_c = _cuda.HeffPlan(L, W, R, l, r, flags=gauge, w_host=W_host)
    t_c8 = timed_plan(plan_c, x, y, args.steps, args.warmup, sync, slices=8)
    y_c8 = y.clone()
    t_c7 = timed_plan(plan_c, x, y, args.steps, 1, sync, slices=7)
    y_c7 = y.clone()
    bound_c7 = plan_c.error_bound()
    # the slice counts of the eigensolver's inexact-Krylov schedule (later steps of a solve): time, error, bound
    by_slices = {}
    for sl in (6, 5):
        t_sl = timed_plan(plan_c, x, y, args.steps, 1, sync, slices=sl)
        by_slices[str(sl)] = {"ms_per_step": t_sl * 1e3, "rel_diff_vs_8_slices": float((y - y_c8).abs().max() / y_c8.abs().max()),
                              "int8_error_bound_frobenius": plan_c.error_bound()}
    bound_c = bound_c7  # the plan keeps the largest bound of its products: read before the reduced counts ran
    mode_c = plan_c.mode
    plan_cf = _cuda.HeffPlan(L, W, R, l, r, flags=gauge, algo=_cuda.GEMM_FP64)
    t_cf = timed_plan(plan_cf, x, y, args.steps, args.warmup, sync)
    y_cf = y.clone()
    plan_cf.close()
    scale = float(y_fp64.abs().max())
    canonical = {
        "ms_per_step": t_c8 * 1e3, "tflops_algorithmic": tf(t_c8), "heff_mode": mode_c, "gauge_flags_measured": gauge,
        "slices": 8, "executed_flop_fraction": ((wl - 1) / wl if wl == wr else None) if gauge == 3 else 1.0,
        "rel_diff_vs_general_fp64_chain": float((y_c8 - y_fp64).abs().max()) / scale,
        "rel_diff_vs_canonical_fp64_chain": float((y_c8 - y_cf).abs().max()) / scale,
        "ms_per_step_7_slices": t_c7 * 1e3, "tflops_algorithmic_7_slices": tf(t_c7),
        "rel_diff_7_vs_8_slices": float((y_c7 - y_c8).abs().max()) / scale,
        "reduced_slices": by_slices,
        "int8_error_bound_frobenius": bound_c, "y_norm": float(y_fp64.norm()),
        "native_fp64_ms_per_step": t_cf * 1e3, "native_fp64_tflops_algorithmic": tf(t_cf),
        "note": "the matvec of every sweep step: L[:,0,:] = R[:,w-1,:] = I measured on this state (mixed-canonical "
                "at the site) and skipped; heff_mode 2 = direct path (two independent tcgen05 GEMMs on operands "
                "premixed and sliced from x, no FP64 intermediate); 7 slices is what tnpy_eig_lowest starts with at tol >= 1e-10, 6 and 5 "
                "what its inexact-Krylov schedule moves to as the residual falls (accepted on the true residual)",
    }
    del y_c8, y_c7, y_cf

    # ---- dominant kernels alone -------------------------------------------------------------------
    # (a) oz2_mma_kernel on the shapes it runs: operands already sliced (phase 2), int8 ops = 36 slice pairs x 2MNK
    def oz_kernel_ms(m_, n_, k_):
        a_ = torch.randn((k_, m_), dtype=torch.float64, device="cuda")
        b_ = torch.randn((k_, n_), dtype=torch.float64, device="cuda")
        c_ = torch.empty((m_, n_), dtype=torch.float64, device="cuda")
        _cuda.ozaki_gemm_tn(a_, b_, out=c_, slices=8, phase=1)
        mm = lambda: _cuda.ozaki_gemm_tn(a_, b_, out=c_, slices=8, phase=2)  # noqa: E731
        mm()
        return cuda_time(mm, args.steps, sync) / args.steps * 1e3

    shapes_general = [(d * r, wl * l, l), (d * l, r, r * wr)]
    shapes_direct = [(l * d, r, (wr - 1) * r), (l, d * r, (wl - 1) * l)]
    oz_g = [oz_kernel_ms(*sh) for sh in shapes_general]
    oz_d = [oz_kernel_ms(*sh) for sh in shapes_direct]
    pops = lambda shapes, ms: sum(36 * 2.0 * m_ * n_ * k_ for m_, n_, k_ in shapes) / (sum(ms) * 1e-3) / 1e15  # noqa: E731

    # the same kernel back to back for ~1.5 s with the SM clock sampled: under this load the board sits at its power
    # cap and the clock well below 1965 MHz, so the fraction of the MMA peak *at the clock the kernel is given* is
    # the number that says how well the kernel uses the pipe
    def oz_sustained(m_, n_, k_, seconds=1.5):
        a_ = torch.randn((k_, m_), dtype=torch.float64, device="cuda")
        b_ = torch.randn((k_, n_), dtype=torch.float64, device="cuda")
        c_ = torch.empty((m_, n_), dtype=torch.float64, device="cuda")
        _cuda.ozaki_gemm_tn(a_, b_, out=c_, slices=8, phase=1)
        mm = lambda: _cuda.ozaki_gemm_tn(a_, b_, out=c_, slices=8, phase=2)  # noqa: E731
        for _ in range(50):
            mm()
        sync()
        smp = ClockSampler(0, period_ms=100)
        smp.start()
        time.sleep(0.15)
        t0, launches_ = time.perf_counter(), 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.perf_counter() - t0 < seconds:
            for _ in range(50):
                mm()
            launches_ += 50
            sync()
        e1.record()
        sync()
        clk = smp.stop()
        ms_ = e0.elapsed_time(e1) / launches_
        rate = 36 * 2.0 * m_ * n_ * k_ / (ms_ * 1e-3) / 1e15
        mhz = clk.get("sm_mhz") or 1965.0
        return {"shape_mnk": [m_, n_, k_], "ms_per_launch": ms_, "launches": launches_, "achieved": rate, "sm_mhz_median": mhz,
                "power_w_max": clk.get("power_w_max"), "reasons": clk.get("reasons"),
                "peak_at_that_clock": INT8_PEAK_POPS * mhz / 1965.0, "frac_at_that_clock": rate / (INT8_PEAK_POPS * mhz / 1965.0),
                "frac_of_1965mhz_peak": rate / INT8_PEAK_POPS}

    oz_sus = oz_sustained(*shapes_direct[0])
    # (b) gemm_tn_dmma on the two GEMM shapes of the FP64 chain
    t1 = torch.empty((d * r, wl * l), dtype=torch.float64, device="cuda")
    t2 = torch.randn((r * wr, d * l), dtype=torch.float64, device="cuda")
    xm, Lm, Rm = x.reshape(l, d * r), L.reshape(l, wl * l), R.reshape(r * wr, r)
    yq = torch.empty((d * l, r), dtype=torch.float64, device="cuda")
    g1 = lambda: _cuda.gemm_tn(xm, Lm, out=t1, algo=_cuda.GEMM_DMMA)  # noqa: E731
    g3 = lambda: _cuda.gemm_tn(t2, Rm, out=yq, algo=_cuda.GEMM_DMMA)  # noqa: E731
    g1(); g3()
    tg1 = cuda_time(g1, args.steps, sync) / args.steps
    tg3 = cuda_time(g3, args.steps, sync) / args.steps
    gemm_flops = 2.0 * (d * r) * (wl * l) * l + 2.0 * (d * l) * r * (r * wr)
    del t1, t2, yq

    # ---- e2e through the reference-facing seam: host (pinned) vector in, host vector out ------------
    def e2e_seconds(use_identity):
        env.use_identity_channels = use_identity
        op = env.one_site_matvec(site, zero_copy=True)
        x_host = torch.empty(x.numel(), dtype=torch.float64).pin_memory()
        x_host.copy_(x.reshape(-1))
        x_np = x_host.numpy()
        for _ in range(max(2, args.warmup // 2)):
            op.matvec(x_np)
        sync()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y_np = op.matvec(x_np)
        e1.record()
        sync()
        sec = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0) / args.steps
        return sec, np.array(y_np, copy=True)

    # one on-device local solve, cut off after 60 matvecs: what a Lanczos step costs on top of its matvec
    env.use_identity_channels = True
    psi_try = x.clone()
    img = torch.empty_like(psi_try)
    _cuda.eig_lowest(L, W, R, psi_try.clone(), tol=args.tol, max_matvec=3, flags=gauge)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = _cuda.eig_lowest(L, W, R, psi_try, tol=args.tol, max_matvec=60, flags=gauge, image=img)
    e1.record()
    sync()
    lanczos = {"matvecs": st["n_matvec"], "ms_per_step": e0.elapsed_time(e1) / max(st["n_matvec"], 1), "slices": st["slices"],
               "heff_mode": st["heff_mode"], "restarts": st["n_restart"], "int8_error_bound": st["int8_error_bound"],
               "note": "tnpy_eig_lowest at the bench site (measured gauge flags), stopped after 60 matvecs: whole-step time = matvec + "
                       "Gram-Schmidt passes + Ritz solves + status read-backs + environment slicing once"}
    del psi_try, img

    t_e2e, y_np = e2e_seconds(False)
    parity = float(np.abs(y_np - y_general.reshape(-1).cpu().numpy()).max())
    t_e2e_c, _ = e2e_seconds(True)
    clocks = sampler.stop()
    canonical["e2e_ms_per_step"] = t_e2e_c * 1e3
    canonical["e2e_tflops_algorithmic"] = tf(t_e2e_c)
    plan_g.close(); plan_c.close()

    value = flops * args.steps / dt / 1e12
    roofline_oz = {
        "bound": "tensor", "achieved": pops(shapes_general, oz_g), "peak": INT8_PEAK_POPS, "unit": "POP/s (int8)",
        "frac": pops(shapes_general, oz_g) / INT8_PEAK_POPS,
        "library_int8_gemm": int8_lib,
        "frac_of_library_int8_gemm_burst": (pops(shapes_general, oz_g) / int8_lib["burst_pops"]) if int8_lib.get("burst_pops") else None,
        "traffic": 1.63e9,  # dram read + write per launch, mean of the two shapes, ncu --set full (profiles/r01_oz2_mma_kernel_ncu_full_raw.csv)
        "algorithmic_bytes_per_launch": 8.0 * (shapes_general[0][0] + shapes_general[0][1]) * shapes_general[0][2] + 8.0 * shapes_general[0][0] * shapes_general[0][1],
        "kernel": "oz2_mma_kernel<8> (tcgen05.mma.cta_group::2.kind::i8; 2 launches per matvec, operands already sliced)",
        "ms_gemm1": oz_g[0], "ms_gemm3": oz_g[1], "shapes_mnk": shapes_general,
        "fp64_equivalent_tflops": gemm_flops / (sum(oz_g) * 1e-3) / 1e12,
        "algorithmic_ops": "36 slice pairs x 2 M N K int8 multiply-adds per launch (8 slices; DESIGN 2.2)",
        "peak_source": "ncu sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained at 1965 MHz (profiles/r01_oz2_mma_kernel_ncu_full_raw.csv)",
        "direct_path_shapes": {"shapes_mnk": shapes_direct, "ms": oz_d, "achieved": pops(shapes_direct, oz_d),
                               "frac": pops(shapes_direct, oz_d) / INT8_PEAK_POPS,
                               "traffic": 1.64e9,  # ncu --set full: 1.51 GB read + 0.13 GB written per launch (profiles/r02_direct_path_ncu_full_raw.csv)
                               "algorithmic_bytes_per_launch": 8.0 * (shapes_direct[0][0] + shapes_direct[0][1]) * shapes_direct[0][2] + 16.0 * shapes_direct[0][0] * shapes_direct[0][1]},
        "sustained_at_measured_clock": oz_sus,
    }
    roofline_dmma = {
        "bound": "tensor", "achieved": gemm_flops / (tg1 + tg3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": gemm_flops / (tg1 + tg3) / 1e12 / fp64_peak, "traffic": 1.29e9 if chi == 2048 else None,
        "kernel": "gemm_tn_dmma (2 launches per matvec)", "ms_gemm1": tg1 * 1e3, "ms_gemm3": tg3 * 1e3,
        "peak_source": f"cuBLAS dgemm {m8}^3 via torch.matmul, best of 5 in this process; hardware DMMA peak and the "
                       "sustained figure are tracked in profiles/r02_fp64_peak.json (MEASURED_PEAKS.json has no FP64 entry)",
    }
    line = {
        "metric": "heff_matvec_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 (results); large GEMMs as exact int8 slice products with int32 accumulation, FP64 recombination",
        "data": "synthetic",
        "config": {
            "workload": workload_name(1, chi), "n": n, "chi": chi, "w": wl, "d": d, "site": site,
            "operands": "general chain: all w channels of L and R multiplied out (flags 0), executed = algorithmic flops",
            "path": {0: "fp64 chain", 1: "chain, GEMMs on tcgen05 int8 (Ozaki, 8 slices)", 2: "direct"}[mode_g],
            "flops_per_step": flops,
            "l2_policy": "operands (L 168 MB + R 168 MB + int8 slices 0.3-0.6 GB + T1/T2 670 MB at chi=2048) exceed the 126 MB L2; no flush needed",
            "setup_s": t_setup,
        },
        "rel_diff_vs_native_fp64_chain": diff_general,
        "int8_error_bound_frobenius": bound_g,
        "roofline": roofline_oz if mode_g != 0 else roofline_dmma,
        "e2e": {
            "value": tf(t_e2e), "unit": "TFLOP/s", "h2d_bytes_per_step": int(x.numel() * 8),
            "d2h_bytes_per_step": int(x.numel() * 8), "ms_per_step": t_e2e * 1e3,
            "api": "Environment.one_site_matvec(site).matvec(x_host), use_identity_channels off (general chain)",
            "max_abs_diff_vs_device_path": parity,
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "canonical_gauge": canonical,
        "lanczos_step": lanczos,
        "native_fp64": {
            "ms_per_step": t_f * 1e3, "tflops": tf(t_f), "roofline": roofline_dmma,
            "note": "the same general chain with TNPY_GEMM_ALGO=fp64: gemm_tn_dmma (mma.sync DMMA) -> wmix -> gemm_tn_dmma",
        },
    }
    del plan_g, plan_c, L, R, x, y, y_general, y_fp64, env, dmrg
    torch.cuda.empty_cache()
    if args.sweep_n > 0:
        line["sweep_measured"] = measure_sweeps(args.sweep_n, chi, args.tol, args.sweep_count)
    if not args.no_readme_run:
        line["readme_config"] = measure_readme_config(args.tol)
    if args.scale_chi > 0:
        line["scaling_reference_n1"] = single_gpu_matvec(args.scale_chi, max(2, args.steps // 4))
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(chi)
    print(json.dumps(line), flush=True)


def measure_sweeps(n, chi, tol, count):
    """Whole sweeps, measured (not extrapolated): XXZ delta=0.5 at the stated (reduced) n and the workload's chi from
    a random right-canonical MPS, default solver / split / GEMM selection, wall seconds with a device synchronise
    on both sides of every sweep.  Sweep 1 is the cold sweep (local solves dominate), later ones are warm."""
    import torch

    from tnpy_b200 import _cuda
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import Direction
    from tnpy_b200.model import XXZ

    tensors = random_right_canonical_device(n, chi, 2, seed=1)
    dmrg = FiniteDMRG(XXZ(n=n, delta=0.5).mpo, bond_dim=chi, mps=tensors, compute_variance=False)
    del tensors
    dmrg.phase_seconds = {}
    out = {"n": n, "chi": chi, "tol": tol, "sweep_s": [], "matvecs": [], "energies": [], "phase_s_cumulative": [],
           "launches": [], "sites_at_full_chi": sum(1 for i in range(n) if dmrg.environment.device_tensor(i).shape[0] == chi
                                                    and dmrg.environment.device_tensor(i).shape[2] == chi)}
    for k in range(count):
        torch.cuda.synchronize()
        l0, t0 = _cuda.launch_count(), time.perf_counter()
        e = dmrg.sweep(Direction.RIGHTWARD if k % 2 == 0 else Direction.LEFTWARD, tol=tol)
        torch.cuda.synchronize()
        out["sweep_s"].append(time.perf_counter() - t0)
        out["launches"].append(_cuda.launch_count() - l0)
        out["matvecs"].append(sum(st.get("n_matvec", 0) for st in dmrg.solver_stats))
        out.setdefault("looks", []).append(sum(st.get("looks", 0) for st in dmrg.solver_stats))
        out.setdefault("extra_gs_passes", []).append(sum(st.get("extra_gs_passes", 0) for st in dmrg.solver_stats))
        out.setdefault("reduced_slice_matvecs", []).append(sum(st.get("reduced_slice_matvecs", 0) for st in dmrg.solver_stats))
        out.setdefault("failed_residual_checks", []).append(sum(st.get("failed_residual_checks", 0) for st in dmrg.solver_stats))
        out.setdefault("five_slice_matvecs", []).append(sum(st.get("five_slice_matvecs", 0) for st in dmrg.solver_stats))
        out["energies"].append(e)
        out["phase_s_cumulative"].append(dict(dmrg.phase_seconds))
    out["split_counts"] = dict(dmrg.environment.split_counts)
    out["note"] = ("measured whole sweeps at reduced n (the full n=100 chi=2048 run is recorded in profiles/, minutes per "
                   "cold sweep); the reference's tol=1e-8 stopping rule per local solve")
    return out


def measure_readme_config(tol):
    """BASELINE configs[0], the reference README's own example, through the public API to convergence: XXZ n=100
    delta=0.5, FiniteDMRG(chi=60).update(tol=1e-8) from the seeded random MPS -- the small-site regime, where whole
    Lanczos steps run in one cooperative launch (csrc/lanczos_steps.cu).  Wall seconds with a device synchronise on
    both sides; the second run is the warm-process number (the first pays one-time module / allocator set-up)."""
    import torch

    from tnpy_b200 import _cuda
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.model import XXZ

    out = {"config": "XXZ n=100 delta=0.5 chi=60 tol=%g to convergence (FiniteDMRG(mpo, chi=60).update(tol))" % tol, "runs": []}
    for _ in range(2):
        dmrg = FiniteDMRG(XXZ(n=100, delta=0.5).mpo, chi=60, seed=0)
        torch.cuda.synchronize()
        l0, t0 = _cuda.launch_count(), time.perf_counter()
        energies = dmrg.update(tol=tol)
        torch.cuda.synchronize()
        out["runs"].append({"seconds": time.perf_counter() - t0, "launches": int(_cuda.launch_count() - l0),
                            "sweeps": len(energies), "energy": energies[-1]})
    out["seconds"] = min(r["seconds"] for r in out["runs"])
    return out


def single_gpu_matvec(chi, steps):
    """The N>1 workload (random L / R at chi, general chain) on ONE GPU, so that the 2/4/8-GPU lines of the same
    bench have a same-workload N=1 point (the driver's N=1 run is the chi=2048 headline)."""
    import torch

    from tnpy_b200 import _cuda
    from tnpy_b200.model import XXZ

    w, d = 5, 2
    g = torch.Generator(device="cuda").manual_seed(99)
    rnd = lambda *s: torch.randn(s, generator=g, dtype=torch.float64, device="cuda")  # noqa: E731
    L, R, x = rnd(chi, w, chi), rnd(chi, w, chi), rnd(chi, d, chi)
    W = torch.from_numpy(np.ascontiguousarray(XXZ(n=4, delta=0.5).mpo.as_four_leg(1))).cuda()
    y = torch.empty_like(x)
    plan = _cuda.HeffPlan(L, W, R, chi, chi)
    sec = timed_plan(plan, x, y, steps, 1, torch.cuda.synchronize)
    mode = plan.mode
    plan.close()
    flops = f_mv(chi, chi, w, w, d)
    return {"workload": workload_name(2, chi).replace("on 2 GPUs", "on 1 GPU (unsharded)"), "chi": chi, "steps": steps,
            "ms_per_step": sec * 1e3, "value": flops / sec / 1e12, "unit": "TFLOP/s", "heff_mode": mode}


# ---------------------------------------------------------------------------------------------------
def run_sharded(args):
    import torch
    import torch.distributed as dist

    from tnpy_b200 import _cuda
    from tnpy_b200.parallel import row_block

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _cuda.load()
    chi = args.chi or 8192
    w, d = 5, 2
    lo0, lo1 = row_block(chi, world, rank)
    rows = lo1 - lo0
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    L_rows = torch.randn((chi, w, rows), generator=g, dtype=torch.float64, device="cuda")
    g2 = torch.Generator(device="cuda").manual_seed(99)
    R = torch.randn((chi, w, chi), generator=g2, dtype=torch.float64, device="cuda")
    from tnpy_b200.model import XXZ

    W = torch.from_numpy(np.ascontiguousarray(XXZ(n=4, delta=0.5).mpo.as_four_leg(1))).cuda()
    x_full = torch.empty((chi, d, chi), dtype=torch.float64, device="cuda")
    x_rows = torch.randn((rows, d, chi), generator=g, dtype=torch.float64, device="cuda")
    y_rows = torch.empty_like(x_rows)
    equal = chi % world == 0
    flops = f_mv(chi, chi, w, w, d)

    if not equal:
        raise SystemExit(f"bench: chi={chi} does not split evenly over {world} ranks")

    plan = _cuda.HeffPlan(L_rows, W, R, chi, chi, l_rows=rows)  # environments sliced once, as in a local solve

    def gather():
        dist.all_gather_into_tensor(x_full, x_rows)

    def step():
        gather()
        plan.apply(x_full, y_rows)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # correctness of the sharded result, once, outside the timed region: rank 0 rebuilds the full L from every
    # rank's seed, runs the unsharded matvec on the native FP64 chain and broadcasts y; every rank compares its rows
    step()
    y_check = torch.empty((chi, d, chi), dtype=torch.float64, device="cuda")
    if rank == 0:
        L_full = torch.empty((chi, w, chi), dtype=torch.float64, device="cuda")
        for g_ in range(world):
            lo_g, hi_g = row_block(chi, world, g_)
            gg = torch.Generator(device="cuda").manual_seed(1234 + g_)
            L_full[:, :, lo_g:hi_g] = torch.randn((chi, w, hi_g - lo_g), generator=gg, dtype=torch.float64, device="cuda")
        ref_plan = _cuda.HeffPlan(L_full, W, R, chi, chi, algo=_cuda.GEMM_FP64)
        ref_plan.apply(x_full, y_check)
        ref_plan.close()
        del L_full, ref_plan
    dist.broadcast(y_check, src=0)
    diff = ((y_rows - y_check[lo0:lo1]).abs().max() / y_check.abs().max()).reshape(1)
    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    max_rel_diff = float(diff.item())
    del y_check
    torch.cuda.empty_cache()
    if not max_rel_diff < 1e-12:
        raise SystemExit(f"bench: sharded matvec differs from the unsharded FP64 chain by {max_rel_diff:.3e} (> 1e-12)")

    # per-GPU FP64 roofline denominator: cuBLAS dgemm 8192^3 on this rank (yardstick only), max over ranks
    a8 = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    b8 = torch.randn(8192, 8192, dtype=torch.float64, device="cuda")
    torch.matmul(a8, b8)
    best = min(cuda_time(lambda: torch.matmul(a8, b8), 1, torch.cuda.synchronize) for _ in range(3))
    peak = torch.tensor([2.0 * 8192**3 / best / 1e12], dtype=torch.float64, device="cuda")
    dist.all_reduce(peak, op=dist.ReduceOp.MAX)
    fp64_peak = float(peak.item())
    del a8, b8

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _cuda.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    dt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    launches = _cuda.launch_count() - launches0

    # compute-only share (no all-gather) for the breakdown
    barrier()
    e0.record()
    for _ in range(args.steps):
        plan.apply(x_full, y_rows)
    e1.record()
    barrier()
    dt_compute = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt_compute, op=dist.ReduceOp.MAX)

    # e2e: each rank's row block comes from / goes back to pinned host memory every step
    xh = torch.empty(x_rows.numel(), dtype=torch.float64).pin_memory()
    xh.copy_(x_rows.reshape(-1))
    yh = torch.empty(x_rows.numel(), dtype=torch.float64).pin_memory()

    def step_e2e():
        x_rows.reshape(-1).copy_(xh, non_blocking=True)
        step()
        yh.copy_(y_rows.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    dt_e2e = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        t, tc, te = float(dt.item()), float(dt_compute.item()), float(dt_e2e.item())
        gemm_flops = 4.0 * w * d * float(chi) ** 3  # the two GEMMs of the chain, all ranks
        if plan.mode != 0:
            pops = 36 * gemm_flops * args.steps / tc / world / 1e15
            roofline = {
                "bound": "tensor", "achieved": pops, "peak": INT8_PEAK_POPS, "unit": "POP/s (int8) per GPU",
                "frac": pops / INT8_PEAK_POPS, "traffic": None,
                "kernel": "local chain per GPU (x / T2 slicing + wmix + 2 x oz2_mma_kernel<8>), all-gather excluded; "
                          "int8 ops = 36 slice pairs x 2MNK of the two GEMMs",
                "peak_source": "ncu sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained at 1965 MHz (profiles/r01_oz2_mma_kernel_ncu_full_raw.csv)",
                "fp64_equivalent_tflops_per_gpu": flops * args.steps / tc / 1e12 / world, "cublas_dgemm_tflops": fp64_peak,
                "ms_compute_per_step": tc / args.steps * 1e3, "ms_allgather_per_step": (t - tc) / args.steps * 1e3,
            }
        else:
            roofline = {
                "bound": "tensor", "achieved": flops * args.steps / tc / 1e12 / world, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": flops * args.steps / tc / 1e12 / world / fp64_peak, "traffic": None,
                "kernel": "local chain per GPU (gemm_tn_dmma x2 + wmix), all-gather excluded",
                "peak_source": "cuBLAS dgemm 8192^3 per GPU, best of 3, max over ranks",
                "ms_compute_per_step": tc / args.steps * 1e3, "ms_allgather_per_step": (t - tc) / args.steps * 1e3,
            }
        line = {
            "metric": "heff_matvec_fp64_tflops", "value": flops * args.steps / t / 1e12, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(world, chi), "chi": chi, "w": w, "d": d, "rows_per_rank": rows,
                "parallelism": f"chi-row blocks of L x{world}; all-gather(x) per step over NCCL; no reduction",
                "flops_per_step": flops, "l2_policy": "operands exceed L2; no flush needed",
                "path": {0: "fp64 chain", 1: "chain, GEMMs on tcgen05 int8 (Ozaki, 8 slices)", 2: "direct"}[plan.mode],
            },
            "max_rel_diff_vs_unsharded_fp64_chain": max_rel_diff,
            "roofline": roofline,
            "e2e": {
                "value": flops * args.steps / te / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(x_rows.numel() * 8 * world),
                "d2h_bytes_per_step": int(x_rows.numel() * 8 * world), "ms_per_step": te / args.steps * 1e3,
            },
            "gpu_launches": int(launches) * world, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chi", type=int, default=0)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--sweep-n", type=int, default=40, help="chain length of the measured whole sweeps at the workload's chi (0 = skip)")
    ap.add_argument("--sweep-count", type=int, default=2, help="number of measured sweeps (the first is the cold one)")
    ap.add_argument("--scale-chi", type=int, default=8192, help="chi of the same-workload single-GPU point of the scaling run (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-readme-run", action="store_true", help="skip the XXZ n=100 chi=60 run to convergence (about 3 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_sharded(args)
    return run_single(args)


if __name__ == "__main__":
    main()
