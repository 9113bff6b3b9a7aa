#!/usr/bin/env python
"""bench.py -- the fDMRG local-update hot path on B200: H_eff matvec FP64 TFLOP/s (+ sweep seconds).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one H_eff.psi contraction at a full-chi bulk site
(the call primme makes ~10-30x per site in the reference, matrix_product_state.py:411-440).

  N = 1   workload = BASELINE.json configs[2]: XXZ n=100 delta=0.5 chi=2048 (w=5, d=2), mid-chain site
          of a synthetic random right-canonical MPS with genuine L / R from the update recursions.
  N > 1   workload = configs[4]: XXZ n=200 chi=8192, the matvec sharded over chi-row blocks of L
          (SURVEY 8e.1): every rank holds L[:, :, rows_g], the full R and x; per step the ranks
          all-gather x over NCCL and run the local chain; no reduction is needed.

`value` is whole-job algorithmic TFLOP/s (F_mv = 4 w d chi^3 + 2 w^2 d^2 chi^2 per matvec)
with operands resident in HBM; `e2e` is the same metric through the reference-facing seam
(``Environment.one_site_matvec(site).matvec(x_host)``): pinned host vector in, host vector out, both
copies inside the timed region.  `roofline` is for the dominant kernel (gemm_tn_dmma) against the
FP64 GEMM rate cuBLAS reaches in the same process (MEASURED_PEAKS.json carries no FP64 figure).
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

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
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


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this step (NumPy restatement in
    oracle/ -- quimb/primme are not installable here) on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with all_host_threads():
        return _run_reference(args)


def _run_reference(args):
    from oracle import tnpy_oracle as oracle

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chi = args.chi or (2048 if args.gpus == 1 else 8192)
    sample_chi = min(chi, args.cpu_chi)
    w, d = 5, 2
    rng = np.random.default_rng(0)
    L = rng.standard_normal((sample_chi, w, sample_chi))
    R = rng.standard_normal((sample_chi, w, sample_chi))
    W = np.ascontiguousarray(oracle.xxz_mpo(4, 0.5)[1])
    x = rng.standard_normal((sample_chi, d, sample_chi))
    for _ in range(max(args.warmup, 1)):
        oracle.heff_apply(L, W, R, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.heff_apply(L, W, R, x)
    dt = time.perf_counter() - t0
    flops = f_mv(sample_chi, sample_chi, w, w, d)
    value = flops * args.steps / dt / 1e12
    cores = os.cpu_count()
    sample = f"{args.steps} NumPy tensordot-chain matvecs at chi={sample_chi} (w=5, d=2) of the chi={chi} workload"
    line = {
        "impl": "reference", "metric": "heff_matvec_fp64_tflops", "value": value, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus, chi), "chi": chi, "cpu_sample_chi": sample_chi},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(n_gpus, chi):
    if n_gpus == 1:
        return f"XXZ n=100 delta=0.5 chi={chi}: H_eff matvec at the mid-chain site (BASELINE configs[2])"
    return f"XXZ n=200 delta=0.5 chi={chi}: H_eff matvec sharded over chi-row blocks of L on {n_gpus} GPUs (BASELINE configs[4])"


def cpu_baseline(chi, budget_s=20.0):
    """Oracle matvec on the host cores, bounded sample (same shapes when they fit the budget)."""
    with all_host_threads():
        return _cpu_baseline(chi, budget_s)


def _cpu_baseline(chi, budget_s):
    from oracle import tnpy_oracle as oracle

    w, d = 5, 2
    sample_chi = min(chi, 1024)
    rng = np.random.default_rng(0)
    L = rng.standard_normal((sample_chi, w, sample_chi))
    R = rng.standard_normal((sample_chi, w, sample_chi))
    W = np.ascontiguousarray(oracle.xxz_mpo(4, 0.5)[1])
    x = rng.standard_normal((sample_chi, d, sample_chi))
    oracle.heff_apply(L, W, R, x)
    t0 = time.perf_counter()
    n = 0
    while True:
        oracle.heff_apply(L, W, R, x)
        n += 1
        if time.perf_counter() - t0 > budget_s * 0.6 or n >= 100:
            break
    dt = (time.perf_counter() - t0) / n
    return {
        "value": f_mv(sample_chi, sample_chi, w, w, d) / dt / 1e12, "unit": "TFLOP/s", "cores": os.cpu_count(),
        "kind": "port", "ms_per_matvec": dt * 1e3,
        "sample": f"{n} NumPy tensordot-chain matvecs (oracle/tnpy_oracle.heff_apply) at chi={sample_chi}, w=5, d=2",
    }


# ---------------------------------------------------------------------------------------------------
def run_single(args):
    import torch

    from tnpy_b200 import _cuda
    from tnpy_b200.finite_dmrg import FiniteDMRG
    from tnpy_b200.matrix_product_state import Direction
    from tnpy_b200.model import XXZ

    torch.cuda.set_device(0)
    lib = _cuda.load()
    chi = args.chi or 2048
    n = args.n or 100
    d = 2
    model = XXZ(n=n, delta=0.5)
    mpo = model.mpo
    sync = torch.cuda.synchronize

    # FP64 roofline denominator measured here: cuBLAS dgemm 8192^3 (burst, best of 5) -- yardstick only
    m8 = 8192 if chi >= 1024 else 4096
    a = torch.randn(m8, m8, dtype=torch.float64, device="cuda")
    b = torch.randn(m8, m8, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    best = min(cuda_time(lambda: torch.matmul(a, b), 1, sync) for _ in range(5))
    fp64_peak = 2.0 * m8**3 / best / 1e12
    del a, b

    t_setup = time.perf_counter()
    tensors = random_right_canonical_device(n, chi, d, seed=0)
    dmrg = FiniteDMRG(mpo, bond_dim=chi, mps=tensors, compute_variance=False)
    env = dmrg.environment
    del tensors
    sync()
    t_setup = time.perf_counter() - t_setup
    site = n // 2
    L, W, R = env.operands(site)
    x = env.device_tensor(site).clone()
    y = torch.empty_like(x)
    l, _, r = x.shape
    wl, wr = W.shape[0], W.shape[1]
    flops = f_mv(l, r, wl, wr, d)

    step = lambda: _cuda.heff_apply(L, W, R, x, y)  # noqa: E731
    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = _cuda.launch_count()
    dt = cuda_time(step, args.steps, sync)
    launches = _cuda.launch_count() - launches0

    # the same step in the gauge the sweep actually runs in: mixed-canonical MPS => L[:,0,:] = R[:,w-1,:] = I,
    # the library replaces those two channel slices of the GEMMs by transposes (executed flops (w-1)/w)
    Lc = L.clone()
    Lc[:, 0, :] = torch.eye(l, dtype=torch.float64, device="cuda")
    Rc = R.clone()
    Rc[:, wr - 1, :] = torch.eye(r, dtype=torch.float64, device="cuda")
    both = _cuda.LEFT_IDENTITY | _cuda.RIGHT_IDENTITY
    step_c = lambda: _cuda.heff_apply(Lc, W, Rc, x, y, flags=both)  # noqa: E731
    y_ref = _cuda.heff_apply(Lc, W, Rc, x).clone()
    for _ in range(args.warmup):
        step_c()
    dt_c = cuda_time(step_c, args.steps, sync)
    gauge_diff = float((y - y_ref).abs().max() / y_ref.abs().max())
    _cuda.heff_apply(L, W, R, x, y)
    del Lc, Rc, y_ref

    # the tcgen05 path: same matvec with the GEMMs on the int8 tensor cores (Ozaki scheme, 8 slices)
    _cuda.set_gemm_algo(_cuda.GEMM_OZAKI)
    try:
        y_oz = _cuda.heff_apply(L, W, R, x).clone()
        step_o = lambda: _cuda.heff_apply(L, W, R, x, y)  # noqa: E731
        for _ in range(args.warmup):
            step_o()
        dt_o = cuda_time(step_o, args.steps, sync)
        Lc = L.clone()
        Lc[:, 0, :] = torch.eye(l, dtype=torch.float64, device="cuda")
        Rc = R.clone()
        Rc[:, wr - 1, :] = torch.eye(r, dtype=torch.float64, device="cuda")
        step_oc = lambda: _cuda.heff_apply(Lc, W, Rc, x, y, flags=both)  # noqa: E731
        step_oc()
        dt_oc = cuda_time(step_oc, args.steps, sync)
        # as the eigensolver runs it: L and R declared constant, their slices made once per solve
        _cuda.ozaki_const_scope(True)
        try:
            y_sc = _cuda.heff_apply(L, W, R, x).clone()
            scope_diff = float((y_sc - y_oz).abs().max())
            del y_sc
            step_o()
            dt_os = cuda_time(step_o, args.steps, sync)
            step_oc()
            dt_ocs = cuda_time(step_oc, args.steps, sync)
            # 7 slices (28 slice GEMMs): what tnpy_eig_lowest selects when its tolerance is >= 1e-10
            _cuda.set_ozaki_slices(7)
            y_s7 = _cuda.heff_apply(L, W, R, x).clone()
            s7_diff = float((y_s7 - y_oz).abs().max() / y_oz.abs().max())
            del y_s7
            step_oc()
            dt_ocs7 = cuda_time(step_oc, args.steps, sync)
        finally:
            _cuda.set_ozaki_slices(8)
            _cuda.ozaki_const_scope(False)
        del Lc, Rc
    finally:
        _cuda.set_gemm_algo(_cuda.GEMM_AUTO)
    _cuda.heff_apply(L, W, R, x, y)
    ozaki_diff = float((y_oz - y).abs().max() / y.abs().max())
    del y_oz

    # dominant kernel alone: the two gemm_tn_dmma launches of the chain, timed with events per launch group
    ws = torch.empty(lib.tnpy_heff_workspace_bytes(l, r, wl, wr, d), dtype=torch.uint8, device="cuda")
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
    del t1, t2, yq, ws

    # e2e through the reference-facing seam: host (pinned) vector in, host vector out
    op = env.one_site_matvec(site)
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
    dt_e2e = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
    clocks = sampler.stop()
    parity = float(np.abs(y_np - y.reshape(-1).cpu().numpy()).max())

    # local-update breakdown at mid-chain sites (what a sweep is made of)
    sweep = sweep_oz = None
    if args.sweep_sites > 0:
        sweep = measure_local_updates(dmrg, site, args.sweep_sites, Direction.RIGHTWARD, args.tol)
        # the same local updates with the chains' GEMMs on the tcgen05 path (the next sites of the same sweep)
        _cuda.set_gemm_algo(_cuda.GEMM_OZAKI)
        try:
            sweep_oz = measure_local_updates(dmrg, site + args.sweep_sites, args.sweep_sites, Direction.RIGHTWARD, args.tol)
        finally:
            _cuda.set_gemm_algo(_cuda.GEMM_AUTO)

    # the tcgen05 kernel alone on the two GEMM shapes of the step (operands already sliced): int8 roofline
    t1 = torch.empty((d * r, wl * l), dtype=torch.float64, device="cuda")
    t2 = torch.randn((r * wr, d * l), dtype=torch.float64, device="cuda")
    yq = torch.empty((d * l, r), dtype=torch.float64, device="cuda")
    oz_ms = []
    for (am, bm, cm) in ((x.reshape(l, d * r), L.reshape(l, wl * l), t1), (t2, R.reshape(r * wr, r), yq)):
        _cuda.ozaki_gemm_tn(am, bm, out=cm, slices=8, phase=1)
        mm = lambda: _cuda.ozaki_gemm_tn(am, bm, out=cm, slices=8, phase=2)  # noqa: E731
        mm()
        oz_ms.append(cuda_time(mm, args.steps, sync) / args.steps * 1e3)
    del t1, t2, yq
    int8_peak = 4.48  # POP/s: ncu sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained at 1965 MHz (profiles/)
    oz_pops = 36 * gemm_flops / (sum(oz_ms) * 1e-3) / 1e15

    value = flops * args.steps / dt / 1e12
    gemm_flops_oz = 2.0 * (d * r) * (wl * l) * l + 2.0 * (d * l) * r * (r * wr)
    line = {
        "metric": "heff_matvec_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": workload_name(1, chi), "n": n, "chi": chi, "w": wl, "d": d, "site": site,
            "flops_per_step": flops, "l2_policy": "operands (L 168 MB + R 168 MB + x + T1/T2 670 MB at chi=2048) exceed the 126 MB L2; no flush needed",
            "setup_s": t_setup,
        },
        "roofline": {
            "bound": "tensor", "achieved": gemm_flops / (tg1 + tg3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": gemm_flops / (tg1 + tg3) / 1e12 / fp64_peak,
            "traffic": 1.29e9 if chi == 2048 else None,  # dram read + write per launch (mean of the two), ncu --set full (profiles/r01_gemm_tn_dmma_ncu_full_raw.csv)
            "kernel": "gemm_tn_dmma (2 launches per matvec)", "ms_gemm1": tg1 * 1e3, "ms_gemm3": tg3 * 1e3,
            "peak_source": f"cuBLAS dgemm {m8}^3 via torch.matmul, best of 5 in this process (MEASURED_PEAKS.json has no FP64 entry)",
            "whole_step_frac": value / fp64_peak,
        },
        "e2e": {
            "value": flops * args.steps / dt_e2e / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(x.numel() * 8),
            "d2h_bytes_per_step": int(x.numel() * 8), "ms_per_step": dt_e2e / args.steps * 1e3,
            "api": "Environment.one_site_matvec(site).matvec(x_host)", "max_abs_diff_vs_device_path": parity,
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "tcgen05_ozaki": {
            "ms_per_step": dt_o / args.steps * 1e3, "tflops_fp64_equivalent": flops * args.steps / dt_o / 1e12,
            "ms_per_step_canonical_gauge": dt_oc / args.steps * 1e3,
            "tflops_fp64_equivalent_canonical_gauge": flops * args.steps / dt_oc / 1e12,
            "ms_per_step_const_env": dt_os / args.steps * 1e3,
            "tflops_fp64_equivalent_const_env": flops * args.steps / dt_os / 1e12,
            "ms_per_step_canonical_gauge_const_env": dt_ocs / args.steps * 1e3,
            "tflops_fp64_equivalent_canonical_gauge_const_env": flops * args.steps / dt_ocs / 1e12,
            "max_abs_diff_const_env_vs_resliced": scope_diff,
            "ms_per_step_canonical_gauge_const_env_7_slices": dt_ocs7 / args.steps * 1e3,
            "tflops_fp64_equivalent_canonical_gauge_const_env_7_slices": flops * args.steps / dt_ocs7 / 1e12,
            "rel_diff_7_vs_8_slices": s7_diff,
            "rel_diff_vs_dmma_path": ozaki_diff, "slices": 8,
            "int8_pops_const_env": 36 * gemm_flops_oz * args.steps / dt_os / 1e15,
            "roofline": {
                "bound": "tensor", "achieved": oz_pops, "peak": int8_peak, "unit": "POP/s (int8)", "frac": oz_pops / int8_peak,
                "traffic": 1.63e9, "kernel": "oz2_mma_kernel<8> (2 launches per matvec)", "ms_gemm1": oz_ms[0],
                "ms_gemm3": oz_ms[1], "fp64_equivalent_tflops": gemm_flops / (sum(oz_ms) * 1e-3) / 1e12,
                "peak_source": "ncu sm__ops_path_tensor_op_utcimma_src_int8 peak_sustained (profiles/r01_oz2_mma_kernel_ncu_full_raw.csv); "
                               "traffic = dram read + write per launch (mean of the two shapes) from the same capture",
            },
            "note": "same matvec with both GEMMs as 36 exact int8 slice GEMMs on tcgen05 (CTA pairs, TMEM int32 "
                    "accumulators); first figures re-slice every operand on every call, *_const_env keep the slices "
                    "of L and R for the duration of a scope as tnpy_eig_lowest does; int8_pops = 36 x 2MNK of both "
                    "GEMMs / whole-step time (ncu peak 4.48 POP/s at 1965 MHz); opt-in via TNPY_GEMM_ALGO=ozaki",
        },
        "canonical_gauge": {
            "ms_per_step": dt_c / args.steps * 1e3, "tflops_algorithmic": flops * args.steps / dt_c / 1e12,
            "executed_flop_fraction": (wl - 1) / wl if wl == wr else None, "rel_diff_vs_dense_path": gauge_diff,
            "note": "same matvec with L[:,0,:] = R[:,w-1,:] = I flagged (the state of every site inside a sweep); "
                    "`value` above is the general dense path",
        },
    }
    if sweep is not None:
        line["sweep"] = sweep
    if sweep_oz is not None:
        line["sweep_tcgen05"] = sweep_oz
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(chi)
    print(json.dumps(line), flush=True)


def measure_local_updates(dmrg, site, count, direction, tol):
    """Full local updates (eigensolve + perturb + split + env update) at consecutive mid-chain
    sites; per-phase device time and an extrapolated sweep time (sum over sites of the measured
    per-site time scaled by F_mv(site) / F_mv(mid)).  `split` is what the sweep runs (verified
    Cholesky-QR split -- two passes, then the shifted three-pass variant, then the SVD, whichever the
    device-side orthogonality check accepted first, all attempts inside the timed phase); the reference-literal Jacobi SVD split
    of the same tensor is timed beside it on a copy (`svd_split_reference_gauge`, not part of the total)."""
    import torch

    from tnpy_b200.matrix_product_state import Direction, _split_on_device

    env = dmrg.environment
    sync = torch.cuda.synchronize
    phases = {"eigensolve": 0.0, "perturb": 0.0, "split": 0.0, "env_update": 0.0}
    matvecs = 0
    svd_gauge_s = 0.0
    counts0 = dict(env.split_counts)
    for s in range(site, site + count):
        sync(); t = time.perf_counter()
        dmrg._solve_on_device(s, tol)
        sync(); phases["eigensolve"] += time.perf_counter() - t; t = time.perf_counter()
        matvecs += dmrg.solver_stats[-1].get("n_matvec", 0)
        dmrg.perturb_wave_function(s)
        sync(); phases["perturb"] += time.perf_counter() - t; t = time.perf_counter()
        nb_site = s + 1 if direction == Direction.RIGHTWARD else s - 1
        sync(); t = time.perf_counter()
        _split_on_device(env.device_tensor(s), env.device_tensor(nb_site), direction, "svd")
        sync(); svd_gauge_s += time.perf_counter() - t; t = time.perf_counter()
        env.split_tensor(s, direction)
        sync(); phases["split"] += time.perf_counter() - t; t = time.perf_counter()
        env.update(s, direction)
        sync(); phases["env_update"] += time.perf_counter() - t
    per_site = {k: v / count for k, v in phases.items()}
    total = sum(per_site.values())
    n = env.n_sites
    shapes = [tuple(env.device_tensor(i).shape) for i in range(n)]
    w = env.operands(site)[1].shape[0]
    mid = f_mv(shapes[site][0], shapes[site][2], w, w, 2)
    weight = sum(f_mv(sh[0], sh[2], w if i else 1, w if i < n - 1 else 1, 2) for i, sh in enumerate(shapes[:-1])) / mid
    return {
        "sites_measured": count, "per_site_s": per_site, "per_site_total_s": total,
        "matvecs_per_site": matvecs / count, "sweep_s_extrapolated": total * weight,
        "svd_split_reference_gauge_s": svd_gauge_s / count,
        "sweep_s_extrapolated_svd_gauge": (total - per_site["split"] + svd_gauge_s / count) * weight,
        "splits": {k: env.split_counts[k] - counts0[k] for k in counts0},
        "note": "extrapolated = per-site total x sum_sites F_mv(site)/F_mv(mid); eigensolver tol %.0e" % tol,
    }


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

    def gather():
        dist.all_gather_into_tensor(x_full, x_rows)

    def step():
        gather()
        _cuda.heff_apply_rows(L_rows, W, R, x_full, y_rows)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

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
        _cuda.heff_apply_rows(L_rows, W, R, x_full, y_rows)
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
        line = {
            "metric": "heff_matvec_fp64_tflops", "value": flops * args.steps / t / 1e12, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(world, chi), "chi": chi, "w": w, "d": d, "rows_per_rank": rows,
                "parallelism": f"chi-row blocks of L x{world}; all-gather(x) per step over NCCL; no reduction",
                "flops_per_step": flops, "l2_policy": "operands exceed L2; no flush needed",
            },
            "roofline": {
                "bound": "tensor", "achieved": flops * args.steps / tc / 1e12 / world, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": flops * args.steps / tc / 1e12 / world / fp64_peak, "traffic": None,
                "kernel": "local chain per GPU (gemm_tn_dmma x2 + wmix), all-gather excluded",
                "peak_source": "cuBLAS dgemm 8192^3 per GPU, best of 3, max over ranks",
                "ms_compute_per_step": tc / args.steps * 1e3, "ms_allgather_per_step": (t - tc) / args.steps * 1e3,
            },
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
    ap.add_argument("--sweep-sites", type=int, default=2, help="mid-chain local updates to time for the sweep estimate (0 = skip)")
    ap.add_argument("--cpu-chi", type=int, default=1024, help="--impl reference: chi of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return run_sharded(args)
    return run_single(args)


if __name__ == "__main__":
    main()
