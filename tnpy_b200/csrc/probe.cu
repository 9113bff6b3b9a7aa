// Hardware probes used to fix the FP64 roofline denominator on the box the bench runs on
// (MEASURED_PEAKS.json carries HBM and BF16 figures only).  Register-resident loops: no memory traffic.
#include "common.cuh"

namespace tnpy {

template <int ILP>
__global__ void __launch_bounds__(1024) dmma_probe_kernel(double* out, int iters, double seed) {
  double c[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = seed * (i + 1);
  double a = seed + threadIdx.x * 1e-9, b = seed - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(1024) dfma_probe_kernel(double* out, int iters, double seed) {
  double c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i] = seed * (i + 1);
  const double a = 1.0 + seed * 1e-9, b = seed * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

}  // namespace tnpy

using namespace tnpy;

// kind 0: DMMA m8n8k4, kind 1: DFMA.  Returns achieved TFLOP/s (2 flop per FMA) through *tflops.
extern "C" int tnpy_probe_fp64(int kind, int threads_per_block, int blocks_per_sm, int ilp, int iters, double* tflops,
                               void* scratch_dev) {
  TNPY_CHECK_ARG(tflops && scratch_dev && threads_per_block % 32 == 0 && threads_per_block <= 1024, "bad argument");
  const int blocks = sm_count() * blocks_per_sm;
  cudaEvent_t e0, e1;
  TNPY_CUDA_OK(cudaEventCreate(&e0));
  TNPY_CUDA_OK(cudaEventCreate(&e1));
  double* out = static_cast<double*>(scratch_dev);
  auto launch = [&](int n) {
    if (kind == 0) {
      if (ilp <= 4) dmma_probe_kernel<4><<<blocks, threads_per_block>>>(out, n, 1.0);
      else if (ilp <= 8) dmma_probe_kernel<8><<<blocks, threads_per_block>>>(out, n, 1.0);
      else if (ilp <= 16) dmma_probe_kernel<16><<<blocks, threads_per_block>>>(out, n, 1.0);
      else dmma_probe_kernel<32><<<blocks, threads_per_block>>>(out, n, 1.0);
    } else {
      if (ilp <= 4) dfma_probe_kernel<4><<<blocks, threads_per_block>>>(out, n, 1.0);
      else if (ilp <= 8) dfma_probe_kernel<8><<<blocks, threads_per_block>>>(out, n, 1.0);
      else dfma_probe_kernel<16><<<blocks, threads_per_block>>>(out, n, 1.0);
    }
  };
  const int eff_ilp = kind == 0 ? (ilp <= 4 ? 4 : ilp <= 8 ? 8 : ilp <= 16 ? 16 : 32) : (ilp <= 4 ? 4 : ilp <= 8 ? 8 : 16);
  launch(iters / 10 + 1);
  TNPY_CUDA_OK(cudaDeviceSynchronize());
  TNPY_CUDA_OK(cudaEventRecord(e0));
  launch(iters);
  TNPY_CUDA_OK(cudaEventRecord(e1));
  TNPY_CUDA_OK(cudaEventSynchronize(e1));
  TNPY_LAUNCH_OK();
  float ms = 0.f;
  TNPY_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  const double fma_per_thread_instr = kind == 0 ? 256.0 / 32.0 : 1.0;
  const double flops = 2.0 * fma_per_thread_instr * (double)eff_ilp * iters * (double)threads_per_block * blocks;
  *tflops = flops / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return TNPY_OK;
}

static int g_forced_tile = -1;
extern "C" int tnpy_set_gemm_tile(int cfg) {
  g_forced_tile = cfg;
  return TNPY_OK;
}
namespace tnpy {
int forced_gemm_tile() { return g_forced_tile; }
}
