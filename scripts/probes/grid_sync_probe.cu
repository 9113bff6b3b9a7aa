// How long does a grid-wide barrier take on this GPU?  cooperative_groups grid.sync() against a hand-written
// arrive/spin barrier and, for up to 16 CTAs, the hardware cluster barrier.  Decides how the fused small-site Lanczos
// steps (tnpy_b200/csrc/lanczos_steps.cu) synchronise their phases.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/grid_sync_probe scripts/probes/grid_sync_probe.cu && /tmp/grid_sync_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256, 1) cg_kernel(int iters, double* sink) {
  cg::grid_group grid = cg::this_grid();
  double v = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    v = v * 1.0000001 + 1.0;
    grid.sync();
  }
  if (v == -1.0) sink[0] = v;
}

__device__ __forceinline__ void spin_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) spin_kernel(int iters, unsigned* counter, double* sink) {
  double v = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    v = v * 1.0000001 + 1.0;
    spin_barrier(counter, (unsigned)(i + 1) * gridDim.x);
  }
  if (v == -1.0) sink[0] = v;
}

__global__ void __launch_bounds__(256, 1) cluster_kernel(int iters, double* sink) {
  double v = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    v = v * 1.0000001 + 1.0;
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (v == -1.0) sink[0] = v;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0;
  cudaEventSynchronize(b);
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  double* sink;
  unsigned* counter;
  cudaMalloc(&sink, 64);
  cudaMalloc(&counter, 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;
  const int grids[] = {8, 16, 41, 64, 129, 148};
  for (int g : grids) {
    int it = iters;
    void* args[] = {&it, &sink};
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      cudaLaunchCooperativeKernel((void*)cg_kernel, dim3(g), dim3(256), args, 0, 0);
      cudaEventRecord(e1);
    }
    const float cg_ms = time_ms(e0, e1);
    float spin_ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(counter, 0, 4);
      void* sargs[] = {&it, &counter, &sink};
      cudaEventRecord(e0);
      cudaLaunchCooperativeKernel((void*)spin_kernel, dim3(g), dim3(256), sargs, 0, 0);
      cudaEventRecord(e1);
      spin_ms = time_ms(e0, e1);
    }
    printf("{\"ctas\": %d, \"cg_grid_sync_us\": %.3f, \"spin_barrier_us\": %.3f}\n", g, 1e3 * cg_ms / iters, 1e3 * spin_ms / iters);
  }
  for (int c : {8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c);
    cfg.blockDim = dim3(256);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (c > 8) cudaFuncSetAttribute(cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    float ms = 0;
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      err = cudaLaunchKernelEx(&cfg, cluster_kernel, iters, sink);
      cudaEventRecord(e1);
      ms = time_ms(e0, e1);
    }
    printf("{\"cluster_ctas\": %d, \"cluster_barrier_us\": %.3f, \"launch\": \"%s\"}\n", c, 1e3 * ms / iters, cudaGetErrorString(err));
  }
  printf("last error: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
