// Shared helpers for the tnpy_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/tnpy_cuda.h"

namespace tnpy {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define TNPY_CHECK_ARG(cond, msg)                                      \
  do {                                                                 \
    if (!(cond)) {                                                     \
      ::tnpy::set_error("%s: invalid argument: %s", __func__, msg);    \
      return TNPY_EINVAL;                                              \
    }                                                                  \
  } while (0)

#define TNPY_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::tnpy::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(_e));       \
      return TNPY_ECUDA;                                                                     \
    }                                                                                        \
  } while (0)

#define TNPY_LAUNCH_OK()                                                                     \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ::tnpy::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e));   \
      return TNPY_ECUDA;                                                                     \
    }                                                                                        \
    ::tnpy::count_launch();                                                                  \
  } while (0)

#define TNPY_TRY(expr)           \
  do {                           \
    int _rc = (expr);            \
    if (_rc != TNPY_OK) return _rc; \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int sm_count();  // of the current device
int current_device();
// true exactly once per (key, current device): one-time per-device setup such as cudaFuncSetAttribute
bool first_time_on_device(const void* key);
// pinned host scratch of the calling thread (>= 64 bytes, lives as long as the thread): status read-backs
void* thread_pinned_scratch();

template <typename Kernel>
inline int set_max_dynamic_smem(Kernel kernel, int bytes);

// Bump allocator over a caller-supplied workspace (256-byte aligned slices).
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {
    size_t mis = reinterpret_cast<uintptr_t>(p) & 255;
    if (mis) used = 256 - mis;
  }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (base == nullptr || used + bytes > size) return nullptr;
    T* out = reinterpret_cast<T*>(base + used);
    used += bytes;
    return out;
  }
  static size_t need(size_t count, size_t elem = sizeof(double)) { return align_up(count * elem, 256); }
};

// ---- internal cross-file entry points --------------------------------------------------------
// C[m,n] (+)= sum_k A[k,m] B[k,n] with an optional "split-M" row map on the output:
// output row of GEMM row index m is (m / m_inner) * c_outer + (m % m_inner) * c_inner (in elements).
struct GemmOut {
  double* C;
  int64_t c_inner;  // element stride between consecutive m inside one inner block
  int64_t c_outer;  // element stride between inner blocks
  int m_inner;      // extent of the inner block (== M for a plain matrix)
};
int gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
            int accumulate, int algo, cudaStream_t stream);
// the same with scratch for split-K partials (gemm_splitk_doubles(M, N, K) doubles, 0 when the shape is not split:
// few output tiles and a long K); without enough scratch the product runs unsplit
int gemm_tn_ws(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
               int accumulate, int algo, double* scratch, size_t scratch_doubles, cudaStream_t stream);
size_t gemm_splitk_doubles(int M, int N, int K);
inline GemmOut plain_out(double* C, int64_t ldc, int M) { return GemmOut{C, ldc, 0, M}; }
int current_gemm_algo();

// device constant 1.0 used for NULL (unit) environments (per device)
const double* device_one();

template <typename Kernel>
inline int set_max_dynamic_smem(Kernel kernel, int bytes) {
  if (first_time_on_device(reinterpret_cast<const void*>(kernel))) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize = %d) failed: %s", bytes, cudaGetErrorString(e));
      return TNPY_ECUDA;
    }
  }
  return TNPY_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-reduce one value per thread; the result is valid in warp 0 (all lanes).  `sh` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}
#endif

}  // namespace tnpy
