// Bandwidth-bound vector kernels for the on-device eigensolver (replace the host BLAS-1/2 inside
// primme, reference linalg.py:86).  All are HBM-roofline kernels: 16-byte vectorised, coalesced,
// grid sized as a multiple of the SM count, warp-shuffle + shared-memory block reductions, and a
// deterministic fixed-order final reduction by the last block to finish (no floating-point atomics,
// so results are bit-reproducible run to run).
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace tnpy {

constexpr int kRedThreads = 256;
constexpr int kMaxRedBlocks = 1184;  // 148 SMs x 8
constexpr int kMaxMulti = 64;        // max number of basis vectors in one multi_dot / multi_axpy

struct RedScratch {
  double* partials;       // [kMaxMulti][kMaxRedBlocks]
  unsigned int* counter;  // last-block-done ticket
};

// The BLAS-1 entry points take no workspace, so the partial sums and the ticket counter of their two-stage
// reductions live in a library-owned scratch (0.6 MB), one per (device, stream): kernels of different streams never
// share partials or tickets, and a second device gets its own allocation.
static RedScratch red_scratch(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, RedScratch> table;
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair(current_device(), stream);
  auto it = table.find(key);
  if (it != table.end()) return it->second;
  RedScratch s{nullptr, nullptr};
  void* p = nullptr;
  if (cudaMalloc(&p, sizeof(double) * kMaxMulti * kMaxRedBlocks + 256) == cudaSuccess) {
    s.partials = static_cast<double*>(p);
    s.counter = reinterpret_cast<unsigned int*>(s.partials + (size_t)kMaxMulti * kMaxRedBlocks);
    if (cudaMemset(s.counter, 0, 256) != cudaSuccess) s = RedScratch{nullptr, nullptr};
  }
  if (s.partials) table.emplace(key, s);
  return s;
}

// Grid of the streaming kernels: one double2 per thread and trip for short vectors (a chi = 60 site has 7200
// entries: with eight per thread only four CTAs worked and multi_dot took 23 us), up to eight per thread for long
// ones, capped at eight CTAs per SM.
static int red_blocks(int64_t n) {
  int64_t b = (n + (int64_t)kRedThreads * 2 - 1) / ((int64_t)kRedThreads * 2);
  const int cap = sm_count() * 8 < kMaxRedBlocks ? sm_count() * 8 : kMaxRedBlocks;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// h[j] = sum_i V[j][i] * w[i]   for j < m  (m == 1 and V == w gives a squared norm)
// mode: 0 plain, 1 sqrt of the result (nrm2)
template <int MB>
__global__ void __launch_bounds__(kRedThreads) multi_dot_kernel(const double* __restrict__ V, int64_t ldv, int m,
                                                                const double* __restrict__ w, int64_t n,
                                                                double* __restrict__ h, double* __restrict__ partials,
                                                                unsigned int* __restrict__ counter, int mode,
                                                                const int* __restrict__ skip) {
  __shared__ double sh[32];
  __shared__ bool is_last;
  if (skip != nullptr && *skip != 0) return;  // uniform over the grid: the caller decided on the device
  const int64_t n2 = n / 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int j0 = 0; j0 < m; j0 += MB) {
    double acc[MB];
#pragma unroll
    for (int j = 0; j < MB; ++j) acc[j] = 0.0;
    // two trips per iteration: twice the loads in flight per thread (a single vector pair -- dot, nrm2 -- otherwise
    // has one 16-byte load per operand outstanding and reached 0.4 of the HBM rate)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n2; i += 2 * stride) {
      const double2 wa = reinterpret_cast<const double2*>(w)[i];
      const double2 wb = reinterpret_cast<const double2*>(w)[i + stride];
      double2 va[MB], vb[MB];
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        const bool on = j0 + j < m;
        va[j] = on ? *reinterpret_cast<const double2*>(V + (int64_t)(j0 + j) * ldv + 2 * i) : make_double2(0.0, 0.0);
        vb[j] = on ? *reinterpret_cast<const double2*>(V + (int64_t)(j0 + j) * ldv + 2 * (i + stride)) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        acc[j] = fma(va[j].x, wa.x, acc[j]);
        acc[j] = fma(va[j].y, wa.y, acc[j]);
        acc[j] = fma(vb[j].x, wb.x, acc[j]);
        acc[j] = fma(vb[j].y, wb.y, acc[j]);
      }
    }
    for (; i < n2; i += stride) {
      const double2 wv = reinterpret_cast<const double2*>(w)[i];
#pragma unroll
      for (int j = 0; j < MB; ++j) {
        if (j0 + j < m) {
          const double2 vv = *reinterpret_cast<const double2*>(V + (int64_t)(j0 + j) * ldv + 2 * i);
          acc[j] = fma(vv.x, wv.x, acc[j]);
          acc[j] = fma(vv.y, wv.y, acc[j]);
        }
      }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
      for (int j = 0; j < MB; ++j)
        if (j0 + j < m) acc[j] = fma(V[(int64_t)(j0 + j) * ldv + n - 1], w[n - 1], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < MB; ++j) {
      const double r = block_sum(acc[j], sh);
      if (threadIdx.x == 0 && j0 + j < m) partials[(int64_t)(j0 + j) * gridDim.x + blockIdx.x] = r;
    }
  }
  // last block to finish reduces the per-block partials in a fixed order
  __threadfence();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int j = 0; j < m; ++j) {
      double a = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a += partials[(int64_t)j * gridDim.x + b];
      const double r = block_sum(a, sh);
      if (threadIdx.x == 0) h[j] = mode == 1 ? sqrt(r) : r;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// Unvectorised variant for odd strides / unaligned bases.
__global__ void __launch_bounds__(kRedThreads) multi_dot_scalar_kernel(const double* __restrict__ V, int64_t ldv, int m,
                                                                       const double* __restrict__ w, int64_t n,
                                                                       double* __restrict__ h,
                                                                       double* __restrict__ partials,
                                                                       unsigned int* __restrict__ counter, int mode,
                                                                       const int* __restrict__ skip) {
  __shared__ double sh[32];
  __shared__ bool is_last;
  if (skip != nullptr && *skip != 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int j = 0; j < m; ++j) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      acc = fma(V[(int64_t)j * ldv + i], w[i], acc);
    const double r = block_sum(acc, sh);
    if (threadIdx.x == 0) partials[(int64_t)j * gridDim.x + blockIdx.x] = r;
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    for (int j = 0; j < m; ++j) {
      double a = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a += partials[(int64_t)j * gridDim.x + b];
      const double r = block_sum(a, sh);
      if (threadIdx.x == 0) h[j] = mode == 1 ? sqrt(r) : r;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

int multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h, int mode,
              cudaStream_t stream, const int* skip) {
  TNPY_CHECK_ARG(V && w && h && n > 0 && m > 0 && m <= kMaxMulti, "bad argument");
  const RedScratch s = red_scratch(stream);
  if (!s.partials) {
    set_error("multi_dot: could not allocate reduction scratch");
    return TNPY_ECUDA;
  }
  const int blocks = red_blocks(n);
  const bool vec = (reinterpret_cast<uintptr_t>(V) % 16 == 0) && (reinterpret_cast<uintptr_t>(w) % 16 == 0) &&
                   (ldv % 2 == 0 || m == 1);
  if (!vec)
    multi_dot_scalar_kernel<<<blocks, kRedThreads, 0, stream>>>(V, ldv, m, w, n, h, s.partials, s.counter, mode, skip);
  else if (m == 1)
    multi_dot_kernel<1><<<blocks, kRedThreads, 0, stream>>>(V, ldv, m, w, n, h, s.partials, s.counter, mode, skip);
  else if (m == 2)
    multi_dot_kernel<2><<<blocks, kRedThreads, 0, stream>>>(V, ldv, m, w, n, h, s.partials, s.counter, mode, skip);
  else
    multi_dot_kernel<4><<<blocks, kRedThreads, 0, stream>>>(V, ldv, m, w, n, h, s.partials, s.counter, mode, skip);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// w[i] -= sum_j h[j] V[j][i];  optionally also accumulates ||w_new||^2 -> nrm (sqrt taken) in the same pass
template <bool WITH_NORM>
__global__ void __launch_bounds__(kRedThreads) multi_axpy_kernel(const double* __restrict__ V, int64_t ldv, int m,
                                                                 const double* __restrict__ h, double* __restrict__ w,
                                                                 int64_t n, int vec, double* __restrict__ nrm,
                                                                 double* __restrict__ partials,
                                                                 unsigned int* __restrict__ counter,
                                                                 const int* __restrict__ skip) {
  __shared__ double hs[kMaxMulti];
  __shared__ double sh[32];
  __shared__ bool is_last;
  if (skip != nullptr && *skip != 0) return;
  for (int j = threadIdx.x; j < m; j += blockDim.x) hs[j] = h[j];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double nacc = 0.0;
  if (vec) {
    const int64_t n2 = n / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 wv = reinterpret_cast<double2*>(w)[i];
      for (int j = 0; j < m; ++j) {
        const double2 vv = *reinterpret_cast<const double2*>(V + (int64_t)j * ldv + 2 * i);
        wv.x = fma(-hs[j], vv.x, wv.x);
        wv.y = fma(-hs[j], vv.y, wv.y);
      }
      reinterpret_cast<double2*>(w)[i] = wv;
      if (WITH_NORM) nacc = fma(wv.x, wv.x, fma(wv.y, wv.y, nacc));
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double x = w[n - 1];
      for (int j = 0; j < m; ++j) x = fma(-hs[j], V[(int64_t)j * ldv + n - 1], x);
      w[n - 1] = x;
      if (WITH_NORM) nacc = fma(x, x, nacc);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      double x = w[i];
      for (int j = 0; j < m; ++j) x = fma(-hs[j], V[(int64_t)j * ldv + i], x);
      w[i] = x;
      if (WITH_NORM) nacc = fma(x, x, nacc);
    }
  }
  if (WITH_NORM) {
    const double r = block_sum(nacc, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = r;
    __threadfence();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      double a = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) a += partials[b];
      const double t = block_sum(a, sh);
      if (threadIdx.x == 0) {
        *nrm = sqrt(t);
        *counter = 0u;
      }
    }
  }
}

int multi_axpy(const double* V, int64_t ldv, int m, const double* h, double* w, int64_t n, double* nrm_out,
               cudaStream_t stream, const int* skip) {
  TNPY_CHECK_ARG(V && h && w && n > 0 && m > 0 && m <= kMaxMulti, "bad argument");
  const RedScratch s = red_scratch(stream);
  if (!s.partials) {
    set_error("multi_axpy: could not allocate reduction scratch");
    return TNPY_ECUDA;
  }
  const int blocks = red_blocks(n);
  const int vec = (reinterpret_cast<uintptr_t>(V) % 16 == 0) && (reinterpret_cast<uintptr_t>(w) % 16 == 0) &&
                  (ldv % 2 == 0 || m == 1);
  if (nrm_out)
    multi_axpy_kernel<true><<<blocks, kRedThreads, 0, stream>>>(V, ldv, m, h, w, n, vec, nrm_out, s.partials, s.counter, skip);
  else
    multi_axpy_kernel<false><<<blocks, kRedThreads, 0, stream>>>(V, ldv, m, h, w, n, vec, nullptr, s.partials, s.counter, skip);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// y = (*num / *den or 1) * alpha * x + beta * y  style helpers -------------------------------------
// out[i] = x[i] * scale, scale = alpha * (s_dev ? (inv ? 1 / *s_dev : *s_dev) : 1);  out may alias x.
__global__ void __launch_bounds__(256) scale_copy_kernel(const double* __restrict__ x, double* __restrict__ out,
                                                         int64_t n, double alpha, const double* __restrict__ s_dev,
                                                         int inv, int vec) {
  double scale = alpha;
  if (s_dev) {
    const double s = *s_dev;
    scale = inv ? (s != 0.0 ? alpha / s : 0.0) : alpha * s;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n2 = n / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 v = reinterpret_cast<const double2*>(x)[i];
      v.x *= scale;
      v.y *= scale;
      reinterpret_cast<double2*>(out)[i] = v;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = x[n - 1] * scale;
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[i] * scale;
  }
}

int scale_copy(const double* x, double* out, int64_t n, double alpha, const double* s_dev, int inv,
               cudaStream_t stream) {
  TNPY_CHECK_ARG(x && out && n > 0, "bad argument");
  const int vec = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  scale_copy_kernel<<<red_blocks(n), 256, 0, stream>>>(x, out, n, alpha, s_dev, inv, vec);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// y[i] += (alpha * (a_dev ? *a_dev : 1)) * x[i]
__global__ void __launch_bounds__(256) axpy_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n,
                                                   double alpha, const double* __restrict__ a_dev, int vec) {
  const double a = a_dev ? alpha * (*a_dev) : alpha;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n2 = n / 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      const double2 xv = reinterpret_cast<const double2*>(x)[i];
      double2 yv = reinterpret_cast<double2*>(y)[i];
      yv.x = fma(a, xv.x, yv.x);
      yv.y = fma(a, xv.y, yv.y);
      reinterpret_cast<double2*>(y)[i] = yv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) y[n - 1] = fma(a, x[n - 1], y[n - 1]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = fma(a, x[i], y[i]);
  }
}

int axpy(double alpha, const double* a_dev, const double* x, double* y, int64_t n, cudaStream_t stream) {
  TNPY_CHECK_ARG(x && y && n > 0, "bad argument");
  const int vec = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0);
  axpy_kernel<<<red_blocks(n), 256, 0, stream>>>(x, y, n, alpha, a_dev, vec);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// out[i] = sum_j c[j] * V[j][i]   (Ritz vector assembly; c on device)
__global__ void __launch_bounds__(256) combine_kernel(const double* __restrict__ V, int64_t ldv, int m,
                                                      const double* __restrict__ c, int64_t ldc_, int nout,
                                                      double* __restrict__ out, int64_t ldo, int64_t n) {
  // out[o][i] = sum_j c[j * ldc_ + o] * V[j][i]  for o < nout
  extern __shared__ double cs[];
  for (int idx = threadIdx.x; idx < m * nout; idx += blockDim.x) cs[idx] = c[(idx / nout) * ldc_ + (idx % nout)];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    for (int o0 = 0; o0 < nout; o0 += 4) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      for (int j = 0; j < m; ++j) {
        const double v = V[(int64_t)j * ldv + i];
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (o0 + o < nout) acc[o] = fma(cs[j * nout + o0 + o], v, acc[o]);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o)
        if (o0 + o < nout) out[(int64_t)(o0 + o) * ldo + i] = acc[o];
    }
  }
}

int combine(const double* V, int64_t ldv, int m, const double* c, int64_t ldc_, int nout, double* out, int64_t ldo,
            int64_t n, cudaStream_t stream) {
  TNPY_CHECK_ARG(V && c && out && m > 0 && nout > 0 && n > 0, "bad argument");
  const size_t sh = sizeof(double) * m * nout;
  TNPY_CHECK_ARG(sh <= 48 * 1024, "combine: coefficient block too large");
  combine_kernel<<<red_blocks(n), 256, sh, stream>>>(V, ldv, m, c, ldc_, nout, out, ldo, n);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

}  // namespace tnpy

using namespace tnpy;

extern "C" int tnpy_dot(const double* x, const double* y, int64_t n, double* result, void* stream) {
  return multi_dot(x, n, 1, y, n, result, 0, static_cast<cudaStream_t>(stream), nullptr);
}
extern "C" int tnpy_nrm2(const double* x, int64_t n, double* result, void* stream) {
  return multi_dot(x, n, 1, x, n, result, 1, static_cast<cudaStream_t>(stream), nullptr);
}
extern "C" int tnpy_axpy(double alpha, const double* x, double* y, int64_t n, void* stream) {
  return axpy(alpha, nullptr, x, y, n, static_cast<cudaStream_t>(stream));
}
extern "C" int tnpy_axpy_dev(const double* alpha_dev, double alpha_scale, const double* x, double* y, int64_t n,
                             void* stream) {
  TNPY_CHECK_ARG(alpha_dev != nullptr, "null alpha_dev");
  return axpy(alpha_scale, alpha_dev, x, y, n, static_cast<cudaStream_t>(stream));
}
extern "C" int tnpy_scal(double alpha, double* x, int64_t n, void* stream) {
  return scale_copy(x, x, n, alpha, nullptr, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int tnpy_multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h,
                              void* stream) {
  return multi_dot(V, ldv, m, w, n, h, 0, static_cast<cudaStream_t>(stream), nullptr);
}
extern "C" int tnpy_multi_axpy(const double* V, int64_t ldv, int m, const double* h, double* w, int64_t n,
                               void* stream) {
  return multi_axpy(V, ldv, m, h, w, n, nullptr, static_cast<cudaStream_t>(stream), nullptr);
}
