// On-device bond SVD (reference linalg.py:9-23 -> numpy.linalg.svd / LAPACK gesdd, called from
// matrix_product_state.py:206, :218), the dense lowest-eigenpair solve for tiny sites (reference
// linalg.py:42-61) and the neighbour absorb (reference matrix_product_state.py:207-223).
//
// SVD = one-sided Jacobi (Hestenes) on the SHORT side.  The matrix is held as G (n x m, n <= m, rows
// are the vectors being orthogonalised, contiguous => coalesced); row rotations are accumulated into
// P (n x n, starts as I), so that   G_in = P^T * G_out   with G_out rows mutually orthogonal:
//     s_k = ||G_out[k]||,   W[k] = G_out[k] / s_k,   G_in = (P^T) diag(s) W.
// Singular values are sorted descending with ties broken by row index (deterministic).
//   wide input  A (rows <  cols): G = A,    U = P^T, Vt = W.
//   tall input  A (rows >= cols): G = A^T,  U = W^T, Vt = P.
// Two execution paths: a single-CTA kernel with G and P resident in shared memory (edge bonds,
// small chi: no launch or sync overhead), and a per-round multi-CTA kernel (one CTA per row pair,
// rows staged in shared memory, L2-resident matrix) for large bonds.
#include <math.h>

#include "common.cuh"

namespace tnpy {

int multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h, int mode,
              cudaStream_t stream);

__device__ __forceinline__ void rr_pair(int round, int i, int me, int& p, int& q) {
  // round-robin tournament over `me` (even) players: player me-1 fixed, the rest rotate
  p = (i == 0) ? me - 1 : (round + i) % (me - 1);
  q = (round + me - 1 - i) % (me - 1);
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
}

// Hestenes rotation from the 2x2 Gram entries; returns false when the pair is already orthogonal.
__device__ __forceinline__ bool hestenes_cs(double a, double b, double c, double tol, double& cs, double& sn) {
  if (!(fabs(c) > tol * sqrt(a * b)) || c == 0.0) return false;
  const double zeta = (b - a) / (2.0 * c);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  cs = 1.0 / sqrt(1.0 + t * t);
  sn = cs * t;
  return true;
}

// ---------------------------------------------------------------------------------------------
// small path: everything in shared memory, one CTA, warp per pair
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) hestenes_small_kernel(double* __restrict__ G, int n, int m, int64_t ldg,
                                                             double* __restrict__ P, double tol, int max_sweeps,
                                                             int* __restrict__ info) {
  extern __shared__ double sm[];
  double* gs = sm;                      // n x m
  double* ps = sm + (size_t)n * m;      // n x n
  __shared__ int rot;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
  for (int idx = tid; idx < n * m; idx += nt) gs[idx] = G[(int64_t)(idx / m) * ldg + idx % m];
  for (int idx = tid; idx < n * n; idx += nt) ps[idx] = (idx / n == idx % n) ? 1.0 : 0.0;
  __syncthreads();
  const int me = (n + 1) & ~1, half = me / 2;
  int sweep = 0;
  for (; sweep < max_sweeps && n > 1; ++sweep) {
    if (tid == 0) rot = 0;
    __syncthreads();
    for (int round = 0; round < me - 1; ++round) {
      for (int i = warp; i < half; i += nwarps) {
        int p, q;
        rr_pair(round, i, me, p, q);
        if (q >= n) continue;
        double* gp = gs + (size_t)p * m;
        double* gq = gs + (size_t)q * m;
        double a = 0.0, b = 0.0, c = 0.0;
        for (int k = lane; k < m; k += 32) {
          const double x = gp[k], y = gq[k];
          a = fma(x, x, a);
          b = fma(y, y, b);
          c = fma(x, y, c);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        c = warp_sum(c);
        double cs, sn;
        if (!hestenes_cs(a, b, c, tol, cs, sn)) continue;
        for (int k = lane; k < m; k += 32) {
          const double x = gp[k], y = gq[k];
          gp[k] = cs * x - sn * y;
          gq[k] = sn * x + cs * y;
        }
        double* pp = ps + (size_t)p * n;
        double* pq = ps + (size_t)q * n;
        for (int k = lane; k < n; k += 32) {
          const double x = pp[k], y = pq[k];
          pp[k] = cs * x - sn * y;
          pq[k] = sn * x + cs * y;
        }
        if (lane == 0) atomicAdd(&rot, 1);
      }
      __syncthreads();
    }
    const int r = rot;
    __syncthreads();
    if (r == 0) break;
  }
  for (int idx = tid; idx < n * m; idx += nt) G[(int64_t)(idx / m) * ldg + idx % m] = gs[idx];
  for (int idx = tid; idx < n * n; idx += nt) P[idx] = ps[idx];
  if (tid == 0 && info) info[0] = sweep;
}

// ---------------------------------------------------------------------------------------------
// large path: one launch per round, one CTA per pair
// ---------------------------------------------------------------------------------------------
__global__ void identity_kernel(double* __restrict__ P, int n) {
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    P[e] = (e / n == e % n) ? 1.0 : 0.0;
}

template <bool STAGE>
__global__ void __launch_bounds__(256) hestenes_round_kernel(double* __restrict__ G, int n, int m, int64_t ldg,
                                                             double* __restrict__ P, int round, int me, double tol,
                                                             unsigned int* __restrict__ rot_count) {
  extern __shared__ double rows[];  // STAGE: 2 x m doubles
  __shared__ double sh[32];
  __shared__ double rot_cs[2];
  __shared__ int do_rot;
  int p, q;
  rr_pair(round, blockIdx.x, me, p, q);
  if (q >= n) return;
  double* gp = G + (int64_t)p * ldg;
  double* gq = G + (int64_t)q * ldg;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const double x = gp[k], y = gq[k];
    if (STAGE) {
      rows[k] = x;
      rows[m + k] = y;
    }
    a = fma(x, x, a);
    b = fma(y, y, b);
    c = fma(x, y, c);
  }
  a = block_sum(a, sh);
  b = block_sum(b, sh);
  c = block_sum(c, sh);
  if (threadIdx.x == 0) {
    double cs = 1.0, sn = 0.0;
    do_rot = hestenes_cs(a, b, c, tol, cs, sn) ? 1 : 0;
    rot_cs[0] = cs;
    rot_cs[1] = sn;
    if (do_rot) atomicAdd(rot_count, 1u);
  }
  __syncthreads();
  if (!do_rot) return;
  const double cs = rot_cs[0], sn = rot_cs[1];
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const double x = STAGE ? rows[k] : gp[k];
    const double y = STAGE ? rows[m + k] : gq[k];
    gp[k] = cs * x - sn * y;
    gq[k] = sn * x + cs * y;
  }
  double* pp = P + (int64_t)p * n;
  double* pq = P + (int64_t)q * n;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double x = pp[k], y = pq[k];
    pp[k] = cs * x - sn * y;
    pq[k] = sn * x + cs * y;
  }
}

// ---------------------------------------------------------------------------------------------
// finalize: norms, deterministic descending rank, scatter into U / s / Vt
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_norm_kernel(const double* __restrict__ G, int n, int m, int64_t ldg,
                                                       double* __restrict__ norms) {
  __shared__ double sh[32];
  const int k = blockIdx.x;
  double a = 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double x = G[(int64_t)k * ldg + i];
    a = fma(x, x, a);
  }
  a = block_sum(a, sh);
  if (threadIdx.x == 0) norms[k] = sqrt(a);
}

__global__ void rank_kernel(const double* __restrict__ norms, int n, int* __restrict__ rank, double* __restrict__ s) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const double v = norms[k];
  int r = 0;
  for (int j = 0; j < n; ++j) {
    const double u = norms[j];
    r += (u > v) || (u == v && j < k);
  }
  rank[k] = r;
  s[r] = v;
}

// w_out[rank[k]*wa + i*wb] = G[k][i] / s_k  (i < m);   p_out[rank[k]*pa + j*pb] = P[k][j]  (j < n)
__global__ void __launch_bounds__(256) scatter_kernel(const double* __restrict__ G, int n, int m, int64_t ldg,
                                                      const double* __restrict__ P, const double* __restrict__ norms,
                                                      const int* __restrict__ rank, double* __restrict__ w_out,
                                                      int64_t wa, int64_t wb, double* __restrict__ p_out, int64_t pa,
                                                      int64_t pb) {
  const int k = blockIdx.x;
  const int r = rank[k];
  const double nk = norms[k];
  const double inv = nk > 0.0 ? 1.0 / nk : 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) w_out[r * wa + i * wb] = G[(int64_t)k * ldg + i] * inv;
  for (int j = threadIdx.x; j < n; j += blockDim.x) p_out[r * pa + j * pb] = P[(int64_t)k * n + j];
}

__global__ void __launch_bounds__(256) transpose2d_kernel(const double* __restrict__ in, int rows, int cols,
                                                          double* __restrict__ out, const double* __restrict__ colscale) {
  // out[c][r] = in[r][c] * (colscale ? colscale[r] : 1)
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = in[(int64_t)r * cols + c] * (colscale ? colscale[r] : 1.0);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) out[(int64_t)c * rows + r] = tile[tx][i];
  }
}

static int transpose2d(const double* in, int rows, int cols, double* out, const double* rowscale, cudaStream_t stream) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose2d_kernel<<<grid, 256, 0, stream>>>(in, rows, cols, out, rowscale);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// out[i][k] = in[i][k] * s[k]
__global__ void colscale_kernel(const double* __restrict__ in, const double* __restrict__ s, double* __restrict__ out,
                                int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = in[e] * s[e % cols];
}

struct PinnedWord {
  unsigned int* host = nullptr;
  PinnedWord() { cudaMallocHost(&host, 64); }
};

// Orthogonalise the rows of G (n x m, n <= m) in place; P (n x n) receives the accumulated rotations.
static int hestenes(double* G, int n, int m, int64_t ldg, double* P, unsigned int* counter_dev, cudaStream_t stream,
                    int* sweeps_out) {
  const double tol = fmax(1e-15, 2.220446049250313e-16 * sqrt((double)m));
  const int max_sweeps = 60;
  const size_t small_bytes = sizeof(double) * ((size_t)n * m + (size_t)n * n);
  if (small_bytes <= 200 * 1024) {
    static bool configured = false;
    if (!configured) {
      TNPY_CUDA_OK(cudaFuncSetAttribute(hestenes_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      configured = true;
    }
    const int threads = n >= 32 ? 512 : (n >= 8 ? 256 : 64);
    hestenes_small_kernel<<<1, threads, small_bytes, stream>>>(G, n, m, ldg, P, tol, max_sweeps, nullptr);
    TNPY_LAUNCH_OK();
    if (sweeps_out) *sweeps_out = -1;
    return TNPY_OK;
  }
  static PinnedWord pinned;
  if (!pinned.host) {
    set_error("hestenes: pinned allocation failed");
    return TNPY_ECUDA;
  }
  identity_kernel<<<sm_count() * 4, 256, 0, stream>>>(P, n);
  TNPY_LAUNCH_OK();
  const int me = (n + 1) & ~1, half = me / 2;
  const size_t stage_bytes = sizeof(double) * 2 * (size_t)m;
  const bool stage = stage_bytes <= 96 * 1024;
  static bool configured2 = false;
  if (stage && !configured2) {
    TNPY_CUDA_OK(cudaFuncSetAttribute(hestenes_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured2 = true;
  }
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    TNPY_CUDA_OK(cudaMemsetAsync(counter_dev, 0, sizeof(unsigned int), stream));
    for (int round = 0; round < me - 1; ++round) {
      if (stage)
        hestenes_round_kernel<true><<<half, 256, stage_bytes, stream>>>(G, n, m, ldg, P, round, me, tol, counter_dev);
      else
        hestenes_round_kernel<false><<<half, 256, 0, stream>>>(G, n, m, ldg, P, round, me, tol, counter_dev);
    }
    TNPY_LAUNCH_OK();
    count_launch(me - 2);
    TNPY_CUDA_OK(cudaMemcpyAsync(pinned.host, counter_dev, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    TNPY_CUDA_OK(cudaStreamSynchronize(stream));
    if (*pinned.host == 0u) break;
  }
  if (sweeps_out) *sweeps_out = sweep;
  return TNPY_OK;
}

// lowest eigenpair from the sorted SVD of the shifted matrix: eval = s[n-1] - shift, evec = Vt[n-1]
__global__ void eigh_pick_kernel(const double* __restrict__ s, const double* __restrict__ Vt, int n,
                                 const double* __restrict__ shift, double* __restrict__ eval, double* __restrict__ evec) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *eval = s[n - 1] - *shift;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    evec[i] = Vt[(int64_t)(n - 1) * n + i];
}

__global__ void add_diag_kernel(double* __restrict__ H, int n, const double* __restrict__ shift) {
  const double s = *shift;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) H[(int64_t)i * n + i] += s;
}

}  // namespace tnpy

using namespace tnpy;

extern "C" size_t tnpy_svd_workspace_bytes(int rows, int cols) {
  const size_t n = rows < cols ? rows : cols, m = rows < cols ? cols : rows;
  return Workspace::need(n * m) + Workspace::need(n * n) + Workspace::need(n) + Workspace::need(n, sizeof(int)) + 1024;
}

extern "C" int tnpy_svd(double* A, int rows, int cols, double* U, double* s, double* Vt, void* workspace,
                        size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(A && U && s && Vt && rows > 0 && cols > 0, "bad argument");
  const bool tall = rows >= cols;
  const int n = tall ? cols : rows, m = tall ? rows : cols;
  Workspace ws(workspace, workspace_bytes);
  double* Gt = ws.take<double>((size_t)n * m);
  double* P = ws.take<double>((size_t)n * n);
  double* norms = ws.take<double>(n);
  int* rank = ws.take<int>(n);
  unsigned int* counter = ws.take<unsigned int>(64);
  if (!Gt || !P || !norms || !rank || !counter) {
    set_error("tnpy_svd: workspace too small");
    return TNPY_EWORKSPACE;
  }
  double* G = A;
  if (tall) {
    TNPY_TRY(transpose2d(A, rows, cols, Gt, nullptr, stream));
    G = Gt;
  }
  TNPY_TRY(hestenes(G, n, m, m, P, counter, stream, nullptr));
  row_norm_kernel<<<n, 256, 0, stream>>>(G, n, m, m, norms);
  TNPY_LAUNCH_OK();
  rank_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(norms, n, rank, s);
  TNPY_LAUNCH_OK();
  if (tall)  // U = W^T (rows x n): U[i][r] ; Vt = P (n x cols): Vt[r][j]
    scatter_kernel<<<n, 256, 0, stream>>>(G, n, m, m, P, norms, rank, U, 1, n, Vt, cols, 1);
  else       // Vt = W (n x cols): Vt[r][i] ; U = P^T (rows x n): U[j][r]
    scatter_kernel<<<n, 256, 0, stream>>>(G, n, m, m, P, norms, rank, Vt, cols, 1, U, 1, n);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

extern "C" size_t tnpy_eigh_workspace_bytes(int n) {
  return tnpy_svd_workspace_bytes(n, n) + 2 * Workspace::need((size_t)n * n) + Workspace::need(n) + 1024;
}

extern "C" int tnpy_eigh_lowest(double* H, int n, double* eval_dev, double* evec, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(H && eval_dev && evec && n > 0, "bad argument");
  Workspace ws(workspace, workspace_bytes);
  double* U = ws.take<double>((size_t)n * n);
  double* Vt = ws.take<double>((size_t)n * n);
  double* s = ws.take<double>(n);
  double* shift = ws.take<double>(8);
  if (!U || !Vt || !s || !shift) {
    set_error("tnpy_eigh_lowest: workspace too small");
    return TNPY_EWORKSPACE;
  }
  // shift = ||H||_F >= max |lambda|  =>  H + shift*I is positive semi-definite and its SVD is its
  // eigendecomposition; the lowest eigenvalue is the smallest singular value minus the shift.
  TNPY_TRY(multi_dot(H, (int64_t)n * n, 1, H, (int64_t)n * n, shift, 1, stream));
  add_diag_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(H, n, shift);
  TNPY_LAUNCH_OK();
  TNPY_TRY(tnpy_svd(H, n, n, U, s, Vt, static_cast<char*>(workspace) + ws.used, workspace_bytes - ws.used, stream_));
  eigh_pick_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(s, Vt, n, shift, eval_dev, evec);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

extern "C" size_t tnpy_absorb_workspace_bytes(int k, int n, int nb) {
  // right: n x k scaled transpose; left: (n x nb) transposed neighbour + n x k scaled copy
  const size_t big = (size_t)(k > nb ? k : nb);
  return Workspace::need(big * n) + Workspace::need((size_t)k * n) + 1024;
}

extern "C" int tnpy_absorb_right(const double* s, const double* Vt, int k, int n, const double* Nb, int cols_nb,
                                 double* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(s && Vt && Nb && out && k > 0 && n > 0 && cols_nb > 0, "bad argument");
  Workspace ws(workspace, workspace_bytes);
  double* svt_t = ws.take<double>((size_t)n * k);  // [i][k] = s[k] * Vt[k][i]
  if (!svt_t) {
    set_error("tnpy_absorb_right: workspace too small");
    return TNPY_EWORKSPACE;
  }
  TNPY_TRY(transpose2d(Vt, k, n, svt_t, s, stream));
  // out[k][j] = sum_i svt_t[i][k] * Nb[i][j]
  return gemm_tn(svt_t, k, Nb, cols_nb, plain_out(out, cols_nb, k), k, cols_nb, n, 0, TNPY_GEMM_AUTO, stream);
}

extern "C" int tnpy_absorb_left(const double* U, const double* s, int n, int k, const double* Nb, int rows_nb,
                                double* out, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(U && s && Nb && out && k > 0 && n > 0 && rows_nb > 0, "bad argument");
  Workspace ws(workspace, workspace_bytes);
  double* nb_t = ws.take<double>((size_t)rows_nb * n);  // [j][i] = Nb[i][j]
  double* us = ws.take<double>((size_t)n * k);          // [j][k] = U[j][k] * s[k]
  if (!nb_t || !us) {
    set_error("tnpy_absorb_left: workspace too small");
    return TNPY_EWORKSPACE;
  }
  TNPY_TRY(transpose2d(Nb, rows_nb, n, nb_t, nullptr, stream));
  colscale_kernel<<<sm_count() * 4, 256, 0, stream>>>(U, s, us, n, k);
  TNPY_LAUNCH_OK();
  // out[i][k] = sum_j nb_t[j][i] * us[j][k]
  return gemm_tn(nb_t, rows_nb, us, k, plain_out(out, k, rows_nb), rows_nb, k, n, 0, TNPY_GEMM_AUTO, stream);
}
