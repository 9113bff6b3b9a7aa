// Orthogonal split of a site tensor without the SVD: A = Q T (tall) or A = T Q (wide), Q with orthonormal
// columns / rows and T square -- the "QR-then-small-SVD" form of the bond split (linalg.py:9-23 is only ever
// called with cutoff = current bond, matrix_product_state.py:206/:218, i.e. as an orthogonalisation), with the
// small SVD of T deferred until somebody asks for the bond spectrum.
//
// Algorithm (rows = the n vectors to orthonormalise, length m >= n): Cholesky-QR applied twice on the
// norm-scaled vectors,
//     G = D^-1 X X^T D^-1 = C C^T,  Q1 = C^-1 D^-1 X,   G2 = Q1 Q1^T = C2 C2^T,  Q = C2^-1 Q1,  T = X Q^T,
// with every big product on the FP64 tensor pipe (gemm_tn) and the n x n Cholesky / triangular inverse as
// blocked DFMA kernels.  Scaling by D makes the factorisation insensitive to the 12 decades the Schmidt
// spectrum of a DMRG wave function spans (only the conditioning of the *normalised* vectors matters, which is
// O(1..100) once a sweep is warm).  The result is verified, not trusted: max|Q Q^T - I| is formed on the
// device and returned; the caller falls back to the Jacobi SVD when it is not at rounding level (cold sweeps
// from a random state can be that ill-conditioned) or when a pivot broke down (reported as +inf).
#include <math.h>

#include "chol.cuh"

namespace tnpy {
namespace {

constexpr int NB = 64;  // Cholesky / inverse block
constexpr int KC = 32;  // K chunk of the tile product

// ---- batched 64x64-tile product on the DFMA pipe:  C = alpha * A * B + beta * C -------------------------
// A(i, k) = A[i * sai + k * sak], B(k, j) = B[k * sbk + j * sbj] (one of each stride pair is 1), C row-major.
// M, N multiples of 64 and K a multiple of 32 (the callers pad).  C may alias A when N == 64 and the C tile is
// the A tile (every load of a CTA precedes its stores).
struct MmArgs {
  const double* A;
  int64_t sai, sak, batchA;
  const double* B;
  int64_t sbk, sbj, batchB;
  double* C;
  int64_t ldc, batchC;
  int K;
  double alpha, beta;
  int lower_only;  // skip tiles strictly above the block diagonal
  int k_mode;      // 0: all of K;  1: B is lower triangular (k >= first column of the tile);  2: A is lower
                   // triangular (k <= last row of the tile)
};

__global__ void __launch_bounds__(256) mm_kernel(MmArgs p) {
  if (p.lower_only && blockIdx.x > blockIdx.y) return;
  __shared__ double As[NB][KC + 1];
  __shared__ double Bs[KC][NB + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const double* A = p.A + (int64_t)blockIdx.z * p.batchA + (int64_t)blockIdx.y * NB * p.sai;
  const double* B = p.B + (int64_t)blockIdx.z * p.batchB + (int64_t)blockIdx.x * NB * p.sbj;
  double* C = p.C + (int64_t)blockIdx.z * p.batchC + (int64_t)blockIdx.y * NB * p.ldc + (int64_t)blockIdx.x * NB;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  const int k_begin = p.k_mode == 1 ? (int)blockIdx.x * NB : 0;
  const int k_end = p.k_mode == 2 ? ((int)blockIdx.y + 1) * NB : p.K;
  for (int k0 = k_begin; k0 < k_end; k0 += KC) {
#pragma unroll
    for (int t = 0; t < (NB * KC) / 256; ++t) {
      const int idx = tid + 256 * t;
      int i, k;
      if (p.sak == 1) {
        k = idx & (KC - 1);
        i = idx / KC;
      } else {
        i = idx & (NB - 1);
        k = idx / NB;
      }
      As[i][k] = A[(int64_t)i * p.sai + (int64_t)(k0 + k) * p.sak];
    }
#pragma unroll
    for (int t = 0; t < (NB * KC) / 256; ++t) {
      const int idx = tid + 256 * t;
      int k, j;
      if (p.sbj == 1) {
        j = idx & (NB - 1);
        k = idx / NB;
      } else {
        k = idx & (KC - 1);
        j = idx / KC;
      }
      Bs[k][j] = B[(int64_t)(k0 + k) * p.sbk + (int64_t)j * p.sbj];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[ty + 16 * i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double* c = C + (int64_t)(ty + 16 * i) * p.ldc + (tx + 16 * j);
      double v = p.alpha * acc[i][j];
      if (p.beta != 0.0) v += p.beta * *c;
      *c = v;
    }
}

int mm(const MmArgs& p, int tiles_n, int tiles_m, int batch, cudaStream_t stream) {
  dim3 grid(tiles_n, tiles_m, batch);
  mm_kernel<<<grid, 256, 0, stream>>>(p);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// ---- one 64x64 diagonal block: G_kk = L L^T in place (upper part zeroed) and Dinv = L^-1 ------------------
// The kernel is a chain of 64 dependent column steps on one CTA, i.e. bound by instruction latency (ncu: 120 k
// warp instructions at 0.19 IPC per scheduler in the shared-memory version), so the tile and the inverse under
// construction live in REGISTERS: thread (ty, tx) owns the 4x4 blocks S[4ty.., 4tx..] and X[4ty.., 4tx..].
// Step j: the owners publish column j of S and row j of X (64 + 64 doubles, double-buffered => one barrier per
// step); every thread then takes the reciprocal square root of the pivot itself and applies two rank-1 updates
// from registers,  S -= l l^T  (right-looking Cholesky) and  X -= l x_j  (forward substitution of L X = I by
// columns of L) -- 32 DFMA and 12 LDS per thread and step, no index arithmetic, no serial tail for the inverse.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* G, int64_t ld, double* __restrict__ Dinv,
                                                        int* __restrict__ fail) {
  __shared__ double colbuf[2][NB];
  __shared__ double rowbuf[2][NB];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  double sr[4][4], xr[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sr[i][k] = G[(int64_t)(4 * ty + i) * ld + 4 * tx + k];
      xr[i][k] = (4 * ty + i == 4 * tx + k) ? 1.0 : 0.0;
    }
#pragma unroll 1
  for (int jb = 0; jb < NB / 4; ++jb) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int j = 4 * jb + kk, buf = kk & 1;
      if (tx == jb) {
#pragma unroll
        for (int i = 0; i < 4; ++i) colbuf[buf][4 * ty + i] = sr[i][kk];
      }
      if (ty == jb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) rowbuf[buf][4 * tx + k] = xr[kk][k];
      }
      __syncthreads();
      double p = colbuf[buf][j];
      const bool bad = !(p > 0.0) || !isfinite(p);  // not positive definite to working precision: flag, stay finite
      if (bad) {
        p = 1.0;
        if (tid == 0) *fail = 1;
      }
      const double ri = rsqrt(p);
      double lr[4], lc[4], xj[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) lr[i] = (4 * ty + i > j) ? colbuf[buf][4 * ty + i] * ri : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lc[k] = (4 * tx + k > j) ? colbuf[buf][4 * tx + k] * ri : 0.0;
        xj[k] = rowbuf[buf][4 * tx + k] * ri;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sr[i][k] = fma(-lr[i], lc[k], sr[i][k]);
          xr[i][k] = fma(-lr[i], xj[k], xr[i][k]);
        }
      if (tx == jb) {  // column j of L is final: below the diagonal the scaled column, on it sqrt(p)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (4 * ty + i > j) sr[i][kk] = lr[i];
          if (4 * ty + i == j) sr[i][kk] = p * ri;
        }
      }
      if (ty == jb) {  // row j of X is final
#pragma unroll
        for (int k = 0; k < 4; ++k) xr[kk][k] = xj[k];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 4 * ty + i, c = 4 * tx + k;
      G[(int64_t)r * ld + c] = r >= c ? sr[i][k] : 0.0;
      Dinv[r * NB + c] = r >= c ? xr[i][k] : 0.0;
    }
}

// dinv[i] = 1 / sqrt(G[i][i]) for i < n (1 on the padding); a zero / non-finite vector raises the fail flag
__global__ void gram_dinv_kernel(const double* __restrict__ G, int64_t ld, int n, int np, double* __restrict__ dinv,
                                 int* __restrict__ fail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  double v = 1.0;
  if (i < n) {
    const double g = G[(int64_t)i * ld + i];
    if (g > 0.0 && isfinite(g))
      v = 1.0 / sqrt(g);
    else
      *fail = 1;
  }
  dinv[i] = v;
}

// G[i][j] *= dinv[i] dinv[j] inside n x n (when dinv != null), + shift on the diagonal; identity on the padding
// up to np x np
__global__ void __launch_bounds__(256) gram_scale_pad_kernel(double* __restrict__ G, int64_t ld, int n, int np,
                                                             const double* __restrict__ dinv, double shift) {
  const int64_t total = (int64_t)np * np;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / np), j = (int)(e % np);
    double* g = G + (int64_t)i * ld + j;
    if (i < n && j < n) {
      double v = *g;
      if (dinv) v = v * dinv[i] * dinv[j];
      if (i == j) v += shift;
      *g = v;
    } else {
      *g = i == j ? 1.0 : 0.0;
    }
  }
}

// Cinv = blockdiag(Dinv_0, Dinv_1, ...), zero elsewhere
__global__ void __launch_bounds__(256) inv_init_kernel(double* __restrict__ Cinv, int np, const double* __restrict__ Dk) {
  const int64_t total = (int64_t)np * np;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / np), j = (int)(e % np);
    Cinv[e] = (i / NB == j / NB) ? Dk[(int64_t)(i / NB) * NB * NB + (i % NB) * NB + (j % NB)] : 0.0;
  }
}

// out[c][r] = in[r][c] * (scale ? scale[c] : 1);  in: rows x cols (ld_in), out: cols x rows (ld_out)
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, int rows, int cols, int64_t ld_in,
                                                        double* __restrict__ out, int64_t ld_out,
                                                        const double* __restrict__ scale) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = in[(int64_t)r * ld_in + c] * (scale ? scale[c] : 1.0);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) out[(int64_t)c * ld_out + r] = tile[tx][i];
  }
}

}  // namespace

int transpose(const double* in, int rows, int cols, int64_t ld_in, double* out, int64_t ld_out, const double* scale,
              cudaStream_t stream) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, out, ld_out, scale);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

namespace {

// out = max |G - I| over n x n as the bit pattern of a non-negative double (ordered like uint64); NaN or a
// raised fail flag gives +inf
__global__ void __launch_bounds__(256) defect_kernel(const double* __restrict__ G, int64_t ld, int n,
                                                     const int* __restrict__ fail, unsigned long long* __restrict__ out) {
  const int64_t total = (int64_t)n * n;
  double mx = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e % n);
    double v = fabs(G[(int64_t)i * ld + j] - (i == j ? 1.0 : 0.0));
    if (!(v == v)) v = INFINITY;
    mx = fmax(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
  if (blockIdx.x == 0 && threadIdx.x == 0 && *fail)
    atomicMax(out, (unsigned long long)__double_as_longlong((double)INFINITY));
}

// native FP64 only (DMMA when the operands are TMA-describable, else the generic DFMA kernel): the operand of
// the C^-1 D^-1 X product mixes 12 decades inside one column, which the int8-sliced tcgen05 path cannot carry
int gemm_fp64(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M, int N, int K,
              cudaStream_t stream) {
  auto describable = [](const double* p, int64_t ld) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && ld % 2 == 0; };
  const bool tiny = (int64_t)M * N < 64 * 64 || K < 16 || M < 32 || N < 32;
  const int algo = (!tiny && describable(A, lda) && describable(B, ldb)) ? TNPY_GEMM_DMMA : TNPY_GEMM_GENERIC;
  return gemm_tn(A, lda, B, ldb, plain_out(C, ldc, M), M, N, K, 0, algo, stream);
}

int stream_grid(int64_t total) {
  int64_t want = (total + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

int chol_padded_dim(int n) {
  int blocks = ceil_div(n, NB), p = 1;
  while (p < blocks) p <<= 1;
  return p * NB;
}

// G (np x np, leading n x n = a symmetric positive definite matrix) -> D^-1 G D^-1 with D = sqrt(diag G), identity
// on the padding; dinv (np) receives 1 / D (1 on the padding).  A non-positive diagonal entry raises *fail.
int spd_scale_pad(double* G, int n, int np, double* dinv, int* fail, cudaStream_t stream) {
  gram_dinv_kernel<<<ceil_div(np, 256), 256, 0, stream>>>(G, np, n, np, dinv, fail);
  TNPY_LAUNCH_OK();
  gram_scale_pad_kernel<<<stream_grid((int64_t)np * np), 256, 0, stream>>>(G, np, n, np, dinv, 0.0);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// G (np x np, lower part = SPD matrix, destroyed: holds L afterwards)  ->  Cinv = L^-1 (np x np, lower)
int cholesky_inverse(double* G, int np, double* Cinv, double* Tmp, double* Dk, int* fail, cudaStream_t stream) {
  const int nblk = np / NB;
  for (int kb = 0; kb < nblk; ++kb) {
    double* diag = G + (int64_t)kb * NB * (np + 1);
    double* dinv_k = Dk + (int64_t)kb * NB * NB;
    chol_diag_kernel<<<1, 256, 0, stream>>>(diag, np, dinv_k, fail);
    TNPY_LAUNCH_OK();
    const int rem = nblk - kb - 1;
    if (rem == 0) break;
    double* panel = diag + (int64_t)NB * np;  // rows below the diagonal block, same columns
    // panel <- panel * L_kk^-T   (B(k, j) = Dinv[j][k])
    MmArgs ps{panel, np, 1, 0, dinv_k, 1, NB, 0, panel, np, 0, NB, 1.0, 0.0, 0, 0};
    TNPY_TRY(mm(ps, 1, rem, 1, stream));
    // trailing(lower) -= panel * panel^T
    MmArgs up{panel, np, 1, 0, panel, 1, np, 0, diag + (int64_t)NB * (np + 1), np, 0, NB, -1.0, 1.0, 1, 0};
    TNPY_TRY(mm(up, rem, rem, 1, stream));
  }
  inv_init_kernel<<<stream_grid((int64_t)np * np), 256, 0, stream>>>(Cinv, np, Dk);
  TNPY_LAUNCH_OK();
  // recursive doubling: inv([[A, 0], [B, C]]) = [[A^-1, 0], [-C^-1 B A^-1, C^-1]]
  for (int s = NB; s < np; s *= 2) {
    const int pairs = np / (2 * s), tiles = s / NB;
    const int64_t bstride = (int64_t)2 * s * (np + 1);
    MmArgs t{G + (int64_t)s * np, np, 1, bstride, Cinv, np, 1, bstride, Tmp + (int64_t)s * np, np, bstride, s, 1.0, 0.0, 0, 1};
    TNPY_TRY(mm(t, tiles, tiles, pairs, stream));
    MmArgs x{Cinv + (int64_t)s * (np + 1), np, 1, bstride, Tmp + (int64_t)s * np, np, 1, bstride,
             Cinv + (int64_t)s * np,       np, bstride, s, -1.0, 0.0, 0, 2};
    TNPY_TRY(mm(x, tiles, tiles, pairs, stream));
  }
  return TNPY_OK;
}

namespace {
int padded_dim(int n) { return chol_padded_dim(n); }
}  // namespace
}  // namespace tnpy

using namespace tnpy;

extern "C" size_t tnpy_qr_split_workspace_bytes(int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const size_t n = rows < cols ? rows : cols, m = rows < cols ? cols : rows;
  const size_t np = padded_dim((int)n);
  return 3 * Workspace::need(n * m) + 4 * Workspace::need(np * np) + Workspace::need(np * NB) + Workspace::need(np) +
         Workspace::need(64, 1) + 512;
}

extern "C" int tnpy_qr_split(const double* A, int rows, int cols, double* Q, double* T, double* defect_dev, int flags,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  TNPY_CHECK_ARG(A && Q && T && defect_dev, "null pointer");
  TNPY_CHECK_ARG(rows > 0 && cols > 0, "non-positive dimension");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  // a square matrix can be split either way: Q first (columns orthonormalised) unless the caller asks for T Q
  const bool tall = rows > cols || (rows == cols && !(flags & TNPY_QR_T_FIRST));
  const int n = tall ? cols : rows, m = tall ? rows : cols;
  const int np = padded_dim(n);
  Workspace ws(workspace, workspace_bytes);
  double* bufA = ws.take<double>((size_t)n * m);
  double* Q1 = ws.take<double>((size_t)n * m);
  double* Q1t = ws.take<double>((size_t)n * m);
  double* G = ws.take<double>((size_t)np * np);
  double* Cinv = ws.take<double>((size_t)np * np);
  double* Tmp = ws.take<double>((size_t)np * np);
  double* Aop = ws.take<double>((size_t)np * np);
  double* Dk = ws.take<double>((size_t)np * NB);
  double* dinv = ws.take<double>((size_t)np);
  int* fail = ws.take<int>(16);
  if (!bufA || !Q1 || !Q1t || !G || !Cinv || !Tmp || !Aop || !Dk || !dinv || !fail) {
    set_error("tnpy_qr_split: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              tnpy_qr_split_workspace_bytes(rows, cols));
    return TNPY_EWORKSPACE;
  }
  TNPY_CUDA_OK(cudaMemsetAsync(fail, 0, 16 * sizeof(int), stream));
  TNPY_CUDA_OK(cudaMemsetAsync(defect_dev, 0, sizeof(double), stream));

  // X: n x m (rows = the vectors), Xt: m x n.  The caller's matrix is one of them, the other is a transpose.
  const double *X, *Xt;
  if (tall) {
    Xt = A;
    TNPY_TRY(transpose(A, m, n, n, bufA, m, nullptr, stream));
    X = bufA;
  } else {
    X = A;
    TNPY_TRY(transpose(A, n, m, m, bufA, n, nullptr, stream));
    Xt = bufA;
  }
  const int sgrid = stream_grid((int64_t)np * np);

  // Cholesky-QR passes.  State entering a pass: the current vectors as rows (cur, n x m) and transposed (curT,
  // m x n).  The Gram matrix is formed from curT, whose buffer then receives the new rows; their transpose goes
  // to the other buffer.  Pass 0 works on the norm-scaled vectors (and, in the shifted variant, on G + sigma I:
  // always factorisable, leaves cond(Q) ~ sqrt(sigma) cond(X) for the next two passes to remove).
  const int passes = (flags & TNPY_QR_SHIFTED) ? 3 : 2;
  const double sigma = (flags & TNPY_QR_SHIFTED) ? 100.0 * 1.1102230246251565e-16 * (double)n : 0.0;
  const double *cur = X, *curT = Xt;
  double *rows_buf = Q1, *trans_buf = Q1t;  // where this pass puts its rows / their transpose
  for (int pass = 0; pass < passes; ++pass) {
    const bool last = pass == passes - 1;
    TNPY_TRY(gemm_fp64(curT, n, curT, n, G, np, n, n, m, stream));
    if (pass == 0) {
      gram_dinv_kernel<<<ceil_div(np, 256), 256, 0, stream>>>(G, np, n, np, dinv, fail);
      TNPY_LAUNCH_OK();
    }
    gram_scale_pad_kernel<<<sgrid, 256, 0, stream>>>(G, np, n, np, pass == 0 ? dinv : nullptr, pass == 0 ? sigma : 0.0);
    TNPY_LAUNCH_OK();
    TNPY_TRY(cholesky_inverse(G, np, Cinv, Tmp, Dk, fail, stream));
    // Aop[k][i] = Cinv[i][k] (* dinv[k] in pass 0): the TN operand of  next = Cinv (D^-1) cur
    TNPY_TRY(transpose(Cinv, np, np, np, Aop, np, pass == 0 ? dinv : nullptr, stream));
    double* next = (last && !tall) ? Q : rows_buf;   // the last pass lands in the caller's Q (wide: rows)
    TNPY_TRY(gemm_fp64(Aop, np, cur, m, next, m, n, m, n, stream));
    double* nextT = (last && tall) ? Q : trans_buf;  // (tall: transposed)
    TNPY_TRY(transpose(next, n, m, m, nextT, n, nullptr, stream));
    cur = next;
    curT = nextT;
    // the following pass forms its Gram matrix from curT and may then overwrite that buffer with its rows; its
    // transpose goes over this pass's rows, which its GEMM has consumed by then
    double* swap = rows_buf;
    rows_buf = trans_buf;
    trans_buf = swap;
  }
  const double* Qt = curT;  // m x n

  // verification: max |Q Q^T - I|
  TNPY_TRY(gemm_fp64(Qt, n, Qt, n, G, np, n, n, m, stream));
  defect_kernel<<<stream_grid((int64_t)n * n), 256, 0, stream>>>(G, np, n, fail,
                                                                reinterpret_cast<unsigned long long*>(defect_dev));
  TNPY_LAUNCH_OK();

  // T: tall  A = Q T  => T = Q^T A  = sum_c Qt[c][i] Xt[c][j];   wide  A = T Q => T = A Q^T = sum_c Xt[c][i] Qt[c][j]
  if (tall)
    TNPY_TRY(gemm_fp64(Qt, n, Xt, n, T, n, n, n, m, stream));
  else
    TNPY_TRY(gemm_fp64(Xt, n, Qt, n, T, n, n, n, m, stream));
  return TNPY_OK;
}
