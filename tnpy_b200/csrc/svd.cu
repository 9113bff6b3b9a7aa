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
// small chi: no launch or sync overhead), and block Jacobi on the FP64 tensor pipe for large bonds
// (32-row blocks paired round-robin; per round a Gram kernel, a 64x64 rotation kernel and an apply
// kernel -- see the comment above jb_gram_kernel).
#include <math.h>

#include "common.cuh"

namespace tnpy {

int multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h, int mode,
              cudaStream_t stream, const int* skip = nullptr);

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
// large path: block one-sided Jacobi.  Rows are grouped in blocks of kJB; a round pairs the blocks
// round-robin and, for every pair (a 64-row panel), runs three kernels:
//   jb_gram    partial Gram matrices of the panel over column chunks of G      (DMMA, all SMs)
//   jb_rotate  fixed-order reduction of the partials, one cyclic Jacobi sweep on the 64x64 Gram
//              in shared memory -> the accumulated 64x64 rotation J^T          (one CTA per pair)
//   jb_apply   panel <- J^T panel over column chunks of [G | P], in place      (DMMA, all SMs)
// The matrix is touched 3x per round instead of once per row pair: (n/kJB - 1) rounds per sweep
// instead of (n - 1), and both heavy kernels run on the FP64 tensor pipe.  Rotations are computed
// from freshly formed Gram entries at every visit, so the iteration is self-correcting and the
// singular values keep Hestenes' relative accuracy (the final norms come from the rows themselves).
// ---------------------------------------------------------------------------------------------
constexpr int kJB = 32;             // rows per block
constexpr int kJP = 2 * kJB;        // rows per panel (block pair)
constexpr int kJCH = 128;           // columns per shared-memory chunk
constexpr int kJST = kJCH + 4;      // panel row stride in shared memory (conflict-free DMMA fragments)
constexpr int kJTS = kJP + 4;       // rotation-matrix row stride in shared memory
constexpr int kJGramCols = 256;     // columns of G per jb_gram CTA
constexpr int kJApplyCols = 256;    // columns of [G | P] per jb_apply CTA (kJCH at a time)
// Vectors below theta * max norm would be orthogonalised among themselves first (cluster phase).
// Measured on DMRG wave functions (scripts/svd_trace.py): the cluster phase converges, but the full
// phase that follows still needs as many sweeps as without it, so the schedule is disabled (theta = 0
// => single phase).  The norm-sorted placement of the vectors is kept.
constexpr double kClusterTheta = 0.0;

__device__ __forceinline__ void dmma884_(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// GP = [G | P] (n_pad x ld): G from A (transposed when `tall`), P = identity, padding rows zero.
// Row k of GP holds source vector src[k] (vectors sorted by decreasing norm: src = inverse of rank),
// and P starts as the matching permutation matrix so that G_in = P^T G_out still holds.
__global__ void __launch_bounds__(256) jb_init_kernel(const double* __restrict__ A, int rows, int cols, int tall,
                                                      const int* __restrict__ src, double* __restrict__ GP, int n,
                                                      int m, int n_pad, int64_t ld) {
  const int64_t total = (int64_t)n_pad * ld;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(e / ld), c = (int)(e % ld);
    double v = 0.0;
    if (k < n) {
      const int o = src[k];
      if (c < m) v = tall ? A[(int64_t)c * cols + o] : A[(int64_t)o * cols + c];
      else if (c - m == o) v = 1.0;
    }
    GP[e] = v;
  }
}

// norms of the vectors of A that the SVD orthogonalises: columns when `tall`, rows otherwise
__global__ void __launch_bounds__(256) vec_norm_kernel(const double* __restrict__ A, int rows, int cols, int tall,
                                                       double* __restrict__ norms) {
  if (tall) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cols) return;
    double a = 0.0;
    for (int i = 0; i < rows; ++i) {
      const double x = A[(int64_t)i * cols + k];
      a = fma(x, x, a);
    }
    norms[k] = sqrt(a);
  } else {
    __shared__ double sh[32];
    for (int k = blockIdx.x; k < rows; k += gridDim.x) {
      double a = 0.0;
      for (int i = threadIdx.x; i < cols; i += blockDim.x) {
        const double x = A[(int64_t)k * cols + i];
        a = fma(x, x, a);
      }
      a = block_sum(a, sh);
      if (threadIdx.x == 0) norms[k] = sqrt(a);
    }
  }
}

// src[rank[k]] = k;  n_big = number of vectors with norm >= theta * max norm (written to info[0])
__global__ void invert_rank_kernel(const int* __restrict__ rank, const double* __restrict__ sorted, int n, double theta,
                                   int* __restrict__ src, int* __restrict__ info) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    src[rank[k]] = k;
    const double cut = theta * sorted[0];
    if (sorted[k] >= cut && (k == n - 1 || sorted[k + 1] < cut)) info[0] = k + 1;
    if (k == 0 && !(sorted[0] >= cut)) info[0] = 0;
  }
}

// Which blocks take part and which pairs matter in the current phase of the two-phase schedule:
//   phase 1: all nb blocks, but panels made of two "small" blocks (index >= nbig) are skipped and only
//            pairs touching a big block count -- the well-separated part converges in a few sweeps;
//   phase 2: only the blocks [b0, b0 + nb) (the cluster of tiny, noise-dominated vectors), every pair.
struct JbPhase {
  int b0;    // first block of the round-robin
  int nb;    // number of blocks in the round-robin (even)
  int nbig;  // blocks with absolute index >= nbig are "small"
};

// block pair of CTA `i` in `round` (round -1: adjacent blocks); false when the panel is skipped
__device__ __forceinline__ bool block_pair(int round, int i, const JbPhase& ph, int& I, int& J) {
  if (round < 0) {
    I = 2 * i;
    J = 2 * i + 1;
  } else {
    rr_pair(round, i, ph.nb, I, J);
  }
  I += ph.b0;
  J += ph.b0;
  return !(I >= ph.nbig && J >= ph.nbig);
}

__device__ __forceinline__ int panel_row(int I, int J, int r) { return r < kJB ? I * kJB + r : J * kJB + (r - kJB); }

// Load a kJP x kJCH chunk of the panel (columns [c0, c0 + kJCH) clipped to c_end) into shared memory.
__device__ __forceinline__ void load_panel(const double* __restrict__ GP, int64_t ld, int I, int J, int c0, int c_end,
                                           double* __restrict__ panel) {
  for (int idx = threadIdx.x; idx < kJP * (kJCH / 2); idx += blockDim.x) {
    const int r = idx / (kJCH / 2), c = (idx % (kJCH / 2)) * 2;
    const double* src = GP + (int64_t)panel_row(I, J, r) * ld + c0 + c;
    double2 v = make_double2(0.0, 0.0);
    if (c0 + c + 1 < c_end) v = *reinterpret_cast<const double2*>(src);
    else if (c0 + c < c_end) v.x = src[0];
    panel[r * kJST + c] = v.x;
    panel[r * kJST + c + 1] = v.y;
  }
}

__global__ void __launch_bounds__(256) jb_gram_kernel(const double* __restrict__ GP, int64_t ld, int m, JbPhase ph,
                                                      int round, int gchunks, double* __restrict__ partial) {
  extern __shared__ double sm[];
  double* panel = sm;
  int I, J;
  if (!block_pair(round, blockIdx.x, ph, I, J)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c_begin = blockIdx.y * kJGramCols, c_end = min(m, c_begin + kJGramCols);
  double acc[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
  for (int c0 = c_begin; c0 < c_end; c0 += kJCH) {
    load_panel(GP, ld, I, J, c0, c_end, panel);
    __syncthreads();
#pragma unroll 4
    for (int k0 = 0; k0 < kJCH; k0 += 4) {
      const double a = panel[(8 * warp + g) * kJST + k0 + t];
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma884_(acc[j][0], acc[j][1], a, panel[(8 * j + g) * kJST + k0 + t]);
    }
    __syncthreads();
  }
  double* out = partial + ((int64_t)blockIdx.x * gchunks + blockIdx.y) * (kJP * kJP);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<double2*>(out + (8 * warp + g) * kJP + 8 * j + 2 * t) = make_double2(acc[j][0], acc[j][1]);
}

// cross_only: rotate only (row of block I, row of block J) pairs -- 32 inner rounds instead of 63; the
// first round of every sweep pairs blocks (0,1),(2,3),... with the full 64x64 sweep so every block is
// also orthogonalised internally once per sweep.
__global__ void __launch_bounds__(512) jb_rotate_kernel(const double* __restrict__ partial, int gchunks, double tol,
                                                        int cross_only, JbPhase ph, int round,
                                                        double* __restrict__ Jt,
                                                        unsigned int* __restrict__ rot_count,
                                                        int* __restrict__ skip) {
  int I, J;
  if (!block_pair(round, blockIdx.x, ph, I, J)) {
    if (threadIdx.x == 0) skip[blockIdx.x] = 1;
    return;
  }
  const bool i_big = I < ph.nbig, j_big = J < ph.nbig;
  extern __shared__ double sm[];
  double(*a)[kJP + 1] = reinterpret_cast<double(*)[kJP + 1]>(sm);
  double(*z)[kJP + 1] = reinterpret_cast<double(*)[kJP + 1]>(sm + kJP * (kJP + 1));
  __shared__ double cs[kJB], sn[kJB];
  __shared__ int pp[kJB], qq[kJB];
  __shared__ int significant, significant_all;
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* src = partial + (int64_t)blockIdx.x * gchunks * (kJP * kJP);
  if (tid == 0) significant = significant_all = 0;
  for (int idx = tid; idx < kJP * kJP; idx += nt) {
    double v = 0.0;
    for (int c = 0; c < gchunks; ++c) v += src[(int64_t)c * (kJP * kJP) + idx];  // fixed order: deterministic
    a[idx / kJP][idx % kJP] = v;
    z[idx / kJP][idx % kJP] = (idx / kJP == idx % kJP) ? 1.0 : 0.0;
  }
  __syncthreads();
  // convergence bookkeeping on the freshly formed Gram matrix
  int mine = 0, mine_all = 0;
  for (int idx = tid; idx < kJP * kJP; idx += nt) {
    const int p = idx / kJP, q = idx % kJP;
    if (p < q && fabs(a[p][q]) > tol * sqrt(a[p][p] * a[q][q])) {
      const bool relevant = (p < kJB ? i_big : j_big) || (q < kJB ? i_big : j_big);
      if (relevant) ++mine_all;  // drives the sweep loop of the current phase
      if (!cross_only || (p < kJB && q >= kJB)) ++mine;  // pairs this round is going to rotate (p < kJB <= q)
    }
  }
  if (mine) atomicAdd(&significant, mine);
  if (mine_all) atomicAdd(&significant_all, mine_all);
  __syncthreads();
  const int nsig = significant;
  if (tid == 0) {
    skip[blockIdx.x] = (nsig == 0);
    if (significant_all) atomicAdd(rot_count, (unsigned int)significant_all);  // drives the sweep loop
  }
  if (nsig == 0) return;  // nothing to rotate here: jb_apply skips this panel
  constexpr int half = kJP / 2;
  {
    // One inner round = `half` disjoint rotations J = prod_i R_i.  a <- J^T a J is applied in a single pass:
    // thread (i, j) owns the 2x2 block a[{p_i, q_i}][{p_j, q_j}], rotates its columns with R_j and its rows with
    // R_i, and nobody else touches those four entries -- so a round costs two barriers (parameters, update)
    // instead of three, and the parameter chain is div -> sqrt -> div -> rsqrt.
    const int inner_rounds = cross_only ? kJB : kJP - 1;
    for (int round = 0; round < inner_rounds; ++round) {
      if (tid < half) {
        int p, q;
        if (cross_only) {
          p = tid;
          q = kJB + ((tid + round) & (kJB - 1));
        } else {
          rr_pair(round, tid, kJP, p, q);
        }
        double c = 1.0, s = 0.0;
        const double apq = a[p][q], app = a[p][p], aqq = a[q][q];
        if (fabs(apq) > tol * sqrt(app * aqq) && apq != 0.0) {
          const double tau = (aqq - app) / (2.0 * apq);
          const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = rsqrt(1.0 + tt * tt);
          s = tt * c;
        }
        cs[tid] = c;
        sn[tid] = s;
        pp[tid] = p;
        qq[tid] = q;
      }
      __syncthreads();
      for (int idx = tid; idx < half * half; idx += nt) {
        const int i = idx / half, j = idx % half;
        const double ci = cs[i], si = sn[i], cj = cs[j], sj = sn[j];
        if (si == 0.0 && sj == 0.0) continue;
        const int pi = pp[i], qi = qq[i], pj = pp[j], qj = qq[j];
        const double m00 = a[pi][pj], m01 = a[pi][qj], m10 = a[qi][pj], m11 = a[qi][qj];
        // columns: (col p_j, col q_j) <- (c_j col p_j - s_j col q_j, s_j col p_j + c_j col q_j)
        const double t00 = cj * m00 - sj * m01, t01 = sj * m00 + cj * m01;
        const double t10 = cj * m10 - sj * m11, t11 = sj * m10 + cj * m11;
        // rows: (row p_i, row q_i) <- (c_i row p_i - s_i row q_i, s_i row p_i + c_i row q_i)
        a[pi][pj] = ci * t00 - si * t10;
        a[pi][qj] = ci * t01 - si * t11;
        a[qi][pj] = si * t00 + ci * t10;
        a[qi][qj] = si * t01 + ci * t11;
      }
      for (int idx = tid; idx < half * kJP; idx += nt) {
        const int i = idx / kJP, k = idx % kJP;
        const double s = sn[i];
        if (s == 0.0) continue;
        const double c = cs[i];
        const int p = pp[i], q = qq[i];
        const double zkp = z[k][p], zkq = z[k][q];
        z[k][p] = c * zkp - s * zkq;
        z[k][q] = s * zkp + c * zkq;
      }
      __syncthreads();
    }
  }
  // new_row[i] = sum_k Z[k][i] * old_row[k]  =>  Jt[i][k] = Z[k][i]
  for (int idx = tid; idx < kJP * kJP; idx += nt) Jt[(int64_t)blockIdx.x * kJP * kJP + idx] = z[idx % kJP][idx / kJP];
}

__global__ void __launch_bounds__(256) jb_apply_kernel(double* __restrict__ GP, int64_t ld, int width, JbPhase ph,
                                                       int round, const double* __restrict__ Jt,
                                                       const int* __restrict__ skip) {
  extern __shared__ double sm[];
  if (skip[blockIdx.x]) return;
  double* jt = sm;                  // kJP x kJTS
  double* panel = sm + kJP * kJTS;  // kJP x kJST
  int I, J;
  block_pair(round, blockIdx.x, ph, I, J);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int c_first = blockIdx.y * kJApplyCols, c_end = min(width, c_first + kJApplyCols);
  const double* jsrc = Jt + (int64_t)blockIdx.x * kJP * kJP;
  for (int idx = threadIdx.x; idx < kJP * kJP; idx += blockDim.x) jt[(idx / kJP) * kJTS + idx % kJP] = jsrc[idx];
  // the 32 KB rotation is loaded once and applied to kJApplyCols / kJCH column chunks of the panel
  for (int c0 = c_first; c0 < c_end; c0 += kJCH) {
  load_panel(GP, ld, I, J, c0, c_end, panel);
  __syncthreads();
  double acc[8][2][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < kJP; k0 += 4) {
    double b[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) b[j] = panel[(k0 + t) * kJST + 16 * warp + 8 * j + g];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double a = jt[(8 * i + g) * kJTS + k0 + t];
#pragma unroll
      for (int j = 0; j < 2; ++j) dmma884_(acc[i][j][0], acc[i][j][1], a, b[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double* dst = GP + (int64_t)panel_row(I, J, 8 * i + g) * ld + c0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = 16 * warp + 8 * j + 2 * t;
      if (c0 + c + 1 < c_end) *reinterpret_cast<double2*>(dst + c) = make_double2(acc[i][j][0], acc[i][j][1]);
      else if (c0 + c < c_end) dst[c] = acc[i][j][0];
    }
  }
  __syncthreads();  // the panel buffer is reloaded for the next chunk
  }
}

// ---------------------------------------------------------------------------------------------
// exact null vectors: LAPACK returns an orthonormal completion for zero singular values, and the DMRG
// sweep needs it (A[site] must stay an isometry).  Rows whose norm is (numerically) zero are replaced
// by pseudo-random rows of norm ~1e-140 * max before the Jacobi iteration: the rotations
// orthogonalise them against everything else without moving the other rows (angles ~1e-140), their
// direction becomes the completion, and their singular value is reported as exactly 0.
// ---------------------------------------------------------------------------------------------
constexpr double kNullFill = 1e-140;   // norm scale of injected rows (squares stay representable)
constexpr double kNullCut = 1e-120;    // singular values below cut * s_max are reported as 0

__global__ void max_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double sh[32];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, v[i]);
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
    *out = m;
  }
}

// rows k < n of G with norms[k] <= 1e-150 * max (or max == 0) get deterministic pseudo-random entries
__global__ void __launch_bounds__(256) fill_null_rows_kernel(double* __restrict__ G, int n, int m, int64_t ld,
                                                             const double* __restrict__ norms,
                                                             const double* __restrict__ max_norm) {
  const int k = blockIdx.x;
  const double mx = *max_norm;
  if (norms[k] > 1e-150 * mx && mx > 0.0) return;
  const double scale = kNullFill * (mx > 0.0 ? mx : 1.0);
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    // splitmix64 hash of (k, i) -> uniform in (-1, 1)
    unsigned long long z = ((unsigned long long)k << 32) ^ (unsigned long long)i ^ 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    G[(int64_t)k * ld + i] = scale * ((double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0);
  }
}

// s is sorted descending: report numerically-null singular values as exactly zero
__global__ void zero_null_s_kernel(double* __restrict__ s, int n) {
  const double top = s[0];
  __syncthreads();
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
    if (k > 0 && s[k] < kNullCut * top) s[k] = 0.0;
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
                                                      const double* __restrict__ P, int64_t ldp,
                                                      const double* __restrict__ norms,
                                                      const int* __restrict__ rank, double* __restrict__ w_out,
                                                      int64_t wa, int64_t wb, double* __restrict__ p_out, int64_t pa,
                                                      int64_t pb) {
  const int k = blockIdx.x;
  const int r = rank[k];
  const double nk = norms[k];
  const double inv = nk > 0.0 ? 1.0 / nk : 0.0;
  for (int i = threadIdx.x; i < m; i += blockDim.x) w_out[r * wa + i * wb] = G[(int64_t)k * ldg + i] * inv;
  for (int j = threadIdx.x; j < n; j += blockDim.x) p_out[r * pa + j * pb] = P[(int64_t)k * ldp + j];
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

// significant-pair count seen in each Jacobi sweep of the last block-path SVD (diagnostics)
static unsigned int g_trace[64];
static int g_trace_len = 0;

static double jacobi_tol(int m) { return fmax(1e-15, 2.220446049250313e-16 * sqrt((double)m)); }

static bool fits_small(int n, int m) { return sizeof(double) * ((size_t)n * m + (size_t)n * n) <= 200 * 1024; }

// small path: G (n x m, ld m) in place, P (n x n) out
static int hestenes_small(double* G, int n, int m, double* P, cudaStream_t stream) {
  TNPY_TRY(set_max_dynamic_smem(hestenes_small_kernel, 200 * 1024));
  const size_t bytes = sizeof(double) * ((size_t)n * m + (size_t)n * n);
  const int threads = n >= 32 ? 512 : (n >= 8 ? 256 : 64);
  hestenes_small_kernel<<<1, threads, bytes, stream>>>(G, n, m, m, P, jacobi_tol(m), 60, nullptr);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

struct BlockPlan {
  int n_pad, nb, pairs, gchunks, achunks;
  int64_t ld;
  size_t gp_elems, partial_elems, jt_elems;
};

static BlockPlan block_plan(int n, int m) {
  BlockPlan p;
  p.n_pad = (n + kJP - 1) / kJP * kJP;  // whole panels => an even number of blocks, no byes
  p.nb = p.n_pad / kJB;
  p.pairs = p.nb / 2;
  p.ld = ((int64_t)m + p.n_pad + 3) / 4 * 4;
  p.gchunks = (m + kJGramCols - 1) / kJGramCols;
  p.achunks = (int)((m + p.n_pad + kJApplyCols - 1) / kJApplyCols);
  p.gp_elems = (size_t)p.n_pad * p.ld;
  p.partial_elems = (size_t)p.pairs * p.gchunks * kJP * kJP;
  p.jt_elems = (size_t)p.pairs * kJP * kJP;
  return p;
}

// large path: GP = [G | P] (n_pad x ld) in place.  n_big = number of leading (largest-norm) vectors
// outside the cluster of tiny vectors; see JbPhase.
static int hestenes_block(double* GP, const BlockPlan& p, int m, int n_big, double* partial, double* Jt, int* skip,
                          unsigned int* counter_dev, cudaStream_t stream, int* sweeps_out) {
  struct { unsigned int* host; } pinned{static_cast<unsigned int*>(thread_pinned_scratch())};
  if (!pinned.host) {
    set_error("hestenes: pinned allocation failed");
    return TNPY_ECUDA;
  }
  const size_t gram_smem = sizeof(double) * kJP * kJST;
  const size_t rot_smem = sizeof(double) * 2 * kJP * (kJP + 1);
  const size_t apply_smem = sizeof(double) * (kJP * kJTS + kJP * kJST);
  TNPY_TRY(set_max_dynamic_smem(jb_gram_kernel, (int)gram_smem));
  TNPY_TRY(set_max_dynamic_smem(jb_rotate_kernel, (int)rot_smem));
  TNPY_TRY(set_max_dynamic_smem(jb_apply_kernel, (int)apply_smem));
  const double tol = jacobi_tol(m);
  const int width = m + p.n_pad;
  const int max_sweeps = 60;
  // phase plan
  int nb_big = (n_big + kJB - 1) / kJB;
  if (nb_big < 1) nb_big = 1;
  if (nb_big > p.nb) nb_big = p.nb;
  const int nb_small = p.nb - nb_big;
  // Two-phase schedule.  The tiny, noise-dominated vectors (norm < kClusterTheta * max) are far from
  // mutually orthogonal and need the bulk of the Jacobi sweeps; they are orthogonalised among
  // themselves first, at (cluster/n)^2 of the cost of a full sweep.  The full round-robin that follows
  // then starts from two internally orthogonal groups and converges in a few sweeps.  (Skipping the
  // cluster-cluster pairs in the full phase instead does NOT work: the non-orthogonal cluster couples
  // the big-small eliminations and the iteration stalls -- measured.)
  JbPhase phases[2];
  int n_phases = 0;
  if (nb_small >= 2) {
    const int b0 = nb_big - (nb_small & 1);  // even number of blocks in the cluster round-robin
    phases[n_phases++] = JbPhase{b0, p.nb - b0, 1 << 30};
  }
  phases[n_phases++] = JbPhase{0, p.nb, 1 << 30};
  int sweep = 0;
  g_trace_len = 0;
  for (int phase = 0; phase < n_phases; ++phase) {
    const JbPhase ph = phases[phase];
    const int pairs = ph.nb / 2;
    for (; sweep < max_sweeps; ++sweep) {
      TNPY_CUDA_OK(cudaMemsetAsync(counter_dev, 0, sizeof(unsigned int), stream));
      for (int round = -1; round < ph.nb - 1; ++round) {
        // round -1: adjacent blocks with the full 64x64 inner sweep (intra-block pairs included);
        // rounds 0..nb-2: round-robin block pairs, cross pairs only.
        jb_gram_kernel<<<dim3(pairs, p.gchunks), 256, gram_smem, stream>>>(GP, p.ld, m, ph, round, p.gchunks, partial);
        jb_rotate_kernel<<<pairs, 512, rot_smem, stream>>>(partial, p.gchunks, tol, round >= 0 ? 1 : 0, ph, round, Jt,
                                                          counter_dev, skip);
        jb_apply_kernel<<<dim3(pairs, p.achunks), 256, apply_smem, stream>>>(GP, p.ld, width, ph, round, Jt, skip);
      }
      TNPY_LAUNCH_OK();
      count_launch(3 * ph.nb - 1);
      TNPY_CUDA_OK(cudaMemcpyAsync(pinned.host, counter_dev, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
      TNPY_CUDA_OK(cudaStreamSynchronize(stream));
      if (g_trace_len < 64) g_trace[g_trace_len++] = *pinned.host | (ph.nb < p.nb ? 0x80000000u : 0u);
      if (*pinned.host == 0u) {
        ++sweep;
        break;
      }
    }
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
  const int n = rows < cols ? rows : cols, m = rows < cols ? cols : rows;
  size_t total = Workspace::need(n) + Workspace::need(n, sizeof(int)) + 1024;
  if (fits_small(n, m)) return total + Workspace::need((size_t)n * m) + Workspace::need((size_t)n * n);
  const BlockPlan p = block_plan(n, m);
  return total + Workspace::need(p.gp_elems) + Workspace::need(p.partial_elems) + Workspace::need(p.jt_elems) +
         Workspace::need(p.pairs, sizeof(int)) + Workspace::need(n, sizeof(int)) + 512;
}

static int g_last_svd_sweeps = 0;
extern "C" int tnpy_last_svd_sweeps(void) { return g_last_svd_sweeps; }
extern "C" int tnpy_last_svd_trace(unsigned int* counts, int max_counts) {
  int n = 0;
  for (; n < tnpy::g_trace_len && n < max_counts; ++n) counts[n] = tnpy::g_trace[n];
  return n;
}

extern "C" int tnpy_svd(double* A, int rows, int cols, double* U, double* s, double* Vt, void* workspace,
                        size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(A && U && s && Vt && rows > 0 && cols > 0, "bad argument");
  const bool tall = rows >= cols;
  const int n = tall ? cols : rows, m = tall ? rows : cols;
  Workspace ws(workspace, workspace_bytes);
  double* norms = ws.take<double>(n);
  int* rank = ws.take<int>(n);
  unsigned int* counter = ws.take<unsigned int>(64);
  const double* G = nullptr;
  const double* P = nullptr;
  int64_t ldg = m, ldp = n;
  if (fits_small(n, m)) {
    double* Gt = ws.take<double>((size_t)n * m);
    double* Pm = ws.take<double>((size_t)n * n);
    if (!norms || !rank || !counter || !Gt || !Pm) {
      set_error("tnpy_svd: workspace too small");
      return TNPY_EWORKSPACE;
    }
    double* Gm = A;
    if (tall) {
      TNPY_TRY(transpose2d(A, rows, cols, Gt, nullptr, stream));
      Gm = Gt;
    }
    row_norm_kernel<<<n, 256, 0, stream>>>(Gm, n, m, m, norms);
    TNPY_LAUNCH_OK();
    max_kernel<<<1, 256, 0, stream>>>(norms, n, reinterpret_cast<double*>(counter) + 1);
    TNPY_LAUNCH_OK();
    fill_null_rows_kernel<<<n, 256, 0, stream>>>(Gm, n, m, m, norms, reinterpret_cast<double*>(counter) + 1);
    TNPY_LAUNCH_OK();
    TNPY_TRY(hestenes_small(Gm, n, m, Pm, stream));
    G = Gm;
    P = Pm;
    g_last_svd_sweeps = -1;
  } else {
    const BlockPlan p = block_plan(n, m);
    double* GP = ws.take<double>(p.gp_elems);
    double* partial = ws.take<double>(p.partial_elems);
    double* Jt = ws.take<double>(p.jt_elems);
    if (!norms || !rank || !counter || !GP || !partial || !Jt) {
      set_error("tnpy_svd: workspace too small");
      return TNPY_EWORKSPACE;
    }
    int* skip = ws.take<int>(p.pairs);
    int* src = ws.take<int>(n);
    int* info = ws.take<int>(64);
    if (!skip || !src || !info) {
      set_error("tnpy_svd: workspace too small");
      return TNPY_EWORKSPACE;
    }
    // sort the vectors by decreasing norm (de Rijk-style ordering) and find the cluster of tiny ones
    vec_norm_kernel<<<tall ? ceil_div(cols, 256) : min(rows, sm_count() * 8), 256, 0, stream>>>(A, rows, cols,
                                                                                             tall ? 1 : 0, norms);
    TNPY_LAUNCH_OK();
    rank_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(norms, n, rank, s);
    TNPY_LAUNCH_OK();
    invert_rank_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(rank, s, n, kClusterTheta, src, info);
    TNPY_LAUNCH_OK();
    int n_big = n;
    TNPY_CUDA_OK(cudaMemcpyAsync(&n_big, info, sizeof(int), cudaMemcpyDeviceToHost, stream));
    TNPY_CUDA_OK(cudaStreamSynchronize(stream));
    jb_init_kernel<<<sm_count() * 8, 256, 0, stream>>>(A, rows, cols, tall ? 1 : 0, src, GP, n, m, p.n_pad, p.ld);
    TNPY_LAUNCH_OK();
    fill_null_rows_kernel<<<n, 256, 0, stream>>>(GP, n, m, p.ld, s, s);  // sorted: norm of row k is s[k], max s[0]
    TNPY_LAUNCH_OK();
    TNPY_TRY(hestenes_block(GP, p, m, n_big, partial, Jt, skip, counter, stream, &g_last_svd_sweeps));
    G = GP;
    P = GP + m;
    ldg = p.ld;
    ldp = p.ld;
  }
  row_norm_kernel<<<n, 256, 0, stream>>>(G, n, m, ldg, norms);
  TNPY_LAUNCH_OK();
  rank_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(norms, n, rank, s);
  TNPY_LAUNCH_OK();
  zero_null_s_kernel<<<1, 256, 0, stream>>>(s, n);
  TNPY_LAUNCH_OK();
  if (tall)  // U = W^T (rows x n): U[i][r] ; Vt = P (n x cols): Vt[r][j]
    scatter_kernel<<<n, 256, 0, stream>>>(G, n, m, ldg, P, ldp, norms, rank, U, 1, n, Vt, cols, 1);
  else  // Vt = W (n x cols): Vt[r][i] ; U = P^T (rows x n): U[j][r]
    scatter_kernel<<<n, 256, 0, stream>>>(G, n, m, ldg, P, ldp, norms, rank, Vt, cols, 1, U, 1, n);
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
