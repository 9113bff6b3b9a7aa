// On-device lowest eigenpair of the symmetric-definite pencil  A x = lambda M x  with A = H_eff of one
// environment and M = H_eff of a second one (ShiftInvertDMRG: A from the MPO of H - eps, M from its
// square; reference finite_dmrg.py:341-355 -> primme.eigsh(A, M=...)).
//
// Generalised Davidson without preconditioner: an orthonormal basis V with A V and M V kept alongside,
// the projected pencil (V^T A V, V^T M V) solved in one CTA (Cholesky of the projected M, reduction to
// a standard problem, parallel Jacobi), residual r = A x - theta M x as the expansion direction
// (orthogonalised twice against V), restart with the current Ritz vector plus the residual direction.
// One 64-byte status read-back per iteration, vectors never leave the device.
#include <math.h>

#include "chol.cuh"

namespace tnpy {

int heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l, int r, int wl,
               int wr, int d, int flags, Workspace& ws, cudaStream_t stream);
int multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h, int mode,
              cudaStream_t stream, const int* skip = nullptr);
int multi_axpy(const double* V, int64_t ldv, int m, const double* h, double* w, int64_t n, double* nrm_out,
               cudaStream_t stream, const int* skip = nullptr);
int scale_copy(const double* x, double* out, int64_t n, double alpha, const double* s_dev, int inv,
               cudaStream_t stream);
int axpy(double alpha, const double* a_dev, const double* x, double* y, int64_t n, cudaStream_t stream);
int combine(const double* V, int64_t ldv, int m, const double* c, int64_t ldc_, int nout, double* out, int64_t ldo,
            int64_t n, cudaStream_t stream);

constexpr int kMaxG = 40;  // largest projected pencil (3 x 40 x 40 doubles of shared memory)

// status record
enum { GS_THETA = 0, GS_RESID = 1, GS_NA = 2, GS_NM = 3, GS_DONE = 4, GS_FAIL = 5, GS_TNORM = 6, GS_SIZE = 8 };

// Fold the new columns ha / hm (length j+1) into GA / GM, then solve the (j+1)x(j+1) pencil.
// y: coefficients of the lowest Ritz vector (y^T GM y = 1); theta -> status[GS_THETA].
__global__ void __launch_bounds__(256) geig_small_kernel(double* __restrict__ GA, double* __restrict__ GM,
                                                         const double* __restrict__ ha, const double* __restrict__ hm,
                                                         int j, double* __restrict__ y, double* __restrict__ status) {
  // leading dimension kMaxG + 1 (odd): column walks are spread over the shared-memory banks
  __shared__ double a[kMaxG][kMaxG + 1];  // GA -> C^-1 GA C^-T -> its eigenvalues on the diagonal
  __shared__ double b[kMaxG][kMaxG + 1];  // GM -> Cholesky factor C (lower)
  __shared__ double z[kMaxG][kMaxG + 1];  // eigenvectors
  __shared__ double cs[kMaxG], sn[kMaxG], red[72];
  __shared__ int pp[kMaxG], qq[kMaxG];
  __shared__ int fail;
  const int m = j + 1, tid = threadIdx.x, nt = blockDim.x;
  if (tid < m) {
    GA[tid * kMaxG + j] = GA[j * kMaxG + tid] = ha[tid];
    GM[tid * kMaxG + j] = GM[j * kMaxG + tid] = hm[tid];
  }
  if (tid == 0) fail = 0;
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += nt) {
    a[idx / m][idx % m] = GA[(idx / m) * kMaxG + idx % m];
    b[idx / m][idx % m] = GM[(idx / m) * kMaxG + idx % m];
  }
  __syncthreads();
  // Cholesky b = C C^T (right-looking, lower triangle)
  for (int k = 0; k < m; ++k) {
    if (tid == 0) {
      const double dkk = b[k][k];
      if (!(dkk > 0.0)) fail = 1;
      b[k][k] = sqrt(dkk > 0.0 ? dkk : 1.0);
    }
    __syncthreads();
    const double ckk = b[k][k];
    for (int i = k + 1 + tid; i < m; i += nt) b[i][k] /= ckk;
    __syncthreads();
    const int rem = m - k - 1;
    for (int idx = tid; idx < rem * rem; idx += nt) {
      const int i = k + 1 + idx / rem, c = k + 1 + idx % rem;
      if (c <= i) b[i][c] -= b[i][k] * b[c][k];
    }
    __syncthreads();
  }
  // a <- C^-1 a   (forward substitution, one thread per column)
  for (int c = tid; c < m; c += nt)
    for (int i = 0; i < m; ++i) {
      double v = a[i][c];
      for (int p = 0; p < i; ++p) v -= b[i][p] * a[p][c];
      a[i][c] = v / b[i][i];
    }
  __syncthreads();
  // a <- a C^-T   (same substitution on the rows)
  for (int rrow = tid; rrow < m; rrow += nt)
    for (int i = 0; i < m; ++i) {
      double v = a[rrow][i];
      for (int p = 0; p < i; ++p) v -= b[i][p] * a[rrow][p];
      a[rrow][i] = v / b[i][i];
    }
  __syncthreads();
  // symmetrise against rounding, then Jacobi
  for (int idx = tid; idx < m * m; idx += nt) {
    const int i = idx / m, c = idx % m;
    if (i < c) {
      const double v = 0.5 * (a[i][c] + a[c][i]);
      a[i][c] = v;
      a[c][i] = v;
    }
  }
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += nt) z[idx / m][idx % m] = (idx / m == idx % m) ? 1.0 : 0.0;
  __syncthreads();
  if (m > 1) {
    const int me = (m + 1) & ~1, half = me / 2;
    double o_prev = 1e300;  // thread 0 only
    for (int sweep = 0; sweep < 40; ++sweep) {
      double off = 0.0, dia = 0.0;
      for (int idx = tid; idx < m * m; idx += nt) {
        const double v = a[idx / m][idx % m];
        if (idx / m == idx % m) dia += v * v; else off += v * v;
      }
      off = warp_sum(off);
      dia = warp_sum(dia);
      if ((tid & 31) == 0) { red[tid >> 5] = off; red[32 + (tid >> 5)] = dia; }
      __syncthreads();
      if (tid == 0) {
        double o = 0.0, d2 = 0.0;
        for (int w = 0; w < (nt >> 5); ++w) { o += red[w]; d2 += red[32 + w]; }
        // converged: off-diagonal mass below 1e-31 of the diagonal mass, or at its rounding floor -- Jacobi converges
        // quadratically, so a sweep that no longer shrinks an already tiny off-diagonal mass by 4x has hit the floor
        // ((m eps)^2-ish; a fixed 1e-31 alone is below it for m > ~16 and made those solves run all 40 sweeps: 1.5 ms)
        const bool stalled = o <= 1e-24 * d2 && o >= 0.25 * o_prev;
        red[64] = (o <= 1e-31 * d2 || o == 0.0 || stalled) ? 1.0 : 0.0;
        o_prev = o;
      }
      __syncthreads();
      const bool converged = red[64] != 0.0;
      __syncthreads();
      if (converged) break;
      for (int round = 0; round < me - 1; ++round) {
        if (tid < half) {
          int p = (tid == 0) ? me - 1 : (round + tid) % (me - 1);
          int q = (round + me - 1 - tid) % (me - 1);
          if (p > q) { const int t = p; p = q; q = t; }
          double c = 1.0, s = 0.0;
          if (q < m && a[p][q] != 0.0) {
            const double tau = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
          }
          cs[tid] = c; sn[tid] = s; pp[tid] = p; qq[tid] = q;
        }
        __syncthreads();
        for (int idx = tid; idx < half * m; idx += nt) {
          const int i = idx / m, k = idx % m, p = pp[i], q = qq[i];
          if (q >= m || sn[i] == 0.0) continue;
          const double c = cs[i], s = sn[i];
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
          const double zkp = z[k][p], zkq = z[k][q];
          z[k][p] = c * zkp - s * zkq;
          z[k][q] = s * zkp + c * zkq;
        }
        __syncthreads();
        for (int idx = tid; idx < half * m; idx += nt) {
          const int i = idx / m, k = idx % m, p = pp[i], q = qq[i];
          if (q >= m || sn[i] == 0.0) continue;
          const double c = cs[i], s = sn[i];
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        __syncthreads();
      }
    }
  }
  if (tid == 0) {
    int lo = 0;
    for (int i = 1; i < m; ++i)
      if (a[i][i] < a[lo][lo]) lo = i;
    status[GS_THETA] = a[lo][lo];
    status[GS_FAIL] = fail ? 1.0 : 0.0;
    // y = C^-T z_lo  (back substitution), sign: first component non-negative
    double yy[kMaxG];
    for (int i = m - 1; i >= 0; --i) {
      double v = z[i][lo];
      for (int p = i + 1; p < m; ++p) v -= b[p][i] * yy[p];
      yy[i] = v / b[i][i];
    }
    const double sgn = yy[0] < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < m; ++i) y[i] = sgn * yy[i];
  }
}

// status[DONE] = ||r|| <= tol * (||A x|| + |theta| ||M x||)
__global__ void geig_status_kernel(double* __restrict__ status, double tol) {
  const double scale = status[GS_NA] + fabs(status[GS_THETA]) * status[GS_NM];
  status[GS_DONE] = (status[GS_RESID] <= tol * scale) ? 1.0 : 0.0;
}

__global__ void geig_reset_kernel(double* __restrict__ GA, double* __restrict__ GM) {
  for (int idx = threadIdx.x; idx < kMaxG * kMaxG; idx += blockDim.x) GA[idx] = GM[idx] = 0.0;
}

static void geig_sizes(int64_t n, int ncv_in, int& ncv) {
  ncv = ncv_in <= 0 ? 24 : ncv_in;
  if (ncv > kMaxG) ncv = kMaxG;
  if (ncv < 3) ncv = 3;
  if (ncv > n) ncv = (int)n;
}

}  // namespace tnpy

using namespace tnpy;

extern "C" size_t tnpy_geig_workspace_bytes(int l, int r, int wl_a, int wr_a, int wl_m, int wr_m, int d, int ncv_in) {
  const int64_t n = (int64_t)l * d * r, ldv = n + (n & 1);
  int ncv;
  geig_sizes(n, ncv_in, ncv);
  const size_t chain_a = tnpy_heff_workspace_bytes(l, r, wl_a, wr_a, d), chain_m = tnpy_heff_workspace_bytes(l, r, wl_m, wr_m, d);
  return 3 * Workspace::need((size_t)(ncv + 1) * ldv) + 3 * Workspace::need(ldv) + 2 * Workspace::need(kMaxG * kMaxG) +
         4 * Workspace::need(64) + (chain_a > chain_m ? chain_a : chain_m) + 1024;
}

extern "C" int tnpy_geig_lowest(const double* LA, const double* WA, const double* RA, const double* LM,
                                const double* WM, const double* RM, double* psi, int l, int r, int wl_a, int wr_a,
                                int wl_m, int wr_m, int d, int flags_a, double tol, int max_iter, int ncv_in,
                                double* stats_host, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(psi && WA && WM, "null pointer");
  TNPY_CHECK_ARG(l > 0 && r > 0 && d > 0 && wl_a > 0 && wr_a > 0 && wl_m > 0 && wr_m > 0, "non-positive dimension");
  const int64_t n = (int64_t)l * d * r, ldv = n + (n & 1);
  int ncv;
  geig_sizes(n, ncv_in, ncv);
  if (tol <= 0.0) tol = 2.220446049250313e-16 * 1e4;
  if (max_iter <= 0) max_iter = 2000;
  Workspace ws(workspace, workspace_bytes);
  double* V = ws.take<double>((size_t)(ncv + 1) * ldv);
  double* AV = ws.take<double>((size_t)(ncv + 1) * ldv);
  double* MV = ws.take<double>((size_t)(ncv + 1) * ldv);
  double* t = ws.take<double>(ldv);
  double* tmp = ws.take<double>(ldv);
  double* tmp2 = ws.take<double>(ldv);
  double* GA = ws.take<double>(kMaxG * kMaxG);
  double* GM = ws.take<double>(kMaxG * kMaxG);
  double* ha = ws.take<double>(64);
  double* hm = ws.take<double>(64);
  double* y = ws.take<double>(64);
  double* status = ws.take<double>(64);
  if (!V || !AV || !MV || !t || !tmp || !tmp2 || !GA || !GM || !ha || !hm || !y || !status) {
    set_error("tnpy_geig_lowest: workspace too small (%zu bytes given)", workspace_bytes);
    return TNPY_EWORKSPACE;
  }
  const size_t chain_off = ws.used;
  double* hst = static_cast<double*>(thread_pinned_scratch());
  if (!hst) {
    set_error("tnpy_geig_lowest: pinned status allocation failed");
    return TNPY_ECUDA;
  }
  auto apply = [&](const double* L, const double* W, const double* R, int wl, int wr, int flags, const double* x,
                   double* out) {
    Workspace chain(static_cast<char*>(workspace) + chain_off, workspace_bytes - chain_off);
    return heff_apply(L, W, R, x, out, l, r, wl, wr, d, flags, chain, stream);
  };

  geig_reset_kernel<<<1, 256, 0, stream>>>(GA, GM);
  TNPY_LAUNCH_OK();
  TNPY_TRY(multi_dot(psi, ldv, 1, psi, n, status + GS_TNORM, 1, stream));
  TNPY_TRY(scale_copy(psi, V, n, 1.0, status + GS_TNORM, 1, stream));
  int j = 0, iters = 0, restarts = 0;
  bool done = false;
  while (true) {
    double* vj = V + (int64_t)j * ldv;
    double* avj = AV + (int64_t)j * ldv;
    double* mvj = MV + (int64_t)j * ldv;
    TNPY_TRY(apply(LA, WA, RA, wl_a, wr_a, flags_a, vj, avj));
    TNPY_TRY(apply(LM, WM, RM, wl_m, wr_m, 0, vj, mvj));
    ++iters;
    TNPY_TRY(multi_dot(V, ldv, j + 1, avj, n, ha, 0, stream));
    TNPY_TRY(multi_dot(V, ldv, j + 1, mvj, n, hm, 0, stream));
    geig_small_kernel<<<1, 256, 0, stream>>>(GA, GM, ha, hm, j, y, status);
    TNPY_LAUNCH_OK();
    const int m = j + 1;
    // residual r = A x - theta M x with x = V y
    TNPY_TRY(combine(AV, ldv, m, y, 1, 1, t, ldv, n, stream));
    TNPY_TRY(multi_dot(t, ldv, 1, t, n, status + GS_NA, 1, stream));
    TNPY_TRY(combine(MV, ldv, m, y, 1, 1, tmp, ldv, n, stream));
    TNPY_TRY(multi_dot(tmp, ldv, 1, tmp, n, status + GS_NM, 1, stream));
    TNPY_TRY(axpy(-1.0, status + GS_THETA, tmp, t, n, stream));
    TNPY_TRY(multi_dot(t, ldv, 1, t, n, status + GS_RESID, 1, stream));
    geig_status_kernel<<<1, 1, 0, stream>>>(status, tol);
    TNPY_LAUNCH_OK();
    TNPY_CUDA_OK(cudaMemcpyAsync(hst, status, sizeof(double) * GS_SIZE, cudaMemcpyDeviceToHost, stream));
    TNPY_CUDA_OK(cudaStreamSynchronize(stream));
    if (hst[GS_FAIL] != 0.0) {
      set_error("tnpy_geig_lowest: projected M is not positive definite (basis size %d)", m);
      return TNPY_EINVAL;
    }
    done = hst[GS_DONE] != 0.0 || m >= n;
    if (done || iters >= max_iter) {
      // psi = V y, normalised like primme / scipy.linalg.eigh(a, b): x^T M x = 1 (y^T GM y = 1 by construction)
      TNPY_TRY(combine(V, ldv, m, y, 1, 1, psi, ldv, n, stream));
      break;
    }
    if (m == ncv) {
      // restart: basis = {x, residual direction}; A x and M x come from the stored products
      TNPY_TRY(combine(V, ldv, m, y, 1, 1, tmp, ldv, n, stream));
      TNPY_TRY(multi_dot(tmp, ldv, 1, tmp, n, status + GS_TNORM, 1, stream));
      TNPY_TRY(combine(AV, ldv, m, y, 1, 1, tmp2, ldv, n, stream));
      TNPY_TRY(scale_copy(tmp2, AV, n, 1.0, status + GS_TNORM, 1, stream));
      TNPY_TRY(combine(MV, ldv, m, y, 1, 1, tmp2, ldv, n, stream));
      TNPY_TRY(scale_copy(tmp2, MV, n, 1.0, status + GS_TNORM, 1, stream));
      TNPY_TRY(scale_copy(tmp, V, n, 1.0, status + GS_TNORM, 1, stream));
      geig_reset_kernel<<<1, 256, 0, stream>>>(GA, GM);
      TNPY_LAUNCH_OK();
      TNPY_TRY(multi_dot(V, ldv, 1, AV, n, ha, 0, stream));
      TNPY_TRY(multi_dot(V, ldv, 1, MV, n, hm, 0, stream));
      geig_small_kernel<<<1, 256, 0, stream>>>(GA, GM, ha, hm, 0, y, status);
      TNPY_LAUNCH_OK();
      j = 0;
      ++restarts;
    }
    // expand with the residual, orthogonalised twice against the basis
    double* vnext = V + (int64_t)(j + 1) * ldv;
    TNPY_TRY(multi_dot(V, ldv, j + 1, t, n, ha, 0, stream));
    TNPY_TRY(multi_axpy(V, ldv, j + 1, ha, t, n, nullptr, stream));
    TNPY_TRY(multi_dot(V, ldv, j + 1, t, n, ha, 0, stream));
    TNPY_TRY(multi_axpy(V, ldv, j + 1, ha, t, n, status + GS_TNORM, stream));
    TNPY_TRY(scale_copy(t, vnext, n, 1.0, status + GS_TNORM, 1, stream));
    ++j;
  }
  TNPY_CUDA_OK(cudaStreamSynchronize(stream));
  if (stats_host) {
    stats_host[0] = hst[GS_THETA];
    stats_host[1] = hst[GS_RESID];
    stats_host[2] = (double)iters;
    stats_host[3] = (double)restarts;
    stats_host[4] = done ? 1.0 : 0.0;
    stats_host[5] = hst[GS_NA];
  }
  if (!done) {
    set_error("tnpy_geig_lowest: not converged after %d iterations (resid %.3e)", iters, hst[GS_RESID]);
    return TNPY_ENOCONV;
  }
  return TNPY_OK;
}

// ---------------------------------------------------------------------------------------------
// Dense route (reference: scipy.linalg.eigh(a, b), finite_dmrg.py:344-348).  The projected H^2 is so
// ill-conditioned (kappa ~ 1e9 already at N = 256) that no inverse-free Krylov iteration converges
// in a useful number of steps, so up to a few thousand unknowns the pencil is solved densely:
//   b = Q diag(s) Q^T (Jacobi SVD of the SPD matrix),  X = Q diag(s)^-1/2,  S = X^T a X,
//   lowest eigenpair (lambda, z) of S (Jacobi),  x = X z  (x^T b x = 1).
// ---------------------------------------------------------------------------------------------
namespace tnpy {
// X[p][k] = U[p][k] / sqrt(s[k]);  Xt[k][p] = Vt[k][p] / sqrt(s[k])
__global__ void geig_scale_kernel(const double* __restrict__ U, const double* __restrict__ Vt,
                                  const double* __restrict__ s, int n, double* __restrict__ X,
                                  double* __restrict__ Xt) {
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), k = (int)(e % n);
    const double sk = s[k], si = s[i];
    X[e] = sk > 0.0 ? U[e] / sqrt(sk) : 0.0;
    Xt[e] = si > 0.0 ? Vt[e] / sqrt(si) : 0.0;
  }
}
}  // namespace tnpy

extern "C" size_t tnpy_geig_dense_workspace_bytes(int n) {
  const size_t nn = (size_t)n * n;
  return 6 * Workspace::need(nn) + 2 * Workspace::need(n) + tnpy_svd_workspace_bytes(n, n) + tnpy_eigh_workspace_bytes(n) + 2048;
}

extern "C" int tnpy_geig_dense_lowest(double* a, double* b, int n, double* theta_dev, double* x, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(a && b && theta_dev && x && n > 0, "bad argument");
  Workspace ws(workspace, workspace_bytes);
  const size_t nn = (size_t)n * n;
  double* U = ws.take<double>(nn);
  double* Vt = ws.take<double>(nn);
  double* X = ws.take<double>(nn);
  double* Xt = ws.take<double>(nn);
  double* Y = ws.take<double>(nn);
  double* S = ws.take<double>(nn);
  double* s = ws.take<double>(n);
  double* z = ws.take<double>(n);
  if (!U || !Vt || !X || !Xt || !Y || !S || !s || !z) {
    set_error("tnpy_geig_dense_lowest: workspace too small");
    return TNPY_EWORKSPACE;
  }
  char* rest = static_cast<char*>(workspace) + ws.used;
  const size_t rest_bytes = workspace_bytes - ws.used;
  TNPY_TRY(tnpy_svd(b, n, n, U, s, Vt, rest, rest_bytes, stream_));  // b is destroyed
  geig_scale_kernel<<<sm_count() * 4, 256, 0, stream>>>(U, Vt, s, n, X, Xt);
  TNPY_LAUNCH_OK();
  // Y = a^T X (= a X, a symmetric);  S = X^T Y
  TNPY_TRY(gemm_tn(a, n, X, n, plain_out(Y, n, n), n, n, n, 0, TNPY_GEMM_AUTO, stream));
  TNPY_TRY(gemm_tn(X, n, Y, n, plain_out(S, n, n), n, n, n, 0, TNPY_GEMM_AUTO, stream));
  TNPY_TRY(tnpy_eigh_lowest(S, n, theta_dev, z, rest, rest_bytes, stream_));
  // x = X z = sum_k z[k] * Xt[k][:]
  TNPY_TRY(combine(Xt, n, n, z, 1, 1, x, n, n, stream));
  return TNPY_OK;
}

// ---- the same pencil through a Cholesky factor of b and the on-device Lanczos solver ----------------------------
// b = D U^T U D (D = sqrt(diag b): the factorisation then only sees the conditioning of the scaled matrix, U upper
// triangular), S = U^-T (D^-1 a D^-1) U^-1, lowest eigenpair (theta, z) of S by the thick-restart Lanczos of
// csrc/lanczos.cu -- S enters it as a "left environment" with one channel (H_eff y = S^T y), so the vectors stay on the
// device and the fused small-site steps apply -- and x = D^-1 U^-1 z, x^T b x = 1.
//
// Everything O(n^3) is a TN GEMM on the FP64 tensor pipe, which is why the factor is the *upper* one, built by row
// panels of 512: with U_kk = L^T from the small blocked Cholesky of the diagonal block (csrc/qr.cu, also its explicit
// inverse),
//     row panel   U_k,rest = L^-1 G_k,rest                       = (L^-T)^T G_k,rest          K = 512
//     trailing    G_rest,rest -= U_k,rest^T U_k,rest              = (-U_k,rest)^T U_k,rest     K = 512
//     U^T Y = R   Y_k = L^-1 (R_k - U_(rows < k, cols k)^T Y_(rows < k))                       K = 512 k
// all have the contracted index slowest in both operands, as gemm_tn wants; S = U^-T (U^-T a')^T is two such forward
// substitutions with a transpose in between.  Unlike tnpy_geig_dense_lowest (Jacobi SVD of b, Jacobi eigensolve of S:
// O(n^3) per sweep of each) this is ~2.7 n^3 tensor-pipe flops plus O(n^2) per Lanczos step, which is what makes
// pencils of 10^4 unknowns practical.  The lowest eigenvalue of S is 1 / (E - eps) for the level just below the shift:
// an outlier of a spectrum clustered around zero, so Lanczos needs few steps.
namespace tnpy {
namespace {
constexpr int kPanel = 512;

// out (rows x cols, ld_out) = scale_row ? in * scale_row[i] * scale_col[j] : in, zero outside the n x n corner of `in`
__global__ void pad_scale_kernel(const double* __restrict__ in, int n, int64_t ld_in, double* __restrict__ out, int np,
                                 const double* __restrict__ dinv) {
  const int64_t total = (int64_t)np * np;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / np), j = (int)(e % np);
    double v = 0.0;
    if (i < n && j < n) {
      v = in[(int64_t)i * ld_in + j];
      if (dinv) v *= dinv[i] * dinv[j];
    }
    out[e] = v;
  }
}
// dst (rows x cols, ld_dst) = src (rows x cols, ld_src); neg (same shape as dst, optional) = -src
__global__ void block_copy_kernel(const double* __restrict__ src, int64_t ld_src, double* __restrict__ dst, int64_t ld_dst,
                                  double* __restrict__ neg, int64_t ld_neg, int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / cols), j = (int)(e % cols);
    const double v = src[(int64_t)i * ld_src + j];
    dst[(int64_t)i * ld_dst + j] = v;
    if (neg) neg[(int64_t)i * ld_neg + j] = -v;
  }
}
// dst (rows x cols, ld_dst) -= sub (rows x cols, ld_sub)
__global__ void block_sub_kernel(double* __restrict__ dst, int64_t ld_dst, const double* __restrict__ sub, int64_t ld_sub,
                                 int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / cols), j = (int)(e % cols);
    dst[(int64_t)i * ld_dst + j] -= sub[(int64_t)i * ld_sub + j];
  }
}
// v[i] = z[i] - sum_j U[i][j] y[j] over `cols` columns (one warp per row); U, y already offset to the panel
__global__ void __launch_bounds__(256) panel_gemv_sub_kernel(const double* __restrict__ U, int64_t ld, const double* __restrict__ y,
                                                             const double* __restrict__ z, double* __restrict__ v, int rows,
                                                             int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  double s = 0.0;
  for (int j = lane; j < cols; j += 32) s = fma(U[(int64_t)row * ld + j], y[j], s);
  s = warp_sum(s);
  if (lane == 0) v[row] = z[row] - s;
}
// deterministic start vector with every component non-zero
__global__ void start_vector_kernel(double* __restrict__ z, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    unsigned h = (unsigned)i * 2654435761u + 12345u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    z[i] = 0.5 + (double)(h & 0xffffu) / 65536.0;
  }
}
__global__ void scale_by_kernel(double* __restrict__ x, const double* __restrict__ dinv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= dinv[i];
}

int grid_for(int64_t total) {
  const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

struct PanelScratch {
  double *diag, *tmp, *dk;           // B x B each (dk: B x 64): the small Cholesky of a diagonal block
  double *panel, *panel_neg;         // B x np
  double *cinv_all, *cinvt_all;      // (np / B) blocks of B x B: L_kk^-1 and its transpose
};

// G (np x np, symmetric, full storage) -> U in its upper triangle (row panels; the diagonal blocks are left as they
// were -- only their inverses, kept in ps.cinv_all / cinvt_all, are used afterwards)
int cholesky_upper_blocked(double* G, int np, int B, const PanelScratch& ps, int* fail, cudaStream_t stream) {
  const int nb = np / B;
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * B, rem = np - k0 - B;
    double* gkk = G + (int64_t)k0 * np + k0;
    block_copy_kernel<<<grid_for((int64_t)B * B), 256, 0, stream>>>(gkk, np, ps.diag, B, nullptr, 0, B, B);
    TNPY_LAUNCH_OK();
    double* cinv = ps.cinv_all + (int64_t)k * B * B;
    double* cinvt = ps.cinvt_all + (int64_t)k * B * B;
    TNPY_TRY(cholesky_inverse(ps.diag, B, cinv, ps.tmp, ps.dk, fail, stream));
    TNPY_TRY(transpose(cinv, B, B, B, cinvt, B, nullptr, stream));
    if (rem == 0) break;
    double* gpanel = gkk + B;  // rows of block k, columns to the right
    // U_k,rest = L^-1 G_k,rest: C[i][j] = sum_l cinvt[l][i] G[k0 + l][k0 + B + j]
    TNPY_TRY(gemm_tn(cinvt, B, gpanel, np, plain_out(ps.panel, rem, B), B, rem, B, 0, TNPY_GEMM_FP64, stream));
    block_copy_kernel<<<grid_for((int64_t)B * rem), 256, 0, stream>>>(ps.panel, rem, gpanel, np, ps.panel_neg, rem, B, rem);
    TNPY_LAUNCH_OK();
    // G_rest,rest += (-U_k,rest)^T U_k,rest
    TNPY_TRY(gemm_tn(ps.panel_neg, rem, ps.panel, rem, plain_out(gkk + (int64_t)B * np + B, np, rem), rem, rem, B, 1,
                     TNPY_GEMM_FP64, stream));
  }
  return TNPY_OK;
}

// R (np x cols, ld) <- U^-T R, block row by block row (U in the upper triangle of G, diagonal blocks through cinvt_all)
int forward_substitute(const double* G, int np, int B, const PanelScratch& ps, double* R, int64_t ld, int cols,
                       cudaStream_t stream) {
  const int nb = np / B;
  for (int k = 0; k < nb; ++k) {
    const int k0 = k * B;
    double* rk = R + (int64_t)k0 * ld;
    if (k > 0) {
      // T = U[0:k0, block k]^T Y[0:k0, :]   (K = k0)
      TNPY_TRY(gemm_tn(G + k0, np, R, ld, plain_out(ps.panel, cols, B), B, cols, k0, 0, TNPY_GEMM_FP64, stream));
      block_sub_kernel<<<grid_for((int64_t)B * cols), 256, 0, stream>>>(rk, ld, ps.panel, cols, B, cols);
      TNPY_LAUNCH_OK();
    }
    // Y_k = L^-1 (.)
    TNPY_TRY(gemm_tn(ps.cinvt_all + (int64_t)k * B * B, B, rk, ld, plain_out(ps.panel, cols, B), B, cols, B, 0, TNPY_GEMM_FP64,
                     stream));
    block_copy_kernel<<<grid_for((int64_t)B * cols), 256, 0, stream>>>(ps.panel, cols, rk, ld, nullptr, 0, B, cols);
    TNPY_LAUNCH_OK();
  }
  return TNPY_OK;
}

// y (np) = U^-1 z, from the last block row upwards; v: B doubles of scratch
int backward_substitute(const double* G, int np, int B, const PanelScratch& ps, const double* z, double* y, double* v,
                        cudaStream_t stream) {
  const int nb = np / B;
  for (int k = nb - 1; k >= 0; --k) {
    const int k0 = k * B, rem = np - k0 - B;
    const double* rhs = z + k0;
    if (rem > 0) {
      panel_gemv_sub_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(G + (int64_t)k0 * np + k0 + B, np, y + k0 + B, z + k0, v, B, rem);
      TNPY_LAUNCH_OK();
      rhs = v;
    }
    // y_k = U_kk^-1 rhs = L^-T rhs: y_k[i] = sum_l cinv[l][i] rhs[l]
    TNPY_TRY(gemm_tn(ps.cinv_all + (int64_t)k * B * B, B, rhs, 1, plain_out(y + k0, 1, B), B, 1, B, 0, TNPY_GEMM_GENERIC, stream));
  }
  return TNPY_OK;
}
}  // namespace
}  // namespace tnpy

static size_t chol_panel(int np) { return (size_t)(np < tnpy::kPanel ? np : tnpy::kPanel); }

extern "C" size_t tnpy_geig_chol_workspace_bytes(int n) {
  if (n <= 0) return 0;
  const size_t np = (size_t)chol_padded_dim(n), B = chol_panel((int)np);
  return 3 * Workspace::need(np * np) + Workspace::need((size_t)n * n) + 2 * Workspace::need(B * B) +
         Workspace::need(B * kCholBlock) + 2 * Workspace::need(B * np) + 2 * Workspace::need(np * B) + 3 * Workspace::need(np) +
         Workspace::need(B) + Workspace::need(64, 1) + tnpy_eig_workspace_bytes(n, 1, 1, 1, 1, 0) + 8192;
}

extern "C" int tnpy_geig_chol_lowest(const double* a, const double* b, int n, double tol, int max_matvec,
                                     double* theta_dev, double* x, double* stats_host, void* workspace,
                                     size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(a && b && theta_dev && x && n > 0, "bad argument");
  const int np = chol_padded_dim(n), B = (int)chol_panel(np);
  Workspace ws(workspace, workspace_bytes);
  const size_t nn = (size_t)np * np;
  double* G = ws.take<double>(nn);   // scaled b -> U (upper triangle)
  double* Y = ws.take<double>(nn);   // scaled a -> U^-T a'
  double* Yt = ws.take<double>(nn);  // its transpose -> S (padded)
  double* S = ws.take<double>((size_t)n * n);
  PanelScratch ps;
  ps.diag = ws.take<double>((size_t)B * B);
  ps.tmp = ws.take<double>((size_t)B * B);
  ps.dk = ws.take<double>((size_t)B * kCholBlock);
  ps.panel = ws.take<double>((size_t)B * np);
  ps.panel_neg = ws.take<double>((size_t)B * np);
  ps.cinv_all = ws.take<double>((size_t)np * B);
  ps.cinvt_all = ws.take<double>((size_t)np * B);
  double* dinv = ws.take<double>(np);
  double* z = ws.take<double>(np);
  double* y = ws.take<double>(np);
  double* v = ws.take<double>(B);
  int* fail = ws.take<int>(16);
  if (!G || !Y || !Yt || !S || !ps.diag || !ps.tmp || !ps.dk || !ps.panel || !ps.panel_neg || !ps.cinv_all ||
      !ps.cinvt_all || !dinv || !z || !y || !v || !fail) {
    set_error("tnpy_geig_chol_lowest: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              tnpy_geig_chol_workspace_bytes(n));
    return TNPY_EWORKSPACE;
  }
  TNPY_CUDA_OK(cudaMemsetAsync(fail, 0, 16 * sizeof(int), stream));
  pad_scale_kernel<<<grid_for((int64_t)nn), 256, 0, stream>>>(b, n, n, G, np, nullptr);
  TNPY_LAUNCH_OK();
  TNPY_TRY(spd_scale_pad(G, n, np, dinv, fail, stream));  // G <- D^-1 b D^-1, identity on the padding
  TNPY_TRY(cholesky_upper_blocked(G, np, B, ps, fail, stream));
  int failed = 0;
  TNPY_CUDA_OK(cudaMemcpyAsync(&failed, fail, sizeof(int), cudaMemcpyDeviceToHost, stream));
  TNPY_CUDA_OK(cudaStreamSynchronize(stream));
  if (failed) {
    set_error("tnpy_geig_chol_lowest: the right-hand matrix is not positive definite to working precision (n = %d)", n);
    return TNPY_ENOCONV;
  }
  // S = U^-T a' U^-1 = U^-T (U^-T a')^T, a' = D^-1 a D^-1 (zero on the padding)
  pad_scale_kernel<<<grid_for((int64_t)nn), 256, 0, stream>>>(a, n, n, Y, np, dinv);
  TNPY_LAUNCH_OK();
  TNPY_TRY(forward_substitute(G, np, B, ps, Y, np, np, stream));
  TNPY_TRY(transpose(Y, np, np, np, Yt, np, nullptr, stream));
  TNPY_TRY(forward_substitute(G, np, B, ps, Yt, np, np, stream));
  block_copy_kernel<<<grid_for((int64_t)n * n), 256, 0, stream>>>(Yt, np, S, n, nullptr, 0, n, n);
  TNPY_LAUNCH_OK();
  start_vector_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, n);
  TNPY_LAUNCH_OK();
  double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  char* rest = static_cast<char*>(workspace) + ws.used;
  const int rc = tnpy_eig_lowest(S, device_one(), device_one(), z, n, 1, 1, 1, 1, 0, tol, max_matvec, 0, stats, rest,
                                 workspace_bytes - ws.used, stream_);
  if (stats_host)
    for (int i = 0; i < 8; ++i) stats_host[i] = stats[i];
  if (rc != TNPY_OK && rc != TNPY_ENOCONV) return rc;
  // x = D^-1 U^-1 z (z zero on the padding)
  if (np > n) TNPY_CUDA_OK(cudaMemsetAsync(z + n, 0, sizeof(double) * (np - n), stream));
  TNPY_TRY(backward_substitute(G, np, B, ps, z, y, v, stream));
  TNPY_CUDA_OK(cudaMemcpyAsync(x, y, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
  scale_by_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(x, dinv, n);
  TNPY_LAUNCH_OK();
  TNPY_CUDA_OK(cudaMemcpyAsync(theta_dev, &stats[0], sizeof(double), cudaMemcpyHostToDevice, stream));
  TNPY_CUDA_OK(cudaStreamSynchronize(stream));
  return rc;
}
