// Contraction chains of the fDMRG local update, built from the TN GEMM plus two bandwidth kernels.
//
//   H_eff.psi   (reference matrix_product_state.py:411-440)
//   update_left / update_right (reference matrix_product_state.py:296-336)
//   dense H_eff for tiny sites (reference matrix_product_state.py:372-409)
//
// Every chain is   GEMM (contract the environment bond)  ->  W-mix (apply the MPO tensor)  ->
// GEMM (contract the other bond + MPO bond).  The intermediate layouts are chosen so that
//   * the contracted index is always the slowest index of both GEMM operands (one TN kernel),
//   * the W-mix reads and writes with the long bond index fastest (fully coalesced, any w / d),
//   * the second GEMM's output rows are scattered straight into the reference layout through the
//     kernel's split-M row map (no transpose pass).
#include "common.cuh"

namespace tnpy {

// ---------------------------------------------------------------------------------------------
// W-mix:  out[X][v'][u'][Y] = sum_{u,v} Wc(u,u',v,v') * in[u][X][v][Y]
//         Wc(u,u',v,v') = W[u*su + u'*sup + v*sv + v'*svp]
// One output channel (v',u') per blockIdx.y; the block first compacts the non-zero coefficients of
// its channel (MPO tensors are sparse: XXZ has 14 non-zeros out of 100) in a fixed order, then
// streams the (X,Y) plane.
// ---------------------------------------------------------------------------------------------
constexpr int kMixMaxTerms = 288;

template <int VEC>
__global__ void __launch_bounds__(256) wmix_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                   const double* __restrict__ W, int nu, int nup, int nv, int nvp,
                                                   int X, int Y, int su, int sup, int sv, int svp) {
  __shared__ double coef[kMixMaxTerms];
  __shared__ int64_t inoff[kMixMaxTerms];
  __shared__ int nterms;
  const int vp = blockIdx.y / nup, up = blockIdx.y % nup;
  const int nuv = nu * nv;
  if (threadIdx.x < 32) {
    // warp-ballot stream compaction keeps the term order (u, v) ascending => deterministic sums
    int count = 0;
    for (int base = 0; base < nuv; base += 32) {
      const int idx = base + threadIdx.x;
      double c = 0.0;
      if (idx < nuv) {
        const int u = idx / nv, v = idx % nv;
        c = W[u * su + up * sup + v * sv + vp * svp];
      }
      const unsigned mask = __ballot_sync(0xffffffffu, c != 0.0);
      if (c != 0.0) {
        const int pos = count + __popc(mask & ((1u << threadIdx.x) - 1));
        const int u = idx / nv, v = idx % nv;
        coef[pos] = c;
        inoff[pos] = ((int64_t)u * X * nv + v) * Y;
      }
      count += __popc(mask);
    }
    if (threadIdx.x == 0) nterms = count;
  }
  __syncthreads();
  const int nt = nterms;
  const int64_t plane = (int64_t)X * Y / VEC;  // elements (or element pairs) of the (X,Y) plane
  const int yv = Y / VEC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = e / yv;
    const int y = (int)(e - x * yv) * VEC;
    const int64_t ibase = x * (int64_t)nv * Y + y;
    const int64_t obase = ((x * nvp + vp) * nup + up) * (int64_t)Y + y;
    if (VEC == 2) {
      double2 acc = make_double2(0.0, 0.0);
      for (int k = 0; k < nt; ++k) {
        const double2 v = *reinterpret_cast<const double2*>(in + inoff[k] + ibase);
        acc.x = fma(coef[k], v.x, acc.x);
        acc.y = fma(coef[k], v.y, acc.y);
      }
      *reinterpret_cast<double2*>(out + obase) = acc;
    } else {
      double acc = 0.0;
      for (int k = 0; k < nt; ++k) acc = fma(coef[k], in[inoff[k] + ibase], acc);
      out[obase] = acc;
    }
  }
}

static int wmix(const double* in, double* out, const double* W, int nu, int nup, int nv, int nvp, int X, int Y, int su,
                int sup, int sv, int svp, cudaStream_t stream) {
  if (nu * nv > kMixMaxTerms) {
    set_error("wmix: MPO bond x physical dimension %d exceeds the compiled limit %d", nu * nv, kMixMaxTerms);
    return TNPY_EINVAL;
  }
  const bool vec = (Y % 2 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  const int64_t plane = (int64_t)X * Y / (vec ? 2 : 1);
  const int channels = nvp * nup;
  int64_t bx = (plane + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = max((int64_t)1, (int64_t)sm_count() * 16 / channels);
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)channels);
  if (vec)
    wmix_kernel<2><<<grid, 256, 0, stream>>>(in, out, W, nu, nup, nv, nvp, X, Y, su, sup, sv, svp);
  else
    wmix_kernel<1><<<grid, 256, 0, stream>>>(in, out, W, nu, nup, nv, nvp, X, Y, su, sup, sv, svp);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// ---------------------------------------------------------------------------------------------
// mirror: out[r][p][l] = in[l][p][r]   (32x32 shared-memory tile transpose per physical index)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mirror_kernel(const double* __restrict__ in, double* __restrict__ out, int l,
                                                     int d, int r) {
  __shared__ double tile[32][33];
  const int p = blockIdx.z;
  const int r0 = blockIdx.x * 32, l0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int i = ty; i < 32; i += 8) {
    const int li = l0 + i, ri = r0 + tx;
    if (li < l && ri < r) tile[i][tx] = in[((int64_t)li * d + p) * r + ri];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int ri = r0 + i, li = l0 + tx;
    if (li < l && ri < r) out[((int64_t)ri * d + p) * l + li] = tile[tx][i];
  }
}

static int mirror(const double* in, double* out, int l, int d, int r, cudaStream_t stream) {
  dim3 grid(ceil_div(r, 32), ceil_div(l, 32), d);
  mirror_kernel<<<grid, 256, 0, stream>>>(in, out, l, d, r);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// ---------------------------------------------------------------------------------------------
// chains
// ---------------------------------------------------------------------------------------------
static int check_dims(int l, int r, int wl, int wr, int d) {
  TNPY_CHECK_ARG(l > 0 && r > 0 && wl > 0 && wr > 0 && d > 0, "non-positive dimension");
  return TNPY_OK;
}

// lo = number of output (bra) rows of L held by the caller: L is (l, wl, lo), y is (lo, d, r).
// lo == l is the ordinary matvec; lo < l is one rank's row block of the chi-sharded matvec.
int heff_apply_rows(const double* L, const double* W, const double* R, const double* x, double* y, int l, int lo,
                    int r, int wl, int wr, int d, Workspace& ws, cudaStream_t stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(lo > 0, "non-positive row count");
  TNPY_CHECK_ARG(W && x && y, "null pointer");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1 && lo == 1), "L may be NULL only for unit left bond");
  TNPY_CHECK_ARG(R || (r == 1 && wr == 1), "R may be NULL only for unit right bond");
  if (!L) L = device_one();
  if (!R) R = device_one();
  double* t1 = ws.take<double>((size_t)d * r * wl * lo);
  double* t2 = ws.take<double>((size_t)r * wr * d * lo);
  if (!t1 || !t2) {
    set_error("heff_apply: workspace too small");
    return TNPY_EWORKSPACE;
  }
  const int algo = TNPY_GEMM_AUTO;
  // T1[p, r, a, m] = sum_l x[l, (p r)] L[l, (a m)]
  TNPY_TRY(gemm_tn(x, (int64_t)d * r, L, (int64_t)wl * lo, plain_out(t1, (int64_t)wl * lo, d * r), d * r, wl * lo, l,
                   0, algo, stream));
  // T2[r, b, q, m] = sum_{a p} W[a, b, p, q] T1[p, r, a, m]          (u=p, u'=q, v=a, v'=b)
  TNPY_TRY(wmix(t1, t2, W, d, d, wl, wr, r, lo, d, 1, wr * d * d, d * d, stream));
  // y[m, q, s] = sum_{r b} T2[(r b), (q m)] R[(r b), s]               rows (q m) -> (m q)
  GemmOut out{y, (int64_t)d * r, (int64_t)r, lo};
  TNPY_TRY(gemm_tn(t2, (int64_t)d * lo, R, (int64_t)r, out, d * lo, r, r * wr, 0, algo, stream));
  return TNPY_OK;
}

int heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l, int r, int wl,
               int wr, int d, Workspace& ws, cudaStream_t stream) {
  return heff_apply_rows(L, W, R, x, y, l, l, r, wl, wr, d, ws, stream);
}

int env_update_left(const double* L, const double* A, const double* W, double* Lout, int l, int r, int wl, int wr,
                    int d, Workspace& ws, cudaStream_t stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(A && W && Lout, "null pointer");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1), "L may be NULL only for unit left bond");
  if (!L) L = device_one();
  double* t1 = ws.take<double>((size_t)wl * l * d * r);
  double* t2 = ws.take<double>((size_t)l * d * wr * r);
  if (!t1 || !t2) {
    set_error("env_update_left: workspace too small");
    return TNPY_EWORKSPACE;
  }
  const int algo = TNPY_GEMM_AUTO;
  // T1[a, m, p, r] = sum_l L[l, (a m)] A[l, (p r)]
  TNPY_TRY(gemm_tn(L, (int64_t)wl * l, A, (int64_t)d * r, plain_out(t1, (int64_t)d * r, wl * l), wl * l, d * r, l, 0,
                   algo, stream));
  // T2[m, q, b, r] = sum_{a p} W[a, b, p, q] T1[a, m, p, r]          (u=a, u'=b, v=p, v'=q)
  TNPY_TRY(wmix(t1, t2, W, wl, wr, d, d, l, r, wr * d * d, d * d, d, 1, stream));
  // Lout[r, b, s] = sum_{m q} T2[(m q), (b r)] A[(m q), s]             rows (b r) -> (r b)
  GemmOut out{Lout, (int64_t)wr * r, (int64_t)r, r};
  TNPY_TRY(gemm_tn(t2, (int64_t)wr * r, A, (int64_t)r, out, wr * r, r, l * d, 0, algo, stream));
  return TNPY_OK;
}

int env_update_right(const double* R, const double* A, const double* W, double* Rout, int l, int r, int wl, int wr,
                     int d, Workspace& ws, cudaStream_t stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(A && W && Rout, "null pointer");
  TNPY_CHECK_ARG(R || (r == 1 && wr == 1), "R may be NULL only for unit right bond");
  if (!R) R = device_one();
  double* at = ws.take<double>((size_t)r * d * l);
  double* t1 = ws.take<double>((size_t)wr * r * d * l);
  double* t2 = ws.take<double>((size_t)r * d * wl * l);
  if (!at || !t1 || !t2) {
    set_error("env_update_right: workspace too small");
    return TNPY_EWORKSPACE;
  }
  const int algo = TNPY_GEMM_AUTO;
  // At[r, p, l] = A[l, p, r]  -- mirror image of the site tensor; the right update is the left
  // update of the mirrored chain with the MPO bond roles swapped.
  TNPY_TRY(mirror(A, at, l, d, r, stream));
  // T1[b, s, p, l] = sum_r R[r, (b s)] At[r, (p l)]
  TNPY_TRY(gemm_tn(R, (int64_t)wr * r, at, (int64_t)d * l, plain_out(t1, (int64_t)d * l, wr * r), wr * r, d * l, r, 0,
                   algo, stream));
  // T2[s, q, a, l] = sum_{b p} W[a, b, p, q] T1[b, s, p, l]          (u=b, u'=a, v=p, v'=q)
  TNPY_TRY(wmix(t1, t2, W, wr, wl, d, d, r, l, d * d, wr * d * d, d, 1, stream));
  // Rout[l, a, m] = sum_{s q} T2[(s q), (a l)] At[(s q), m]            rows (a l) -> (l a)
  GemmOut out{Rout, (int64_t)wl * l, (int64_t)l, l};
  TNPY_TRY(gemm_tn(t2, (int64_t)wl * l, at, (int64_t)l, out, wl * l, l, r * d, 0, algo, stream));
  return TNPY_OK;
}

// dense H[(l p r), (m q s)] = sum_{a b} L[l,a,m] W[a,b,p,q] R[r,b,s]   (N = l d r < ~200)
__global__ void __launch_bounds__(256) heff_dense_kernel(const double* __restrict__ L, const double* __restrict__ W,
                                                         const double* __restrict__ R, double* __restrict__ H, int l,
                                                         int r, int wl, int wr, int d) {
  const int n = l * d * r;
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / n), col = (int)(e % n);
    const int li = row / (d * r), p = (row / r) % d, ri = row % r;
    const int m = col / (d * r), q = (col / r) % d, s = col % r;
    double acc = 0.0;
    for (int a = 0; a < wl; ++a) {
      const double lv = L[((int64_t)li * wl + a) * l + m];
      if (lv == 0.0) continue;
      double inner = 0.0;
      for (int b = 0; b < wr; ++b)
        inner = fma(W[((a * wr + b) * d + p) * d + q], R[((int64_t)ri * wr + b) * r + s], inner);
      acc = fma(lv, inner, acc);
    }
    H[e] = acc;
  }
}

}  // namespace tnpy

using namespace tnpy;

static size_t chain_ws(int l, int r, int wl, int wr, int d) {
  const size_t wmax = (size_t)(wl > wr ? wl : wr);
  return 3 * Workspace::need((size_t)l * r * d * wmax) + 1024;
}

extern "C" size_t tnpy_heff_workspace_bytes(int l, int r, int wl, int wr, int d) { return chain_ws(l, r, wl, wr, d); }
extern "C" size_t tnpy_env_workspace_bytes(int l, int r, int wl, int wr, int d) { return chain_ws(l, r, wl, wr, d); }
extern "C" size_t tnpy_heff_dense_workspace_bytes(int, int, int, int, int) { return 256; }

extern "C" int tnpy_heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l,
                               int r, int wl, int wr, int d, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return heff_apply(L, W, R, x, y, l, r, wl, wr, d, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_heff_apply_rows(const double* L_rows, const double* W, const double* R, const double* x,
                                    double* y_rows, int l, int l_rows, int r, int wl, int wr, int d, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return heff_apply_rows(L_rows, W, R, x, y_rows, l, l_rows, r, wl, wr, d, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_left(const double* L, const double* A, const double* W, double* Lout, int l, int r,
                                    int wl, int wr, int d, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_left(L, A, W, Lout, l, r, wl, wr, d, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_right(const double* R, const double* A, const double* W, double* Rout, int l, int r,
                                     int wl, int wr, int d, void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_right(R, A, W, Rout, l, r, wl, wr, d, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_heff_dense(const double* L, const double* W, const double* R, double* H, int l, int r, int wl,
                               int wr, int d, void* /*workspace*/, size_t /*workspace_bytes*/, void* stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(W && H, "null pointer");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1), "L may be NULL only for unit left bond");
  TNPY_CHECK_ARG(R || (r == 1 && wr == 1), "R may be NULL only for unit right bond");
  if (!L) L = device_one();
  if (!R) R = device_one();
  const int64_t n = (int64_t)l * d * r;
  TNPY_CHECK_ARG(n <= 4096, "dense H_eff limited to N <= 4096");
  const int blocks = (int)((n * n + 255) / 256 < 4096 ? (n * n + 255) / 256 : 4096);
  heff_dense_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(L, W, R, H, l, r, wl, wr, d);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

extern "C" int tnpy_mirror_lpr(const double* in, double* out, int l, int d, int r, void* stream) {
  TNPY_CHECK_ARG(in && out && l > 0 && d > 0 && r > 0, "bad argument");
  return mirror(in, out, l, d, r, static_cast<cudaStream_t>(stream));
}
