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
                                                   int X, int Y, int su, int sup, int sv, int svp, int channel_major) {
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
    // default out[X][v'][u'][Y]; channel_major: out[v'][X][u'][Y] (MPO bond slowest, so a GEMM can drop
    // whole channels from its K range)
    const int64_t obase = (channel_major ? (((int64_t)vp * X + x) * nup + up) : ((x * nvp + vp) * nup + up)) * (int64_t)Y + y;
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
                int sup, int sv, int svp, cudaStream_t stream, int channel_major = 0) {
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
    wmix_kernel<2><<<grid, 256, 0, stream>>>(in, out, W, nu, nup, nv, nvp, X, Y, su, sup, sv, svp, channel_major);
  else
    wmix_kernel<1><<<grid, 256, 0, stream>>>(in, out, W, nu, nup, nv, nvp, X, Y, su, sup, sv, svp, channel_major);
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

// out[z][c][r] = in[z][r][c]: batched strided transpose (32x32 shared-memory tiles).
// in rows have stride ld_in, out rows stride ld_out; batch z advances by in_z / out_z elements.
template <bool ADD>
__global__ void __launch_bounds__(256) transpose_strided_kernel(const double* __restrict__ in, int64_t ld_in,
                                                                int64_t in_z, int rows, int cols,
                                                                double* __restrict__ out, int64_t ld_out,
                                                                int64_t out_z) {
  __shared__ double tile[32][33];
  in += blockIdx.z * in_z;
  out += blockIdx.z * out_z;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i][tx] = in[(int64_t)r * ld_in + c];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) {
      double* dst = out + (int64_t)c * ld_out + r;
      *dst = ADD ? *dst + tile[tx][i] : tile[tx][i];
    }
  }
}

static int transpose_strided(const double* in, int64_t ld_in, int64_t in_z, int rows, int cols, double* out,
                             int64_t ld_out, int64_t out_z, int batch, cudaStream_t stream, bool add = false) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch);
  if (add)
    transpose_strided_kernel<true><<<grid, 256, 0, stream>>>(in, ld_in, in_z, rows, cols, out, ld_out, out_z);
  else
    transpose_strided_kernel<false><<<grid, 256, 0, stream>>>(in, ld_in, in_z, rows, cols, out, ld_out, out_z);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// R2[b][r][s] = R[r][b][s] for b < w_keep  (MPO bond slowest, trailing channels dropped)
__global__ void __launch_bounds__(256) channel_major_kernel(const double* __restrict__ R, double* __restrict__ R2,
                                                            int r, int w, int w_keep, int s_dim) {
  const int64_t total = (int64_t)r * w_keep * s_dim;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(e % s_dim);
    const int64_t br = e / s_dim;
    const int ri = (int)(br % r), b = (int)(br / r);
    R2[e] = R[((int64_t)ri * w + b) * s_dim + s];
  }
}

// max_ij |E[i][c][j] - delta_ij| for an environment E (dim, w, dim): is channel c the identity?
__global__ void __launch_bounds__(256) identity_defect_kernel(const double* __restrict__ E, int dim, int w, int c,
                                                              double* __restrict__ partial) {
  __shared__ double sh[32];
  double worst = 0.0;
  const int64_t total = (int64_t)dim * dim;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / dim), j = (int)(e % dim);
    worst = fmax(worst, fabs(E[((int64_t)i * w + c) * dim + j] - (i == j ? 1.0 : 0.0)));
  }
  // max-reduce (values are >= 0, so a sum-free shuffle max)
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = worst;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}
__global__ void max_reduce_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) v = fmax(v, partial[i]);
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (threadIdx.x == 0) *out = v;
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
//
// flags (canonical-gauge shortcuts, SURVEY 7 "identity channels"): with the upper-triangular MPOs of
// tnpy.model and a mixed-canonical MPS, L[:, 0, :] and R[:, wr-1, :] are identity matrices.
//   TNPY_LEFT_IDENTITY : T1[p,r,0,m] = x[m,p,r] is a transpose instead of 1/wl of the first GEMM;
//   TNPY_RIGHT_IDENTITY: the b = wr-1 slice of the second GEMM is y[m,q,s] += T2[s,wr-1,q,m]; the
//                        intermediate is laid out MPO-bond-slowest so the GEMM just drops that K range.
// The caller vouches for the flags (tnpy_identity_defect measures them); executed flops drop by up to
// 2/w while the algorithmic flop count F_mv is unchanged.
int heff_apply_rows(const double* L, const double* W, const double* R, const double* x, double* y, int l, int lo,
                    int r, int wl, int wr, int d, int flags, Workspace& ws, cudaStream_t stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(lo > 0, "non-positive row count");
  TNPY_CHECK_ARG(W && x && y, "null pointer");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1 && lo == 1), "L may be NULL only for unit left bond");
  TNPY_CHECK_ARG(R || (r == 1 && wr == 1), "R may be NULL only for unit right bond");
  if (!L) L = device_one();
  if (!R) R = device_one();
  const bool left_id = (flags & TNPY_LEFT_IDENTITY) && wl > 1 && lo == l;
  const bool right_id = (flags & TNPY_RIGHT_IDENTITY) && wr > 1;
  double* t1 = ws.take<double>((size_t)d * r * wl * lo);
  double* t2 = ws.take<double>((size_t)r * wr * d * lo);
  double* r2 = right_id ? ws.take<double>((size_t)r * (wr - 1) * r) : nullptr;
  if (!t1 || !t2 || (right_id && !r2)) {
    set_error("heff_apply: workspace too small");
    return TNPY_EWORKSPACE;
  }
  const int algo = TNPY_GEMM_AUTO;
  // T1[p, r, a, m] = sum_l x[l, (p r)] L[l, (a m)]
  if (left_id) {
    TNPY_TRY(transpose_strided(x, (int64_t)d * r, 0, l, d * r, t1, (int64_t)wl * lo, 0, 1, stream));  // a = 0
    TNPY_TRY(gemm_tn(x, (int64_t)d * r, L + lo, (int64_t)wl * lo, plain_out(t1 + lo, (int64_t)wl * lo, d * r), d * r,
                     (wl - 1) * lo, l, 0, algo, stream));
  } else {
    TNPY_TRY(gemm_tn(x, (int64_t)d * r, L, (int64_t)wl * lo, plain_out(t1, (int64_t)wl * lo, d * r), d * r, wl * lo, l,
                     0, algo, stream));
  }
  // T2[r, b, q, m] (or [b, r, q, m]) = sum_{a p} W[a, b, p, q] T1[p, r, a, m]      (u=p, u'=q, v=a, v'=b)
  TNPY_TRY(wmix(t1, t2, W, d, d, wl, wr, r, lo, d, 1, wr * d * d, d * d, stream, right_id ? 1 : 0));
  // y[m, q, s] = sum_{r b} T2[(r b), (q m)] R[(r b), s]               rows (q m) -> (m q)
  GemmOut out{y, (int64_t)d * r, (int64_t)r, lo};
  if (right_id) {
    // y[m, q, s] = (GEMM over the channels b < wr-1) + T2[wr-1, s, q, m]; the identity-channel term is
    // added by a transposing pass afterwards so that the GEMM epilogue stays store-only
    channel_major_kernel<<<sm_count() * 8, 256, 0, stream>>>(R, r2, r, wr, wr - 1, r);
    TNPY_LAUNCH_OK();
    TNPY_TRY(gemm_tn(t2, (int64_t)d * lo, r2, (int64_t)r, out, d * lo, r, r * (wr - 1), 0, algo, stream));
    const double* t2_last = t2 + (size_t)(wr - 1) * r * d * lo;
    TNPY_TRY(transpose_strided(t2_last, (int64_t)d * lo, lo, r, lo, y, (int64_t)d * r, r, d, stream, true));
  } else {
    TNPY_TRY(gemm_tn(t2, (int64_t)d * lo, R, (int64_t)r, out, d * lo, r, r * wr, 0, algo, stream));
  }
  return TNPY_OK;
}

int heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l, int r, int wl,
               int wr, int d, int flags, Workspace& ws, cudaStream_t stream) {
  return heff_apply_rows(L, W, R, x, y, l, l, r, wl, wr, d, flags, ws, stream);
}

int env_update_left(const double* L, const double* A, const double* W, double* Lout, int l, int r, int wl, int wr,
                    int d, int flags, Workspace& ws, cudaStream_t stream) {
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
  // T1[a, m, p, r] = sum_l L[l, (a m)] A[l, (p r)];  identity channel a = 0: T1[0] = A
  if ((flags & TNPY_LEFT_IDENTITY) && wl > 1) {
    TNPY_CUDA_OK(cudaMemcpyAsync(t1, A, sizeof(double) * (size_t)l * d * r, cudaMemcpyDeviceToDevice, stream));
    TNPY_TRY(gemm_tn(L + l, (int64_t)wl * l, A, (int64_t)d * r, plain_out(t1 + (size_t)l * d * r, (int64_t)d * r, (wl - 1) * l),
                     (wl - 1) * l, d * r, l, 0, algo, stream));
  } else {
    TNPY_TRY(gemm_tn(L, (int64_t)wl * l, A, (int64_t)d * r, plain_out(t1, (int64_t)d * r, wl * l), wl * l, d * r, l, 0,
                     algo, stream));
  }
  // T2[m, q, b, r] = sum_{a p} W[a, b, p, q] T1[a, m, p, r]          (u=a, u'=b, v=p, v'=q)
  TNPY_TRY(wmix(t1, t2, W, wl, wr, d, d, l, r, wr * d * d, d * d, d, 1, stream));
  // Lout[r, b, s] = sum_{m q} T2[(m q), (b r)] A[(m q), s]             rows (b r) -> (r b)
  GemmOut out{Lout, (int64_t)wr * r, (int64_t)r, r};
  TNPY_TRY(gemm_tn(t2, (int64_t)wr * r, A, (int64_t)r, out, wr * r, r, l * d, 0, algo, stream));
  return TNPY_OK;
}

int env_update_right(const double* R, const double* A, const double* W, double* Rout, int l, int r, int wl, int wr,
                     int d, int flags, Workspace& ws, cudaStream_t stream) {
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
  // T1[b, s, p, l] = sum_r R[r, (b s)] At[r, (p l)];  identity channel b = wr-1: T1[wr-1] = At
  if ((flags & TNPY_RIGHT_IDENTITY) && wr > 1) {
    TNPY_CUDA_OK(cudaMemcpyAsync(t1 + (size_t)(wr - 1) * r * d * l, at, sizeof(double) * (size_t)r * d * l,
                                 cudaMemcpyDeviceToDevice, stream));
    TNPY_TRY(gemm_tn(R, (int64_t)wr * r, at, (int64_t)d * l, plain_out(t1, (int64_t)d * l, (wr - 1) * r), (wr - 1) * r,
                     d * l, r, 0, algo, stream));
  } else {
    TNPY_TRY(gemm_tn(R, (int64_t)wr * r, at, (int64_t)d * l, plain_out(t1, (int64_t)d * l, wr * r), wr * r, d * l, r, 0,
                     algo, stream));
  }
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
  return 3 * Workspace::need((size_t)l * r * d * wmax) + Workspace::need((size_t)r * wr * r) + 1024;
}

extern "C" size_t tnpy_heff_workspace_bytes(int l, int r, int wl, int wr, int d) { return chain_ws(l, r, wl, wr, d); }
extern "C" size_t tnpy_env_workspace_bytes(int l, int r, int wl, int wr, int d) { return chain_ws(l, r, wl, wr, d); }
extern "C" size_t tnpy_heff_dense_workspace_bytes(int, int, int, int, int) { return 256; }

extern "C" int tnpy_heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l,
                               int r, int wl, int wr, int d, int flags, void* workspace, size_t workspace_bytes,
                               void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return heff_apply(L, W, R, x, y, l, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_identity_defect(const double* E, int dim, int w, int channel, double* defect_dev, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(E && defect_dev && dim > 0 && w > 0 && channel >= 0 && channel < w, "bad argument");
  Workspace ws(workspace, workspace_bytes);
  const int blocks = sm_count() * 4;
  double* partial = ws.take<double>(blocks);
  if (!partial) {
    set_error("tnpy_identity_defect: workspace too small (need %zu bytes)", Workspace::need(blocks) + 256);
    return TNPY_EWORKSPACE;
  }
  identity_defect_kernel<<<blocks, 256, 0, stream>>>(E, dim, w, channel, partial);
  TNPY_LAUNCH_OK();
  max_reduce_kernel<<<1, 32, 0, stream>>>(partial, blocks, defect_dev);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

extern "C" int tnpy_heff_apply_rows(const double* L_rows, const double* W, const double* R, const double* x,
                                    double* y_rows, int l, int l_rows, int r, int wl, int wr, int d, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return heff_apply_rows(L_rows, W, R, x, y_rows, l, l_rows, r, wl, wr, d, 0, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_left(const double* L, const double* A, const double* W, double* Lout, int l, int r,
                                    int wl, int wr, int d, int flags, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_left(L, A, W, Lout, l, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_right(const double* R, const double* A, const double* W, double* Rout, int l, int r,
                                     int wl, int wr, int d, int flags, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_right(R, A, W, Rout, l, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
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
