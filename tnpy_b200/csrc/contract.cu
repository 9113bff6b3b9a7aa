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
#include <new>

#include "heff.cuh"

namespace tnpy {

int axpy(double alpha, const double* a_dev, const double* x, double* y, int64_t n, cudaStream_t stream);

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

// May this call use the int8 tensor-core path?  AUTO defers to the process-wide selection (tnpy_set_gemm_algo /
// TNPY_GEMM_ALGO); the default selection and TNPY_GEMM_OZAKI allow it, GENERIC / DMMA / FP64 mean native FP64.
static bool tcgen05_allowed(int algo) {
  if (algo == TNPY_GEMM_AUTO) algo = current_gemm_algo();
  return algo == TNPY_GEMM_AUTO || algo == TNPY_GEMM_OZAKI;
}

size_t chain_gemm_bytes(int M, int N, int K) {
  if (!ozaki_applicable(M, N, K)) return Workspace::need(gemm_splitk_doubles(M, N, K));  // FP64: split-K partials
  return oz_operand_bytes(M, K) + oz_operand_bytes(N, K) + oz_mma_scratch_bytes(M, N);
}

// FP64 GEMM with split-K partials taken from a copy of the workspace (released again on return)
static int fp64_gemm(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
                     int accumulate, Workspace ws, cudaStream_t stream) {
  const size_t want = gemm_splitk_doubles(M, N, K);
  double* scratch = want ? ws.take<double>(want) : nullptr;
  return gemm_tn_ws(A, lda, B, ldb, out, M, N, K, accumulate, TNPY_GEMM_FP64, scratch, scratch ? want : 0, stream);
}

int chain_gemm(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
               int accumulate, int algo, Workspace& ws, cudaStream_t stream) {
  if (tcgen05_allowed(algo) && ozaki_applicable(M, N, K)) {
    Workspace probe = ws;  // all or nothing: fall back to FP64 when the slices do not fit
    OzOperand a, b;
    if (oz_operand_take(probe, M, K, &a) && oz_operand_take(probe, N, K, &b)) {
      TNPY_TRY(oz_slice_operand(A, lda, oz_plain_rows(K), a, stream));
      TNPY_TRY(oz_slice_operand(B, ldb, oz_plain_rows(K), b, stream));
      return oz_mma(a, b, out, M, N, ozaki_slices(), accumulate, probe, nullptr, stream);
    }
  }
  return fp64_gemm(A, lda, B, ldb, out, M, N, K, accumulate, ws, stream);
}

// ---- prepared H_eff ---------------------------------------------------------------------------
// lo = number of output (bra) rows of L held by the caller, starting at row0: L is (l, wl, lo) = L_full[:, :, row0 :
// row0 + lo], y is (lo, d, r).  lo == l is the ordinary matvec; lo < l is one rank's row block of the chi-sharded one
// (the left identity channel is then L[li, 0, m] = delta(li, row0 + m)).
//
// flags (canonical-gauge shortcuts, SURVEY 7 "identity channels"): with the upper-triangular MPOs of
// tnpy.model and a mixed-canonical MPS, L[:, 0, :] and R[:, wr-1, :] are identity matrices.
//   TNPY_LEFT_IDENTITY : T1[p,r,0,m] = x[m,p,r] is a transpose instead of 1/wl of the first GEMM;
//   TNPY_RIGHT_IDENTITY: the b = wr-1 slice of the second GEMM is y[m,q,s] += T2[s,wr-1,q,m]; the
//                        intermediate is laid out MPO-bond-slowest so the GEMM just drops that K range.
// The caller vouches for the flags (tnpy_identity_defect measures them); executed flops drop by up to
// 2/w while the algorithmic flop count F_mv is unchanged.
struct ChainShape {
  bool left_id, right_id;
  int m1, n1, k1;  // GEMM 1: T1 = x^T L      (M = d r, N = channels * lo, K = l)
  int m3, n3, k3;  // GEMM 3: y  = T2^T R     (M = d lo, N = r, K = r * channels)
};
static ChainShape chain_shape(int l, int lo, int r, int wl, int wr, int d, int flags) {
  ChainShape s;
  s.left_id = (flags & TNPY_LEFT_IDENTITY) && wl > 1;
  s.right_id = (flags & TNPY_RIGHT_IDENTITY) && wr > 1;
  s.m1 = d * r; s.n1 = (s.left_id ? wl - 1 : wl) * lo; s.k1 = l;
  s.m3 = d * lo; s.n3 = r; s.k3 = r * (s.right_id ? wr - 1 : wr);
  return s;
}

static bool direct_shapes_ok(int l, int lo, int r, int wl, int wr, int d) {
  return wl > 1 && wr > 1 && oz_premix_applicable(l, r, wl, wr, d) && ozaki_applicable(lo * d, r, (wr - 1) * r) &&
         ozaki_applicable(lo, d * r, (wl - 1) * l);
}

// no non-zero block W[a, b] with a > 0 and b < wr - 1
static bool no_interior_blocks(const double* W_host, int wl, int wr, int d) {
  for (int a = 1; a < wl; ++a)
    for (int b = 0; b + 1 < wr; ++b)
      for (int e = 0; e < d * d; ++e)
        if (W_host[((size_t)a * wr + b) * d * d + e] != 0.0) return false;
  return true;
}

// K ranges of the direct path's GEMMs that multiply identically zero operand pieces: the piece (channel c, block q) of a
// premixed operand vanishes when column q of that channel's d x d MPO block does (an S+ or S- block has one non-zero
// element: half of its pieces).  channel_block(c) = pointer to the (d x d) block of channel c; a channel spans
// `per_channel` K entries; blocks of tiles of one q are `block_tiles` tiles wide; nonzero(c, q) says whether piece (c, q)
// can be non-zero.  Leaves mode 0 (no skipping) when the geometry does not line up with the 64-entry K chunks / the
// tile grid, or nothing can be skipped.
template <typename NonZero>
static void plan_kskip(OzKSkip* out, int mode, int channels, int d, int per_channel, int block_tiles, NonZero nonzero) {
  *out = OzKSkip{};
  if (per_channel % 64 != 0 || block_tiles <= 0 || d > kOzSkipBlocks) return;
  OzKSkip s{};
  s.mode = mode; s.block_tiles = block_tiles; s.nblocks = d;
  const int chunks = per_channel / 64;
  bool anything = false;
  for (int q = 0; q < d; ++q) {
    int n = 0;
    bool open = false;
    for (int c = 0; c < channels; ++c) {
      if (nonzero(c, q)) {
        if (!open) {
          if (n == kOzSkipRanges) return;
          s.lo[q][n] = c * chunks;
          open = true;
        }
        s.hi[q][n] = (c + 1) * chunks;
      } else {
        anything = true;
        if (open) { ++n; open = false; }
      }
    }
    if (open) ++n;
    if (n == 0) return;  // a block with nothing to multiply: leave it to the plain kernel
    s.nranges[q] = n;
  }
  if (anything) *out = s;
}

size_t heff_plan_bytes(int l, int lo, int r, int wl, int wr, int d) {
  size_t chain = Workspace::need((size_t)r * wr * r);  // r2 of the FP64 chain
  if (ozaki_applicable(d * r, wl * lo, l)) chain += oz_operand_bytes(wl * lo, l);
  if (ozaki_applicable(d * lo, r, r * wr)) chain += oz_operand_bytes(r, r * wr);
  size_t direct = 0;
  if (direct_shapes_ok(l, lo, r, wl, wr, d)) direct = oz_operand_bytes(lo, (wl - 1) * l) + oz_operand_bytes(r, (wr - 1) * r);
  return (chain > direct ? chain : direct) + Workspace::need(1) + 1024;
}

size_t heff_apply_bytes(int l, int lo, int r, int wl, int wr, int d) {
  size_t chain = Workspace::need((size_t)d * r * wl * lo) + Workspace::need((size_t)r * wr * d * lo);  // T1, T2
  if (ozaki_applicable(d * r, wl * lo, l)) chain += oz_operand_bytes(d * r, l) + oz_mma_scratch_bytes(d * r, wl * lo);
  if (ozaki_applicable(d * lo, r, r * wr)) chain += oz_operand_bytes(d * lo, r * wr) + oz_mma_scratch_bytes(d * lo, r);
  {  // split-K partials of the FP64 GEMMs (one at a time)
    const size_t a = gemm_splitk_doubles(d * r, wl * lo, l), b = gemm_splitk_doubles(d * lo, r, r * wr);
    chain += Workspace::need(a > b ? a : b);
  }
  size_t direct = 0;
  if (direct_shapes_ok(l, lo, r, wl, wr, d))
    direct = oz_operand_bytes(lo * d, (wr - 1) * r) + oz_operand_bytes(d * r, (wl - 1) * l) +
             oz_mma_scratch_bytes(lo * d, r) + oz_mma_scratch_bytes(lo, d * r);
  return (chain > direct ? chain : direct) + 1024;
}

int heff_plan_init(HeffPlan* plan, const double* L, const double* W, const double* R, const double* W_host, int l,
                   int lo, int row0, int r, int wl, int wr, int d, int flags, int algo, Workspace& mem,
                   cudaStream_t stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(lo > 0 && row0 >= 0 && row0 + lo <= l, "row block outside the left bond");
  TNPY_CHECK_ARG(W != nullptr, "null pointer");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1 && lo == 1), "L may be NULL only for unit left bond");
  TNPY_CHECK_ARG(R || (r == 1 && wr == 1), "R may be NULL only for unit right bond");
  if (!L) L = device_one();
  if (!R) R = device_one();
  *plan = HeffPlan{};
  plan->L = L; plan->W = W; plan->R = R;
  plan->l = l; plan->lo = lo; plan->row0 = row0; plan->r = r; plan->wl = wl; plan->wr = wr; plan->d = d; plan->flags = flags;
  plan->mode = HEFF_FP64_CHAIN;
  const ChainShape s = chain_shape(l, lo, r, wl, wr, d, flags);
  const bool oz_ok = tcgen05_allowed(algo);
  if (oz_ok) {
    plan->bound = mem.take<double>(1);
    if (!plan->bound) {
      set_error("heff_plan_init: plan memory too small");
      return TNPY_EWORKSPACE;
    }
    TNPY_CUDA_OK(cudaMemsetAsync(plan->bound, 0, sizeof(double), stream));
  }
  if (oz_ok && s.left_id && s.right_id && direct_shapes_ok(l, lo, r, wl, wr, d)) {
    double w_local[kPmMaxCh * kPmMaxCh * kPmMaxD * kPmMaxD];
    if (!W_host && (size_t)wl * wr * d * d <= sizeof(w_local) / sizeof(double)) {
      TNPY_CUDA_OK(cudaMemcpyAsync(w_local, W, sizeof(double) * (size_t)wl * wr * d * d, cudaMemcpyDeviceToHost, stream));
      TNPY_CUDA_OK(cudaStreamSynchronize(stream));
      W_host = w_local;
    }
    if (W_host && no_interior_blocks(W_host, wl, wr, d)) {
      Workspace probe = mem;
      if (oz_operand_take(probe, lo, (wl - 1) * l, &plan->envL) && oz_operand_take(probe, r, (wr - 1) * r, &plan->envR)) {
        mem = probe;
        // L'[(a - 1) l + li, m] = L[li, a, m]  (a >= 1);   R'[b r + ri, s] = R[ri, b, s]  (b < wr - 1)
        TNPY_TRY(oz_slice_operand(L, lo, OzRowMap{l, wl, 1}, plan->envL, stream));
        TNPY_TRY(oz_slice_operand(R, r, OzRowMap{r, wr, 0}, plan->envR, stream));
        // R-side GEMM: rows (q, m), K = (b, ri), b < wr - 1, piece (b, q) from W[0, b]; blocks of lo / 256 m-tiles
        auto column_nonzero = [&](const double* blk, int q) {
          for (int p = 0; p < d; ++p)
            if (blk[p * d + q] != 0.0) return true;
          return false;
        };
        plan_kskip(&plan->skipR, 1, wr - 1, d, r, lo % 256 == 0 ? lo / 256 : 0,
                   [&](int b, int q) { return column_nonzero(W_host + (size_t)b * d * d, q); });
        // L-side GEMM: columns (q, s), K = (a, li), a >= 1, piece (a, q) from W[a, wr - 1]; blocks of r / 128 n-tiles
        plan_kskip(&plan->skipL, 2, wl - 1, d, l, r % 128 == 0 ? r / 128 : 0,
                   [&](int a, int q) { return column_nonzero(W_host + ((size_t)(a + 1) * wr + (wr - 1)) * d * d, q); });
        plan->mode = HEFF_OZ_DIRECT;
        return TNPY_OK;
      }
    }
  }
  plan->g1_oz = oz_ok && ozaki_applicable(s.m1, s.n1, s.k1);
  plan->g3_oz = oz_ok && ozaki_applicable(s.m3, s.n3, s.k3);
  if (plan->g1_oz) {
    Workspace probe = mem;
    if (oz_operand_take(probe, s.n1, s.k1, &plan->envL)) {
      mem = probe;
      TNPY_TRY(oz_slice_operand(L + (s.left_id ? lo : 0), (int64_t)wl * lo, oz_plain_rows(l), plan->envL, stream));
    } else {
      plan->g1_oz = false;
    }
  }
  if (plan->g3_oz) {
    Workspace probe = mem;
    if (oz_operand_take(probe, s.n3, s.k3, &plan->envR)) {
      mem = probe;
      // on the tcgen05 path K always runs channel-major, (b, ri) (with the last channel dropped under the right
      // identity flag), matching the channel-major intermediate: a channel is then a contiguous K range, and the
      // GEMM can skip the ranges whose piece (b, q) of the mixed intermediate vanishes -- W[a, b, p, q] = 0 for all
      // a, p (XXZ: 2 of 10 pieces)
      TNPY_TRY(oz_slice_operand(R, r, OzRowMap{r, wr, 0}, plan->envR, stream));
      double w_local[kPmMaxCh * kPmMaxCh * kPmMaxD * kPmMaxD];
      if (!W_host && (size_t)wl * wr * d * d <= sizeof(w_local) / sizeof(double)) {
        TNPY_CUDA_OK(cudaMemcpyAsync(w_local, W, sizeof(double) * (size_t)wl * wr * d * d, cudaMemcpyDeviceToHost, stream));
        TNPY_CUDA_OK(cudaStreamSynchronize(stream));
        W_host = w_local;
      }
      if (W_host)
        plan_kskip(&plan->skip3, 1, s.k3 / r, d, r, lo % 256 == 0 ? lo / 256 : 0, [&](int b, int q) {
          for (int a = 0; a < wl; ++a)
            for (int pp = 0; pp < d; ++pp)
              if (W_host[(((size_t)a * wr + b) * d + pp) * d + q] != 0.0) return true;
          return false;
        });
    } else {
      plan->g3_oz = false;
    }
  }
  if (plan->g1_oz || plan->g3_oz) plan->mode = HEFF_OZ_CHAIN;
  if (s.right_id && !plan->g3_oz) {
    plan->r2 = mem.take<double>((size_t)r * (wr - 1) * r);
    if (!plan->r2) {
      set_error("heff_plan_init: plan memory too small");
      return TNPY_EWORKSPACE;
    }
    channel_major_kernel<<<sm_count() * 8, 256, 0, stream>>>(R, plan->r2, r, wr, wr - 1, r);
    TNPY_LAUNCH_OK();
  }
  return TNPY_OK;
}

static int apply_direct(const HeffPlan& p, const double* x, double* y, int S, const double* shift_dev, Workspace& ws,
                        cudaStream_t stream) {
  const int l = p.l, lo = p.lo, r = p.r, wl = p.wl, wr = p.wr, d = p.d;
  const double* x_rows = x + (int64_t)p.row0 * d * r;  // the caller's rows of x: all the R-side term needs
  OzOperand xa, xb;
  if (!oz_operand_take(ws, lo * d, (wr - 1) * r, &xa) || !oz_operand_take(ws, d * r, (wl - 1) * l, &xb)) {
    set_error("heff_apply: workspace too small");
    return TNPY_EWORKSPACE;
  }
  // y = W[0, wr-1] x - shift x, then both GEMMs accumulate into it
  TNPY_TRY(oz_premix_a(x_rows, p.W, lo, r, wl, wr, d, xa, y, shift_dev, p.skipR.mode != 0, stream));
  TNPY_TRY(oz_premix_b(x, p.W, l, r, wl, wr, d, xb, p.skipL.mode != 0, stream));
  // y[m, q, s] += sum_{(b ri)} Xa[(b ri), (q m)] R'[(b ri), s]: GEMM rows (q, m) -> y rows (m, q) by the split-M row map
  TNPY_TRY(oz_mma(xa, p.envR, GemmOut{y, (int64_t)d * r, (int64_t)r, lo}, lo * d, r, S, 1, ws, p.bound, stream, &p.skipR));
  // y[m, (q s)] += sum_{(a li)} L'[(a li), m] Xb[(a li), (q s)]
  TNPY_TRY(oz_mma(p.envL, xb, plain_out(y, (int64_t)d * r, lo), lo, d * r, S, 1, ws, p.bound, stream, &p.skipL));
  return TNPY_OK;
}

int heff_plan_apply(const HeffPlan& p, const double* x, double* y, int S, const double* shift_dev, Workspace& ws,
                    cudaStream_t stream) {
  TNPY_CHECK_ARG(x && y, "null pointer");
  if (S <= 0) S = ozaki_slices();
  if (p.mode == HEFF_OZ_DIRECT) return apply_direct(p, x, y, S, shift_dev, ws, stream);
  const int l = p.l, lo = p.lo, r = p.r, wl = p.wl, wr = p.wr, d = p.d;
  const ChainShape s = chain_shape(l, lo, r, wl, wr, d, p.flags);
  double* t1 = ws.take<double>((size_t)d * r * wl * lo);
  double* t2 = ws.take<double>((size_t)r * wr * d * lo);
  if (!t1 || !t2) {
    set_error("heff_apply: workspace too small");
    return TNPY_EWORKSPACE;
  }
  // T1[p, r, a, m] = sum_l x[l, (p r)] L[l, (a m)]
  double* t1_dst = t1 + (s.left_id ? lo : 0);
  if (s.left_id)  // a = 0: T1[p, r, 0, m] = x[row0 + m, p, r]
    TNPY_TRY(transpose_strided(x + (int64_t)p.row0 * d * r, (int64_t)d * r, 0, lo, d * r, t1, (int64_t)wl * lo, 0, 1, stream));
  if (p.g1_oz) {
    Workspace scratch = ws;
    OzOperand xs;
    if (!oz_operand_take(scratch, s.m1, s.k1, &xs)) {
      set_error("heff_apply: workspace too small");
      return TNPY_EWORKSPACE;
    }
    TNPY_TRY(oz_slice_operand(x, (int64_t)d * r, oz_plain_rows(l), xs, stream));
    TNPY_TRY(oz_mma(xs, p.envL, plain_out(t1_dst, (int64_t)wl * lo, s.m1), s.m1, s.n1, S, 0, scratch, p.bound, stream));
  } else {
    TNPY_TRY(fp64_gemm(x, (int64_t)d * r, p.L + (s.left_id ? lo : 0), (int64_t)wl * lo,
                       plain_out(t1_dst, (int64_t)wl * lo, s.m1), s.m1, s.n1, s.k1, 0, ws, stream));
  }
  // T2[r, b, q, m] (or [b, r, q, m]) = sum_{a p} W[a, b, p, q] T1[p, r, a, m]      (u=p, u'=q, v=a, v'=b)
  TNPY_TRY(wmix(t1, t2, p.W, d, d, wl, wr, r, lo, d, 1, wr * d * d, d * d, stream, (s.right_id || p.g3_oz) ? 1 : 0));
  // y[m, q, s] = sum_{r b} T2[(r b), (q m)] R[(r b), s]               rows (q m) -> (m q)
  // (right identity: the GEMM runs over the channels b < wr-1 and the identity-channel term T2[wr-1, s, q, m]
  // is added by a transposing pass afterwards so that the GEMM epilogue stays store-only)
  GemmOut out{y, (int64_t)d * r, (int64_t)r, lo};
  if (p.g3_oz) {
    Workspace scratch = ws;
    OzOperand ts;
    if (!oz_operand_take(scratch, s.m3, s.k3, &ts)) {
      set_error("heff_apply: workspace too small");
      return TNPY_EWORKSPACE;
    }
    TNPY_TRY(oz_slice_operand(t2, (int64_t)d * lo, oz_plain_rows(s.k3), ts, stream));
    TNPY_TRY(oz_mma(ts, p.envR, out, s.m3, s.n3, S, 0, scratch, p.bound, stream, &p.skip3));
  } else {
    TNPY_TRY(fp64_gemm(t2, (int64_t)d * lo, s.right_id ? p.r2 : p.R, (int64_t)r, out, s.m3, s.n3, s.k3, 0, ws, stream));
  }
  if (s.right_id) {
    const double* t2_last = t2 + (size_t)(wr - 1) * r * d * lo;
    TNPY_TRY(transpose_strided(t2_last, (int64_t)d * lo, lo, r, lo, y, (int64_t)d * r, r, d, stream, true));
  }
  if (shift_dev) TNPY_TRY(axpy(-1.0, shift_dev, x + (int64_t)p.row0 * d * r, y, (int64_t)lo * d * r, stream));
  return TNPY_OK;
}

int heff_apply_rows(const double* L, const double* W, const double* R, const double* x, double* y, int l, int lo,
                    int row0, int r, int wl, int wr, int d, int flags, Workspace& ws, cudaStream_t stream) {
  HeffPlan plan;
  TNPY_TRY(heff_plan_init(&plan, L, W, R, nullptr, l, lo, row0, r, wl, wr, d, flags, TNPY_GEMM_AUTO, ws, stream));
  return heff_plan_apply(plan, x, y, 0, nullptr, ws, stream);
}

int heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l, int r, int wl,
               int wr, int d, int flags, Workspace& ws, cudaStream_t stream) {
  return heff_apply_rows(L, W, R, x, y, l, l, 0, r, wl, wr, d, flags, ws, stream);
}

// lo bra rows of L starting at row0 (lo == l, row0 == 0: the whole update).  With a row block the result is this
// block's *contribution* to Lout (the sum over the bra index m runs over the block only): the row-sharded sweep sums
// the contributions of the ranks (one all-reduce) -- A is needed in full (ket index), L only by rows.
int env_update_left_rows(const double* L, const double* A, const double* W, double* Lout, int l, int lo, int row0, int r,
                         int wl, int wr, int d, int flags, Workspace& ws, cudaStream_t stream) {
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(A && W && Lout, "null pointer");
  TNPY_CHECK_ARG(lo > 0 && row0 >= 0 && row0 + lo <= l, "row block outside the left bond");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1), "L may be NULL only for unit left bond");
  if (!L) L = device_one();
  double* t1 = ws.take<double>((size_t)wl * lo * d * r);
  double* t2 = ws.take<double>((size_t)lo * d * wr * r);
  if (!t1 || !t2) {
    set_error("env_update_left: workspace too small");
    return TNPY_EWORKSPACE;
  }
  const int algo = TNPY_GEMM_AUTO;
  const double* A_rows = A + (int64_t)row0 * d * r;
  // T1[a, m, p, r] = sum_l L[l, (a m)] A[l, (p r)];  identity channel a = 0: T1[0, m] = A[row0 + m]
  if ((flags & TNPY_LEFT_IDENTITY) && wl > 1) {
    TNPY_CUDA_OK(cudaMemcpyAsync(t1, A_rows, sizeof(double) * (size_t)lo * d * r, cudaMemcpyDeviceToDevice, stream));
    Workspace scratch = ws;
    TNPY_TRY(chain_gemm(L + lo, (int64_t)wl * lo, A, (int64_t)d * r, plain_out(t1 + (size_t)lo * d * r, (int64_t)d * r, (wl - 1) * lo),
                        (wl - 1) * lo, d * r, l, 0, algo, scratch, stream));
  } else {
    Workspace scratch = ws;
    TNPY_TRY(chain_gemm(L, (int64_t)wl * lo, A, (int64_t)d * r, plain_out(t1, (int64_t)d * r, wl * lo), wl * lo, d * r, l, 0,
                        algo, scratch, stream));
  }
  // T2[m, q, b, r] = sum_{a p} W[a, b, p, q] T1[a, m, p, r]          (u=a, u'=b, v=p, v'=q)
  TNPY_TRY(wmix(t1, t2, W, wl, wr, d, d, lo, r, wr * d * d, d * d, d, 1, stream));
  // Lout[r, b, s] = sum_{m q} T2[(m q), (b r)] A[(m q), s]             rows (b r) -> (r b)
  GemmOut out{Lout, (int64_t)wr * r, (int64_t)r, r};
  Workspace scratch = ws;
  TNPY_TRY(chain_gemm(t2, (int64_t)wr * r, A_rows, (int64_t)r, out, wr * r, r, lo * d, 0, algo, scratch, stream));
  return TNPY_OK;
}

int env_update_left(const double* L, const double* A, const double* W, double* Lout, int l, int r, int wl, int wr,
                    int d, int flags, Workspace& ws, cudaStream_t stream) {
  return env_update_left_rows(L, A, W, Lout, l, l, 0, r, wl, wr, d, flags, ws, stream);
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
    Workspace scratch = ws;
    TNPY_TRY(chain_gemm(R, (int64_t)wr * r, at, (int64_t)d * l, plain_out(t1, (int64_t)d * l, (wr - 1) * r), (wr - 1) * r,
                        d * l, r, 0, algo, scratch, stream));
  } else {
    Workspace scratch = ws;
    TNPY_TRY(chain_gemm(R, (int64_t)wr * r, at, (int64_t)d * l, plain_out(t1, (int64_t)d * l, wr * r), wr * r, d * l, r, 0,
                        algo, scratch, stream));
  }
  // T2[s, q, a, l] = sum_{b p} W[a, b, p, q] T1[b, s, p, l]          (u=b, u'=a, v=p, v'=q)
  TNPY_TRY(wmix(t1, t2, W, wr, wl, d, d, r, l, d * d, wr * d * d, d, 1, stream));
  // Rout[l, a, m] = sum_{s q} T2[(s q), (a l)] At[(s q), m]            rows (a l) -> (l a)
  GemmOut out{Rout, (int64_t)wl * l, (int64_t)l, l};
  Workspace scratch = ws;
  TNPY_TRY(chain_gemm(t2, (int64_t)wl * l, at, (int64_t)l, out, wl * l, l, r * d, 0, algo, scratch, stream));
  return TNPY_OK;
}

// dense H[(l p r), (m q s)] = sum_{a b} L[l,a,m] W[a,b,p,q] R[r,b,s]   (tiny sites, and the dense pencils of
// ShiftInvertDMRG up to N = 32768: 8.6 GB)
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

// The same matrix in two passes for anything but tiny sites: LW[p, q, l, b, m] = sum_a L[l, a, m] W[a, b, p, q]
// (d^2 l^2 wr numbers), then H[(l p r), (m q s)] = sum_b LW[p, q, l, b, m] R[r, b, s] -- wr multiply-adds per entry
// instead of wl * wr, the R reads coalesced along s and the LW reads broadcast.  (The squared MPO of ShiftInvertDMRG has
// wl = wr = 25 .. 36 and pencils of 10^4 unknowns: 10 s per matrix with the one-pass kernel above, milliseconds here.)
__global__ void __launch_bounds__(256) heff_dense_lw_kernel(const double* __restrict__ L, const double* __restrict__ W,
                                                            double* __restrict__ LW, int l, int wl, int wr, int d) {
  const int64_t total = (int64_t)d * d * l * wr * l;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(e % l);
    int64_t t = e / l;
    const int b = (int)(t % wr);
    t /= wr;
    const int li = (int)(t % l);
    t /= l;
    const int q = (int)(t % d), p = (int)(t / d);
    double acc = 0.0;
    for (int a = 0; a < wl; ++a) acc = fma(L[((int64_t)li * wl + a) * l + m], W[((a * wr + b) * d + p) * d + q], acc);
    LW[e] = acc;
  }
}

__global__ void __launch_bounds__(256) heff_dense_from_lw_kernel(const double* __restrict__ LW, const double* __restrict__ R,
                                                                 double* __restrict__ H, int l, int r, int wr, int d) {
  const int n = l * d * r;
  const int64_t total = (int64_t)n * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / n), col = (int)(e % n);
    const int li = row / (d * r), p = (row / r) % d, ri = row % r;
    const int m = col / (d * r), q = (col / r) % d, s = col % r;
    const double* lw = LW + ((((int64_t)p * d + q) * l + li) * wr) * l + m;
    const double* rr = R + (int64_t)ri * wr * r + s;
    double acc = 0.0;
    for (int b = 0; b < wr; ++b) acc = fma(lw[(int64_t)b * l], rr[(int64_t)b * r], acc);
    H[e] = acc;
  }
}

}  // namespace tnpy

using namespace tnpy;

static size_t max3(size_t a, size_t b, size_t c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

extern "C" size_t tnpy_heff_workspace_bytes(int l, int r, int wl, int wr, int d) {
  return heff_plan_bytes(l, l, r, wl, wr, d) + heff_apply_bytes(l, l, r, wl, wr, d) + 1024;
}
extern "C" size_t tnpy_env_workspace_bytes(int l, int r, int wl, int wr, int d) {
  const size_t wmax = (size_t)(wl > wr ? wl : wr);
  const size_t left = max3(chain_gemm_bytes(wl * l, d * r, l), chain_gemm_bytes(wr * r, r, l * d), 0);
  const size_t right = max3(chain_gemm_bytes(wr * r, d * l, r), chain_gemm_bytes(wl * l, l, r * d), 0);
  return 3 * Workspace::need((size_t)l * r * d * wmax) + (left > right ? left : right) + 1024;
}
extern "C" size_t tnpy_heff_dense_workspace_bytes(int l, int /*r*/, int /*wl*/, int wr, int d) {
  return Workspace::need((size_t)d * d * l * wr * l) + 256;
}

extern "C" int tnpy_heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l,
                               int r, int wl, int wr, int d, int flags, void* workspace, size_t workspace_bytes,
                               void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return heff_apply(L, W, R, x, y, l, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
}

// ---- prepared H_eff: the host-side handle is a small struct owned by the library, all device memory is the caller's
struct tnpy_heff_plan {
  HeffPlan plan;
};

extern "C" size_t tnpy_heff_plan_bytes(int l, int r, int wl, int wr, int d) { return heff_plan_bytes(l, l, r, wl, wr, d) + 256; }

static int plan_create(tnpy_heff_plan** handle, const double* L, const double* W, const double* R, const double* W_host,
                       int l, int lo, int row0, int r, int wl, int wr, int d, int flags, int algo, void* plan_memory,
                       size_t plan_bytes, void* stream) {
  TNPY_CHECK_ARG(handle != nullptr, "null handle");
  TNPY_CHECK_ARG(algo >= TNPY_GEMM_AUTO && algo <= TNPY_GEMM_FP64, "unknown algo");
  tnpy_heff_plan* h = new (std::nothrow) tnpy_heff_plan;
  if (!h) {
    set_error("tnpy_heff_plan_create: out of host memory");
    return TNPY_EINVAL;
  }
  Workspace mem(plan_memory, plan_bytes);
  const int rc = heff_plan_init(&h->plan, L, W, R, W_host, l, lo, row0, r, wl, wr, d, flags, algo, mem,
                                static_cast<cudaStream_t>(stream));
  if (rc != TNPY_OK) {
    delete h;
    return rc;
  }
  *handle = h;
  return TNPY_OK;
}

extern "C" int tnpy_heff_plan_create(tnpy_heff_plan** handle, const double* L, const double* W, const double* R,
                                     const double* W_host, int l, int r, int wl, int wr, int d, int flags, int algo,
                                     void* plan_memory, size_t plan_bytes, void* stream) {
  return plan_create(handle, L, W, R, W_host, l, l, 0, r, wl, wr, d, flags, algo, plan_memory, plan_bytes, stream);
}

extern "C" int tnpy_heff_plan_create_rows(tnpy_heff_plan** handle, const double* L_rows, const double* W, const double* R,
                                          const double* W_host, int l, int row0, int l_rows, int r, int wl, int wr, int d,
                                          int flags, int algo, void* plan_memory, size_t plan_bytes, void* stream) {
  return plan_create(handle, L_rows, W, R, W_host, l, l_rows, row0, r, wl, wr, d, flags, algo, plan_memory, plan_bytes, stream);
}

extern "C" int tnpy_heff_plan_mode(const tnpy_heff_plan* handle) { return handle ? handle->plan.mode : TNPY_EINVAL; }

extern "C" int tnpy_heff_plan_apply(const tnpy_heff_plan* handle, const double* x, double* y, int slices,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  TNPY_CHECK_ARG(handle != nullptr, "null handle");
  TNPY_CHECK_ARG(slices == 0 || (slices >= 5 && slices <= kOzMaxSlices), "slices must be 0 (default) or 5 .. 8");
  Workspace ws(workspace, workspace_bytes);
  return heff_plan_apply(handle->plan, x, y, slices, nullptr, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_heff_plan_error_bound(const tnpy_heff_plan* handle, double* bound_dev_out, void* stream) {
  TNPY_CHECK_ARG(handle != nullptr && bound_dev_out != nullptr, "null pointer");
  if (handle->plan.bound)
    TNPY_CUDA_OK(cudaMemcpyAsync(bound_dev_out, handle->plan.bound, sizeof(double), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
  else
    TNPY_CUDA_OK(cudaMemsetAsync(bound_dev_out, 0, sizeof(double), static_cast<cudaStream_t>(stream)));
  return TNPY_OK;
}

extern "C" int tnpy_heff_plan_destroy(tnpy_heff_plan* handle) {
  delete handle;
  return TNPY_OK;
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
                                    double* y_rows, int l, int row0, int l_rows, int r, int wl, int wr, int d, int flags,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return heff_apply_rows(L_rows, W, R, x, y_rows, l, l_rows, row0, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_left(const double* L, const double* A, const double* W, double* Lout, int l, int r,
                                    int wl, int wr, int d, int flags, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_left(L, A, W, Lout, l, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_left_rows(const double* L_rows, const double* A, const double* W, double* Lout_partial, int l,
                                         int row0, int l_rows, int r, int wl, int wr, int d, int flags, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_left_rows(L_rows, A, W, Lout_partial, l, l_rows, row0, r, wl, wr, d, flags, ws,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_env_update_right(const double* R, const double* A, const double* W, double* Rout, int l, int r,
                                     int wl, int wr, int d, int flags, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  Workspace ws(workspace, workspace_bytes);
  return env_update_right(R, A, W, Rout, l, r, wl, wr, d, flags, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int tnpy_heff_dense(const double* L, const double* W, const double* R, double* H, int l, int r, int wl,
                               int wr, int d, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_TRY(check_dims(l, r, wl, wr, d));
  TNPY_CHECK_ARG(W && H, "null pointer");
  TNPY_CHECK_ARG(L || (l == 1 && wl == 1), "L may be NULL only for unit left bond");
  TNPY_CHECK_ARG(R || (r == 1 && wr == 1), "R may be NULL only for unit right bond");
  if (!L) L = device_one();
  if (!R) R = device_one();
  const int64_t n = (int64_t)l * d * r;
  TNPY_CHECK_ARG(n <= 32768, "dense H_eff limited to N <= 32768");
  const int blocks = (int)((n * n + 255) / 256 < 4096 ? (n * n + 255) / 256 : 4096);
  Workspace ws(workspace, workspace_bytes);
  double* LW = n >= 64 ? ws.take<double>((size_t)d * d * l * wr * l) : nullptr;
  if (LW) {  // two passes (a caller without a workspace, and tiny sites, get the one-pass kernel)
    const int64_t lw_total = (int64_t)d * d * l * wr * l;
    heff_dense_lw_kernel<<<(int)((lw_total + 255) / 256 < 4096 ? (lw_total + 255) / 256 : 4096), 256, 0, stream>>>(L, W, LW, l, wl, wr, d);
    TNPY_LAUNCH_OK();
    heff_dense_from_lw_kernel<<<blocks, 256, 0, stream>>>(LW, R, H, l, r, wr, d);
    TNPY_LAUNCH_OK();
    return TNPY_OK;
  }
  heff_dense_kernel<<<blocks, 256, 0, stream>>>(L, W, R, H, l, r, wl, wr, d);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

extern "C" int tnpy_mirror_lpr(const double* in, double* out, int l, int d, int r, void* stream) {
  TNPY_CHECK_ARG(in && out && l > 0 && d > 0 && r > 0, "bad argument");
  return mirror(in, out, l, d, r, static_cast<cudaStream_t>(stream));
}
