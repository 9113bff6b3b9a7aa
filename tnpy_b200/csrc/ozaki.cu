// FP64-accurate GEMMs on the 5th-gen tensor cores (the default path of the large contractions).
// tcgen05 has no f64 kind, so C = A^T B is evaluated with the Ozaki scheme: every operand column is scaled by a
// power of two and split error-free into signed 7-bit slices,
//     A[k][m] = 2^ea[m] * sum_i a_i[k][m] * 2^(-7(i+1)),    |a_i| <= 64,
// all slice products a_i^T b_j are exact int8 x int8 -> int32 GEMMs (tcgen05.mma.kind::i8, accumulators in TMEM),
// slice pairs of equal weight i + j = d share one TMEM accumulator (exact: < 2^31 for K <= 65536), pairs with
// i + j >= S are below the target precision and skipped, and the epilogue recombines the S accumulators in FP64:
//     C[m][n] = 2^(ea[m]+eb[n]) * sum_d acc_d[m][n] * 2^(-7(d+2)).
//
// Error of one product with S slices (rigorous, worst case): every scaled element is its S slices plus a remainder
// |rho| <= 2^(-7S-1); the skipped pairs and the remainders add up to at most e_S = (S+2)/4 * 2^(-7S) per element
// product, so  |C - A^T B|[m][n] <= K e_S sa[m] sb[n]  and  ||C - A^T B||_F <= K e_S ||sa||_2 ||sb||_2
// (sa, sb the column scales; e_8 = 3.5e-17, e_7 = 4.0e-15, e_6 = 4.5e-13).  oz_mma folds that bound into a device
// scalar the eigensolver reads with its status record (typical errors are sqrt(K) x random-sign smaller).
//
// Nothing here allocates: operands, scales and the tail scratch are carved from the caller's workspace.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "ozaki.cuh"

namespace tnpy {

constexpr int kOzBM = 128;  // tile rows per CTA (TMEM lanes)
constexpr int kOzBN = 64;   // B rows staged per CTA (half of the pair's 128 columns)
constexpr int kOzBK = 64;   // k elements (= bytes) per stage row: one 64B swizzle span

// ---------------------------------------------------------------------------------------------
// slicing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t oz_src_row(int k, OzRowMap rows) {
  return (int64_t)(k % rows.kin) * rows.kmul + k / rows.kin + rows.koff;
}

// colmax[c] = max_k |P[row(k)][c]| as the bit pattern of a non-negative double (integer order == value
// order), reduced over k-chunks with atomicMax; colmax must be zeroed first.
__global__ void __launch_bounds__(256) oz_colmax_kernel(const double* __restrict__ P, int64_t ld, OzRowMap rows, int K,
                                                        int MN, int k_chunk, unsigned long long* __restrict__ colmax) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= MN) return;
  const int k0 = blockIdx.y * k_chunk, k1 = min(K, k0 + k_chunk);
  double mx = 0.0;
  for (int k = k0; k < k1; ++k) mx = fmax(mx, fabs(P[oz_src_row(k, rows) * ld + c]));
  if (mx > 0.0) atomicMax(&colmax[c], (unsigned long long)__double_as_longlong(mx));
}

// scale = 2^e with |v| / 2^e <= 0.5 for every |v| <= the bound whose bit pattern is `bits` (0 for a zero column)
__device__ __forceinline__ double oz_scale_of(unsigned long long bits) {
  const double mx = __longlong_as_double((long long)bits);
  return mx > 0.0 ? ldexp(1.0, ilogb(mx) + 2) : 0.0;
}

// digit = rint(128 t) without conversion instructions: adding 1.5 * 2^52 leaves the rounded integer in the low
// mantissa bits (two's complement), subtracting it again gives the rounded value; the remainder is exact.
// Returns the digit (|digit| <= 64) in the low byte of an int.
__device__ __forceinline__ int oz_next_digit(double& t) {
  const double magic = 6755399441055744.0;
  const double u = fma(t, 128.0, magic);
  t = fma(t, 128.0, magic - u);  // exact remainder, |t| <= 0.5
  return __double2loint(u);
}

// The slicing kernels all work on 16 consecutive K entries of one column per thread: 16 scaled values in,
// kOzMaxSlices x 16 digit bytes out, one 16-byte store per slice (byte e of slice sl = digit sl of value e).
// Digits of four values are packed with three byte permutes per slice.
template <typename ValueOf>
__device__ __forceinline__ void oz_digits16(ValueOf&& value_of, uint4 (&out)[kOzMaxSlices]) {
  uint32_t w[kOzMaxSlices][4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    double t0 = value_of(4 * g), t1 = value_of(4 * g + 1), t2 = value_of(4 * g + 2), t3 = value_of(4 * g + 3);
#pragma unroll
    for (int sl = 0; sl < kOzMaxSlices; ++sl) {
      const int d0 = oz_next_digit(t0), d1 = oz_next_digit(t1), d2 = oz_next_digit(t2), d3 = oz_next_digit(t3);
      w[sl][g] = __byte_perm(__byte_perm(d0, d1, 0x0040), __byte_perm(d2, d3, 0x0040), 0x5410);
    }
  }
#pragma unroll
  for (int sl = 0; sl < kOzMaxSlices; ++sl) out[sl] = make_uint4(w[sl][0], w[sl][1], w[sl][2], w[sl][3]);
}

// store the 16 digit bytes of every slice at row + k0 (k0 .. k0 + 15 < Kp); 16-byte stores when aligned
__device__ __forceinline__ void oz_store16(int8_t* __restrict__ row, int64_t slice_stride, const uint4 (&dig)[kOzMaxSlices],
                                           bool aligned) {
  if (aligned) {
#pragma unroll
    for (int sl = 0; sl < kOzMaxSlices; ++sl) *reinterpret_cast<uint4*>(row + sl * slice_stride) = dig[sl];
  } else {
#pragma unroll
    for (int sl = 0; sl < kOzMaxSlices; ++sl) {
      const uint32_t w[4] = {dig[sl].x, dig[sl].y, dig[sl].z, dig[sl].w};
#pragma unroll
      for (int e = 0; e < 16; ++e) row[sl * slice_stride + e] = (int8_t)(w[e >> 2] >> (8 * (e & 3)));
    }
  }
}

// The transposing slicers (source contiguous along the columns, output contiguous along K) stage the digits of a
// 32-column x 128-k tile in shared memory -- 16 bytes per thread and slice -- and write whole 128-byte lines; 16-byte
// stores straight from the registers, one row per thread, were measured 2x slower (partial-sector writes).
constexpr int kSliceTileLd = 144;  // 128 k + 16 pad: the 16-byte chunks of a quarter warp fall into distinct banks
typedef int8_t OzSliceTile[kOzMaxSlices][32][kSliceTileLd];

__device__ __forceinline__ void oz_tile_put(OzSliceTile& tile, int tx, int ty, const uint4 (&dig)[kOzMaxSlices]) {
#pragma unroll
  for (int sl = 0; sl < kOzMaxSlices; ++sl) *reinterpret_cast<uint4*>(&tile[sl][tx][16 * ty]) = dig[sl];
}

// write the tile's columns c0 .. c0 + 31 (< ncols), bytes [kbase, min(kbase + 128, kend)) of their rows
__device__ __forceinline__ void oz_tile_flush(const OzSliceTile& tile, int8_t* __restrict__ slices, int64_t slice_stride,
                                              int64_t Kp, int64_t c0, int64_t ncols, int64_t kbase, int64_t kend) {
  const bool aligned = (kbase & 15) == 0;
  for (int idx = threadIdx.x; idx < kOzMaxSlices * 32 * 8; idx += 256) {
    const int chunk = idx % 8, cc = (idx / 8) % 32, sl = idx / 256;
    const int64_t k = kbase + 16 * chunk;
    if (c0 + cc >= ncols || k >= kend) continue;
    int8_t* dst = slices + sl * slice_stride + (c0 + cc) * Kp + k;
    if (aligned && k + 16 <= kend) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&tile[sl][cc][16 * chunk]);
    } else {
      for (int e = 0; e < 16 && k + e < kend; ++e) dst[e] = tile[sl][cc][16 * chunk + e];
    }
  }
}

// slices[s][c][k] (k contiguous, Kp bytes per row) from P[row(k)][c].  Block = 32 columns x 8 chunks of 16 k: the
// loads are coalesced along the columns, every thread then owns 16 consecutive bytes of one (slice, column) row.
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ P, int64_t ld, OzRowMap rows, int K,
                                                       int MN, const unsigned long long* __restrict__ colmax,
                                                       double* __restrict__ scale, double* __restrict__ sumsq,
                                                       int8_t* __restrict__ slices, int64_t Kp, int64_t slice_stride) {
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = blockIdx.x * 32 + tx;
  const int k0 = (blockIdx.y * 8 + ty) * 16;
  const double sc = c < MN ? oz_scale_of(colmax[c]) : 0.0;
  const double inv = sc > 0.0 ? 1.0 / sc : 0.0;  // exact: power of two
  if (blockIdx.y == 0 && ty == 0) {
    if (c < MN) scale[c] = sc;
    const double part = warp_sum(sc * sc);
    if (tx == 0 && part > 0.0) atomicAdd(sumsq, part);  // feeds an error *bound*: summation order is immaterial
  }
  __shared__ __align__(16) OzSliceTile tile;
  double t[16];
  // source row of k0 by one division, of the following 15 entries by counting (the gather wraps at multiples of kin)
  int rem = k0 % rows.kin, quo = k0 / rows.kin;
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const int64_t row = (int64_t)rem * rows.kmul + quo + rows.koff;
    t[e] = (c < MN && k0 + e < K) ? P[row * ld + c] * inv : 0.0;  // |t| <= 0.5
    if (++rem == rows.kin) { rem = 0; ++quo; }
  }
  uint4 dig[kOzMaxSlices];
  oz_digits16([&t](int e) { return t[e]; }, dig);
  oz_tile_put(tile, tx, ty, dig);
  __syncthreads();
  oz_tile_flush(tile, slices, slice_stride, Kp, (int64_t)blockIdx.x * 32, MN, (int64_t)blockIdx.y * 128, Kp);
}

// ---------------------------------------------------------------------------------------------
// premixed operands of the direct path (mixed-canonical gauge, no interior-to-interior MPO blocks)
// ---------------------------------------------------------------------------------------------
// A side: one block per left-bond index m.  Xa[(b, ri), (m, q)] = sum_p W[0, b, p, q] x[m, p, ri] for b < wr - 1:
// for a fixed column (m, q) the K index (b, ri) runs over contiguous ri, so the row x[m, :, :] (d * r doubles) is
// read once into shared memory, its column maxima are found, and every thread then turns 16 consecutive ri of one
// (q, b) piece into digits.  The same pass writes y0[m, q, :] = sum_p W[0, wr - 1, p, q] x[m, p, :] - shift x[m, q, :].
__global__ void __launch_bounds__(256, 4) oz_premix_a_kernel(const double* __restrict__ x, const double* __restrict__ W,
                                                          int r, int wr, int d, double* __restrict__ scale,
                                                          double* __restrict__ sumsq, int8_t* __restrict__ slices,
                                                          int64_t Kp, int64_t slice_stride, double* __restrict__ y0,
                                                          const double* __restrict__ shift_dev, int skip_zero_pieces) {
  extern __shared__ double xs[];  // [d][r + r / 16 + 1]
  __shared__ double wc[kPmMaxCh][kPmMaxD][kPmMaxD];  // [b][p][q], b = wr - 1 holds the y0 block
  __shared__ double red[kPmMaxD][8];
  __shared__ double sc_sh[kPmMaxD];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int nb = wr - 1;
  for (int idx = tid; idx < wr * d * d; idx += blockDim.x) {
    const int b = idx / (d * d), p = (idx / d) % d, q = idx % d;
    wc[b][p][q] = W[(b * d + p) * d + q];  // W[0, b, p, q]: the a = 0 row of the (wl, wr, d, d) tensor
  }
  const double* xrow = x + (int64_t)m * d * r;
  // one pad double per 16 entries: the digit loop reads 16 consecutive ri per thread (stride 17 doubles between the
  // threads of a warp: conflict-free), the other loops read consecutive ri
  const int rp = r + r / 16 + 1;
  auto xi = [rp](int p, int ri) { return p * rp + ri + (ri >> 4); };
  for (int idx = tid; idx < d * r; idx += blockDim.x) xs[xi(idx / r, idx % r)] = xrow[idx];
  __syncthreads();
  // column maxima over (b, ri) for every q
  double mx[kPmMaxD];
#pragma unroll
  for (int q = 0; q < kPmMaxD; ++q) mx[q] = 0.0;
  for (int ri = tid; ri < r; ri += blockDim.x) {
    double xv[kPmMaxD];
#pragma unroll
    for (int p = 0; p < kPmMaxD; ++p) xv[p] = p < d ? xs[xi(p, ri)] : 0.0;
    for (int b = 0; b < nb; ++b)
#pragma unroll
      for (int q = 0; q < kPmMaxD; ++q)
        if (q < d) {
          double v = 0.0;
#pragma unroll
          for (int p = 0; p < kPmMaxD; ++p)
            if (p < d) v = fma(wc[b][p][q], xv[p], v);
          mx[q] = fmax(mx[q], fabs(v));
        }
  }
#pragma unroll
  for (int q = 0; q < kPmMaxD; ++q) {
    double v = mx[q];
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[q][tid >> 5] = v;
  }
  __syncthreads();
  if (tid < d) {
    double v = 0.0;
    for (int w8 = 0; w8 < (int)(blockDim.x >> 5); ++w8) v = fmax(v, red[tid][w8]);
    const double sc = oz_scale_of((unsigned long long)__double_as_longlong(v));
    sc_sh[tid] = sc;
    scale[(int64_t)tid * gridDim.x + m] = sc;  // columns are ordered (q, m): a tile of the GEMM then has one q
    if (sc > 0.0) atomicAdd(sumsq, sc * sc);
  }
  __syncthreads();
  // digits: a thread owns 16 consecutive ri of one (q, b) row piece; K = (b, ri), padded with zeros up to Kp
  const int r16 = (r + 15) / 16;
  for (int item = tid; item < d * nb * r16; item += blockDim.x) {
    const int i16 = item % r16, b = (item / r16) % nb, q = item / (r16 * nb);
    const double sc = sc_sh[q];
    const double inv = sc > 0.0 ? 1.0 / sc : 0.0;
    double cw[kPmMaxD];
    bool any = false;
#pragma unroll
    for (int p = 0; p < kPmMaxD; ++p) {
      cw[p] = p < d ? wc[b][p][q] * inv : 0.0;  // exact: inv is a power of two
      any = any || (p < d && wc[b][p][q] != 0.0);
    }
    if (skip_zero_pieces && !any) continue;  // column q of W[0, b] vanishes: the GEMM skips this K range for block q
    uint4 dig[kOzMaxSlices];
    oz_digits16(
        [&](int e) {
          const int ri = 16 * i16 + e;
          double v = 0.0;
          if (ri < r) {
#pragma unroll
            for (int p = 0; p < kPmMaxD; ++p)
              if (p < d) v = fma(cw[p], xs[xi(p, ri)], v);
          }
          return v;
        },
        dig);
    const int64_t k0 = (int64_t)b * r + 16 * i16;
    int8_t* row = slices + ((int64_t)q * gridDim.x + m) * Kp + k0;
    const bool last_piece = (b == nb - 1) && (16 * i16 + 16 > r);  // runs into the zero padding: still inside Kp?
    if (16 * i16 + 16 <= r || (last_piece && k0 + 16 <= Kp)) {
      oz_store16(row, slice_stride, dig, (k0 & 15) == 0);
    } else {
      // ragged end of a piece that is followed by the next channel's bytes: store only the valid ones
      const int valid = r - 16 * i16;
#pragma unroll
      for (int sl = 0; sl < kOzMaxSlices; ++sl) {
        const uint32_t w[4] = {dig[sl].x, dig[sl].y, dig[sl].z, dig[sl].w};
        for (int e = 0; e < valid; ++e) row[sl * slice_stride + e] = (int8_t)(w[e >> 2] >> (8 * (e & 3)));
      }
    }
  }
  // zero what is left of the K padding of this block's columns (Kp - nb * r < 64 bytes per row)
  const int kreal = nb * r, kpad = (int)(Kp - kreal);
  for (int item = tid; item < d * kOzMaxSlices * kpad; item += blockDim.x) {
    const int e = item % kpad, sl = (item / kpad) % kOzMaxSlices, q = item / (kpad * kOzMaxSlices);
    slices[sl * slice_stride + ((int64_t)q * gridDim.x + m) * Kp + kreal + e] = 0;
  }
  if (y0 != nullptr) {
    const double shift = shift_dev ? *shift_dev : 0.0;
    for (int idx = tid; idx < d * r; idx += blockDim.x) {
      const int q = idx / r, ri = idx % r;
      double v = -shift * xs[xi(q, ri)];
      for (int p = 0; p < d; ++p) v = fma(wc[nb][p][q], xs[xi(p, ri)], v);
      y0[(int64_t)m * d * r + idx] = v;
    }
  }
}

// B side, column maxima: colmax[(q, s)] = max over (a, li) of |sum_p W[a + 1, wr - 1, p, q] x[li, p, s]|
__global__ void __launch_bounds__(256) oz_premix_b_colmax_kernel(const double* __restrict__ x,
                                                                 const double* __restrict__ W, int l, int r, int wl,
                                                                 int wr, int d, int l_chunk,
                                                                 unsigned long long* __restrict__ colmax) {
  __shared__ double wc[kPmMaxCh][kPmMaxD][kPmMaxD];  // [a][p][q] = W[a + 1, wr - 1, p, q]
  const int na = wl - 1;
  for (int idx = threadIdx.x; idx < na * d * d; idx += blockDim.x) {
    const int a = idx / (d * d), p = (idx / d) % d, q = idx % d;
    wc[a][p][q] = W[((((int64_t)(a + 1)) * wr + (wr - 1)) * d + p) * d + q];
  }
  __syncthreads();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= r) return;
  const int l0 = blockIdx.y * l_chunk, l1 = min(l, l0 + l_chunk);
  double mx[kPmMaxD];
#pragma unroll
  for (int q = 0; q < kPmMaxD; ++q) mx[q] = 0.0;
#pragma unroll 4
  for (int li = l0; li < l1; ++li) {
    double xv[kPmMaxD];
#pragma unroll
    for (int p = 0; p < kPmMaxD; ++p) xv[p] = p < d ? x[((int64_t)li * d + p) * r + s] : 0.0;
    for (int a = 0; a < na; ++a)
#pragma unroll
      for (int q = 0; q < kPmMaxD; ++q)
        if (q < d) {
          double v = 0.0;
#pragma unroll
          for (int p = 0; p < kPmMaxD; ++p)
            if (p < d) v = fma(wc[a][p][q], xv[p], v);
          mx[q] = fmax(mx[q], fabs(v));
        }
  }
#pragma unroll
  for (int q = 0; q < kPmMaxD; ++q)
    if (q < d && mx[q] > 0.0) atomicMax(&colmax[(int64_t)q * r + s], (unsigned long long)__double_as_longlong(mx[q]));
}

// B side, digits: Xb[(a, li), (q, s)] for a < wl - 1.  Block = 32 values of s x 8 chunks of 16 li: the x loads are
// coalesced along s, a thread keeps its 16 x d entries of x in registers and forms every (a, q) product from them,
// each one leaving as 16 consecutive bytes of the (slice, column (q, s)) rows.
template <int D>
__global__ void __launch_bounds__(256, D <= 2 ? 2 : 1) oz_premix_b_kernel(const double* __restrict__ x, const double* __restrict__ W,
                                                          int l, int r, int wl, int wr,
                                                          const unsigned long long* __restrict__ colmax,
                                                          double* __restrict__ scale, double* __restrict__ sumsq,
                                                          int8_t* __restrict__ slices, int64_t Kp,
                                                          int64_t slice_stride, int skip_zero_pieces) {
  constexpr int d = D;
  __shared__ double wc[kPmMaxCh][D][D];
  const int na = wl - 1;
  for (int idx = threadIdx.x; idx < na * d * d; idx += blockDim.x) {
    const int a = idx / (d * d), p = (idx / d) % d, q = idx % d;
    wc[a][p][q] = W[((((int64_t)(a + 1)) * wr + (wr - 1)) * d + p) * d + q];
  }
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int s = blockIdx.x * 32 + tx;
  const int l0 = (blockIdx.y * 8 + ty) * 16;
  double sc[D], inv[D];
#pragma unroll
  for (int q = 0; q < D; ++q) {
    sc[q] = s < r ? oz_scale_of(colmax[(int64_t)q * r + s]) : 0.0;
    inv[q] = sc[q] > 0.0 ? 1.0 / sc[q] : 0.0;
  }
  if (blockIdx.y == 0 && ty == 0) {
    double part = 0.0;
#pragma unroll
    for (int q = 0; q < D; ++q) {
      if (s < r) scale[(int64_t)q * r + s] = sc[q];
      part += sc[q] * sc[q];
    }
    part = warp_sum(part);
    if (tx == 0 && part > 0.0) atomicAdd(sumsq, part);
  }
  __shared__ __align__(16) OzSliceTile tile;
  __syncthreads();
  double xv[16][D];
#pragma unroll
  for (int e = 0; e < 16; ++e)
#pragma unroll
    for (int p = 0; p < D; ++p) xv[e][p] = (s < r && l0 + e < l) ? x[((int64_t)(l0 + e) * d + p) * r + s] : 0.0;
  const int64_t lb = (int64_t)blockIdx.y * 128;  // first li of this block's tile
  for (int a = 0; a < na; ++a)
#pragma unroll
    for (int q = 0; q < D; ++q) {
      double cw[D];
      bool any = false;
#pragma unroll
      for (int p = 0; p < D; ++p) {
        cw[p] = wc[a][p][q] * inv[q];  // exact scaling: inv is a power of two
        any = any || wc[a][p][q] != 0.0;
      }
      if (skip_zero_pieces && !any) continue;  // uniform over the block: the GEMM skips this K range for block q
      uint4 dig[kOzMaxSlices];
      oz_digits16(
          [&](int e) {
            double v = 0.0;
#pragma unroll
            for (int p = 0; p < D; ++p) v = fma(cw[p], xv[e][p], v);
            return v;
          },
          dig);
      oz_tile_put(tile, tx, ty, dig);
      __syncthreads();
      // K = (a, li): this tile holds li in [lb, lb + 128); the last channel may run on into the zero padding
      const int64_t kend = (a == na - 1) ? Kp : (int64_t)(a + 1) * l;
      oz_tile_flush(tile, slices + (int64_t)q * r * Kp, slice_stride, Kp, (int64_t)blockIdx.x * 32, r, (int64_t)a * l + lb,
                    min(kend, (int64_t)a * l + lb + 128));
      __syncthreads();
    }
}

// *bound = max(*bound, coef * sqrt(sa2 * sb2)): the normwise error bound of one product (header comment)
__global__ void oz_bound_kernel(const double* __restrict__ sa2, const double* __restrict__ sb2, double coef,
                                double* __restrict__ bound) {
  const double b = coef * sqrt(*sa2 * *sb2);
  if (b > *bound) *bound = b;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = oz_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// CTA pair (cta_group::2), 256 x 128 tile, two accumulator passes.
// ncu on the first-generation kernel (one CTA per 128 x 64 tile, all eight accumulators at once; removed, its
// capture is profiles/r01_oz_mma_kernel_ncu_full_raw.csv) showed the tensor cores' shared-memory
// operand pipe at 74-89 % of peak with the MMA pipe only 50-60 % busy: an M=128, N=64, K=32 int8 MMA reads
// 4 KB of A + 2 KB of B = 48 wavefronts of 128 B for 32 cycles of math.  N per instruction is capped by TMEM
// (S accumulators x N columns <= 512), so this kernel (a) runs the S diagonals in two passes of <= 4
// accumulators, which allows N = 128, and (b) pairs two CTAs (cta_group::2, M = 256): each CTA stages its own
// 128 rows of A and only half (64 rows) of B, i.e. 4 + 2 KB per 64 cycles of math = 75 % of the operand pipe.
//
//   pass 0: diagonals d = 4 .. S-1 (small weights; needs all slices: two 48 KB slots per 64-byte K chunk)
//   pass 1: diagonals d = 0 .. 3   (needs slices 0..3 of both operands: one slot per K chunk)
//
// Shared-memory ring: 4 slots of {A slices[4][128 rows][64 B], B slices[4][64 rows][64 B]} per CTA.  Both CTAs'
// TMA loads complete on the leader's full barrier; the leader's elected thread issues every MMA and frees
// slots / publishes accumulators in both CTAs with multicast commits; each CTA's epilogue warps drain their
// own 128 TMEM lanes (pass 0 writes C, pass 1 adds to it) and release TMEM to the leader between the passes.
constexpr int kOz2SlotA = 4 * kOzBM * kOzBK;  // 32 KB
constexpr int kOz2SlotB = 4 * kOzBN * kOzBK;  // 16 KB
constexpr int kOz2Slot = kOz2SlotA + kOz2SlotB;
constexpr int kOz2Slots = 4;
constexpr int kOz2TileN = 2 * kOzBN;  // 128 columns per accumulator
constexpr int kOz2EpiWarps = 8;       // two per TMEM lane quadrant
constexpr int kOz2PartDepth = 4;      // 16-column chunks of C requested ahead of use in the epilogue

// Tail splitting: tiles of the last, partly filled wave are cut along K (see oz2_mma_kernel).
struct Oz2Tail {
  int n_full;       // work items [0, n_full) are whole tiles
  int rem;          // number of split tiles (tile indices n_full .. n_full + rem - 1)
  int splits;       // K ranges per split tile
  double* scratch;  // (splits - 1) * rem plain 256 x 128 tiles receiving the K ranges ks > 0
};
constexpr int kOz2Threads = 128 + 32 * kOz2EpiWarps;

// K chunks a tile may skip because one operand is identically zero there (OzKSkip in ozaki.cuh): the tile's chunk
// sequence is the concatenation of the ranges of its block.
__device__ __forceinline__ int oz2_chunk_count(const OzKSkip& s, int blk) {
  int n = 0;
  for (int i = 0; i < s.nranges[blk]; ++i) n += s.hi[blk][i] - s.lo[blk][i];
  return n;
}
__device__ __forceinline__ int oz2_chunk_at(const OzKSkip& s, int blk, int seq) {
  for (int i = 0; i < s.nranges[blk]; ++i) {
    const int len = s.hi[blk][i] - s.lo[blk][i];
    if (seq < len) return s.lo[blk][i] + seq;
    seq -= len;
  }
  return 0;
}

__device__ __forceinline__ uint32_t oz_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void oz_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t oz_mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void oz_tma_3d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                               int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(oz_smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void oz_tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void oz_mma_i8_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n}\n"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z),
      "r"(z), "r"(z)
      : "memory");
}
// arrive (once the preceding MMAs have completed) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void oz_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %1;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n}\n"
      ::"r"(oz_smem_u32(bar)), "r"(3u)
      : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Issue the MMAs of one 64-byte K chunk for pass PASS (0: diagonals 4..S-1, 1: diagonals 0..3).  lo4 / hi4 are the
// (address >> 4) fields of the slot(s) holding slice groups 0..3 / 4..7; everything else is a compile-time
// constant so the single issuing thread spends two integer adds per MMA.
template <int S, int PASS>
__device__ __forceinline__ void oz2_issue_chunk(uint32_t lo4, uint32_t hi4, uint32_t tmem_base, bool first_chunk) {
  constexpr int d_lo = PASS == 0 ? 4 : 0, d_hi = PASS == 0 ? S - 1 : 3;
  // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 128, M = 256 (128 rows per CTA)
  constexpr uint32_t idesc =
      (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kOz2TileN >> 3) << 17) | ((uint32_t)((2 * kOzBM) >> 4) << 24);
  // K-major, 64B swizzle: SBO = 8 rows * 64 B, descriptor version 1, layout type SWIZZLE_64B
  constexpr uint64_t desc_hi = ((uint64_t)((8 * kOzBK) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
#pragma unroll
  for (int kk = 0; kk < kOzBK / 32; ++kk) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const uint64_t da = desc_hi | (uint64_t)((i < 4 ? lo4 : hi4) + (uint32_t)(((i & 3) * (kOzBM * kOzBK) + kk * 32) >> 4));
#pragma unroll
      for (int j = 0; j < S; ++j) {
        if (i + j < d_lo || i + j > d_hi) continue;
        const uint64_t db = desc_hi | (uint64_t)((j < 4 ? lo4 : hi4) +
                                                 (uint32_t)((kOz2SlotA + (j & 3) * (kOzBN * kOzBK) + kk * 32) >> 4));
        // the first contribution to every diagonal of a pass comes from slice i = 0 at k-step 0 of chunk 0
        const uint32_t acc = (kk == 0 && i == 0) ? (first_chunk ? 0u : 1u) : 1u;
        oz_mma_i8_pair(tmem_base + (uint32_t)((i + j - d_lo) * kOz2TileN), da, db, idesc, acc);
      }
    }
  }
}

template <int S, int PASS>
__device__ __forceinline__ void oz2_issue_pass(uint32_t smem0, uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tmem_full,
                                               uint32_t tmem_base, int KT, uint32_t& slot, uint32_t& phase) {
  for (int kt = 0; kt < KT; ++kt) {
    const uint32_t lo = smem0 + slot * kOz2Slot;
    oz_mbar_wait(&full_bar[slot], phase);
    if (PASS == 0) oz_mbar_wait(&full_bar[slot + 1], phase);  // slots are taken in pairs (0,1) / (2,3): same phase
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    oz2_issue_chunk<S, PASS>(lo >> 4, (lo + kOz2Slot) >> 4, tmem_base, kt == 0);
    oz_commit_pair(&empty_bar[slot]);
    if (PASS == 0) oz_commit_pair(&empty_bar[slot + 1]);
    slot += PASS == 0 ? 2 : 1;
    if (slot == kOz2Slots) { slot = 0; phase ^= 1; }
  }
  oz_commit_pair(tmem_full);
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kOz2Threads, 1)
    oz2_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const double* __restrict__ scaleA, const double* __restrict__ scaleB, GemmOut out, int M, int N,
                   int KT, int accumulate, Oz2Tail tail, const __grid_constant__ OzKSkip skip) {
  static_assert(S > 4 && S <= 8, "two passes of at most four diagonals");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOz2Slots * kOz2Slot);
  uint64_t* empty_bar = full_bar + kOz2Slots;
  uint64_t* tmem_full = empty_bar + kOz2Slots;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = oz_cluster_rank();
  int tm, tn, ks = 0, kt0 = 0, kt1 = KT, kblk = 0;  // [kt0, kt1): positions in the tile's chunk sequence
  double* tail_tile = nullptr;
  {
    const int tiles_m = (M + 2 * kOzBM - 1) / (2 * kOzBM), tiles_n = (N + kOz2TileN - 1) / kOz2TileN;
    constexpr int GROUP = 4;
    // work items: the first tail.n_full items are whole tiles; the tiles of the last, partly filled wave are cut
    // into tail.splits K ranges each so that the wave fills the machine (part ks > 0 goes to a scratch tile)
    int tile = blockIdx.x >> 1;
    int split_part = -1;
    if (tile >= tail.n_full) {
      const int j = tile - tail.n_full;
      tile = tail.n_full + j % tail.rem;
      ks = j / tail.rem;
      split_part = ks;
      if (ks > 0) tail_tile = tail.scratch + (int64_t)((ks - 1) * tail.rem + (tile - tail.n_full)) * (2 * kOzBM * kOz2TileN);
    }
    const int per_group = GROUP * tiles_n;
    const int gid = tile / per_group, first_m = gid * GROUP;
    const int gsz = min(tiles_m - first_m, GROUP), rem = tile - gid * per_group;
    tm = first_m + rem % gsz;
    tn = rem / gsz;
    // the K chunks this tile runs: all KT of them, or the ranges of its block when an operand has zero blocks
    if (skip.mode != 0) {
      kblk = (skip.mode == 1 ? tm : tn) / skip.block_tiles;
      kt1 = oz2_chunk_count(skip, kblk);
    }
    if (split_part >= 0) {
      const int nk_tile = kt1;
      kt0 = (int)((int64_t)nk_tile * split_part / tail.splits);
      kt1 = (int)((int64_t)nk_tile * (split_part + 1) / tail.splits);
    }
  }
  const int m0 = tm * 2 * kOzBM + (int)rank * kOzBM;  // this CTA's rows of C / columns of A
  const int n0 = tn * kOz2TileN;                      // the pair's columns of C
  const int nb0 = n0 + (int)rank * kOzBN;             // the half of the B tile this CTA stages
  if (threadIdx.x == 0) {
    for (int s = 0; s < kOz2Slots; ++s) {
      oz_mbar_init(&full_bar[s], 1);
      oz_mbar_init(&empty_bar[s], 1);
    }
    oz_mbar_init(tmem_full, 1);
    oz_mbar_init(tmem_empty, 2 * kOz2EpiWarps);  // every epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  oz_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // load sequence: t in [0, 2 KT) = pass 0 (chunk t / 2, slice group t % 2), t in [2 KT, 3 KT) = pass 1 (group 0)
      const int nk = kt1 - kt0, T = 3 * nk;
      uint32_t slot = 0, phase = 0;
      for (int t = 0; t < T; ++t) {
        int kt, sub;
        if (t < 2 * nk) { kt = kt0 + (t >> 1); sub = t & 1; } else { kt = kt0 + t - 2 * nk; sub = 0; }
        if (skip.mode != 0) kt = oz2_chunk_at(skip, kblk, kt);
        oz_mbar_wait(&empty_bar[slot], phase ^ 1);
        if (rank == 0) oz_mbar_expect_tx(&full_bar[slot], 2 * kOz2Slot);  // both CTAs' boxes land on this barrier
        const uint32_t leader_full = oz_mapa(oz_smem_u32(&full_bar[slot]), 0);
        uint8_t* dst = smem + slot * kOz2Slot;
        oz_tma_3d_pair(dst, &tmA, leader_full, kt * kOzBK, m0, sub * 4);
        oz_tma_3d_pair(dst + kOz2SlotA, &tmB, leader_full, kt * kOzBK, nb0, sub * 4);
        if (++slot == kOz2Slots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t smem0 = oz_smem_u32(smem);
      uint32_t slot = 0, phase = 0;
      oz2_issue_pass<S, 0>(smem0, full_bar, empty_bar, tmem_full, tmem_base, kt1 - kt0, slot, phase);
      oz_mbar_wait(tmem_empty, 0);  // both CTAs' epilogues have drained the pass-0 accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      oz2_issue_pass<S, 1>(smem0, full_bar, empty_bar, tmem_full, tmem_base, kt1 - kt0, slot, phase);
    }
  } else if (warp >= 4) {
    // Epilogue: 8 warps; warp w drains TMEM lanes 32 (w % 4) + 16 ((w - 4) / 4) .. + 15 with 16x256b loads, whose
    // fragment is the mma accumulator layout: thread t holds lanes t/4 and t/4 + 8, columns 8 b + 2 (t % 4) + {0, 1}
    // of every 8-column block b -- four threads cover 64 contiguous bytes of a C row, so global accesses are
    // whole sectors without a shared-memory transpose.
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int lane_base = q * 32 + half * 16;
    const int r0 = lane_base + (lane >> 2), r1 = r0 + 8;
    const int mrow0 = m0 + r0, mrow1 = m0 + r1;
    const double sa0 = (mrow0 < M) ? scaleA[mrow0] : 0.0, sa1 = (mrow1 < M) ? scaleA[mrow1] : 0.0;
    // row pointers are biased by the tile's first column: element (row, n0 + c) lives at crow[c]
    double* crow0 = nullptr;
    double* crow1 = nullptr;
    if (tail_tile != nullptr) {  // K part ks > 0 of a split tile: plain 256 x 128 scratch tile
      if (mrow0 < M) crow0 = tail_tile + (int64_t)((int)rank * kOzBM + r0) * kOz2TileN;
      if (mrow1 < M) crow1 = tail_tile + (int64_t)((int)rank * kOzBM + r1) * kOz2TileN;
    } else {
      if (mrow0 < M)
        crow0 = out.C + (int64_t)(mrow0 / out.m_inner) * out.c_outer + (int64_t)(mrow0 % out.m_inner) * out.c_inner + n0;
      if (mrow1 < M)
        crow1 = out.C + (int64_t)(mrow1 / out.m_inner) * out.c_outer + (int64_t)(mrow1 % out.m_inner) * out.c_inner + n0;
    }
    const int cpair = 2 * (lane & 3);
    const int ncols = min(kOz2TileN, N - n0);  // valid columns of this tile
    // 16-byte accesses when every pair this thread touches is aligned and in range (uniform per thread)
    const bool vec = (ncols == kOz2TileN) && ((reinterpret_cast<uintptr_t>(crow0) | reinterpret_cast<uintptr_t>(crow1)) & 15) == 0;
    const uint32_t leader_empty = oz_mapa(oz_smem_u32(tmem_empty), 0);
    for (int pass = 0; pass < 2; ++pass) {
      const int d_lo = pass == 0 ? 4 : 0, nd = pass == 0 ? S - 4 : 4;
      const bool add = pass == 1 || ((accumulate & 1) && tail_tile == nullptr);
      double wgt[4];
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) wgt[dd] = ldexp(1.0, -7 * (d_lo + dd + 2));
      // Current values of the elements this thread owns (read-modify-write of pass 1), kept kOz2PartDepth column
      // chunks ahead of their use; the first chunks are requested while the MMAs of this pass are still running.
      double part[kOz2PartDepth][2][2][2];  // [chunk ring][row][block][column of the pair]
      auto load_part = [&](int slot, int c0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int c = c0 + 8 * b + cpair;
          if (!add) {
            part[slot][0][b][0] = part[slot][0][b][1] = part[slot][1][b][0] = part[slot][1][b][1] = 0.0;
          } else if (vec) {
            const double2 z = make_double2(0.0, 0.0);
            const double2 p0 = crow0 ? *reinterpret_cast<const double2*>(crow0 + c) : z;
            const double2 p1 = crow1 ? *reinterpret_cast<const double2*>(crow1 + c) : z;
            part[slot][0][b][0] = p0.x; part[slot][0][b][1] = p0.y;
            part[slot][1][b][0] = p1.x; part[slot][1][b][1] = p1.y;
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              part[slot][0][b][e] = (crow0 != nullptr && c + e < ncols) ? crow0[c + e] : 0.0;
              part[slot][1][b][e] = (crow1 != nullptr && c + e < ncols) ? crow1[c + e] : 0.0;
            }
          }
        }
      };
#pragma unroll
      for (int pc = 0; pc < kOz2PartDepth; ++pc) load_part(pc, 16 * pc);
      oz_mbar_wait(tmem_full, (uint32_t)pass);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ci = 0; ci < kOz2TileN / 16; ++ci) {
        const int c0 = 16 * ci;
        uint32_t v[4][8];
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          if (dd < nd) {
            const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(dd * kOz2TileN + c0);
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[dd][0]), "=r"(v[dd][1]), "=r"(v[dd][2]), "=r"(v[dd][3]), "=r"(v[dd][4]), "=r"(v[dd][5]),
                           "=r"(v[dd][6]), "=r"(v[dd][7])
                         : "r"(taddr));
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double res[2][2][2];
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = c0 + 8 * b + cpair + e;
            const double sb = c < ncols ? scaleB[n0 + c] : 0.0;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {  // registers 4 b + 2 rr + e: lane t/4 + 8 rr, column 8 b + 2 (t % 4) + e
              double acc = 0.0;
#pragma unroll
              for (int dd = 3; dd >= 0; --dd)  // smallest weights first
                if (dd < nd) acc = fma((double)(int)v[dd][4 * b + 2 * rr + e], wgt[dd], acc);
              res[rr][b][e] = fma(acc * (rr == 0 ? sa0 : sa1), sb, part[ci % kOz2PartDepth][rr][b][e]);
            }
          }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int c = c0 + 8 * b + cpair;
          if (vec) {
            if (crow0) *reinterpret_cast<double2*>(crow0 + c) = make_double2(res[0][b][0], res[0][b][1]);
            if (crow1) *reinterpret_cast<double2*>(crow1 + c) = make_double2(res[1][b][0], res[1][b][1]);
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (crow0 != nullptr && c + e < ncols) crow0[c + e] = res[0][b][e];
              if (crow1 != nullptr && c + e < ncols) crow1[c + e] = res[1][b][e];
            }
          }
        }
        if (ci + kOz2PartDepth < kOz2TileN / 16) load_part(ci % kOz2PartDepth, c0 + 16 * kOz2PartDepth);
      }
      if (pass == 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) oz_mbar_arrive_remote(leader_empty);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  oz_cluster_sync();  // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}
// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 oz_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// slices[s][row][k]: dims (Kp, rows, nslices), box (kOzBK, box_rows, 4), 64B swizzle.  nslices = the slices the product
// uses: the box of the second slice group then reaches past the tensor for S < 8, and TMA fills the part outside with
// zeros without reading it -- the unused slices cost no L2 / HBM traffic (the MMAs never touch them).
static int oz_make_map(CUtensorMap* map, const int8_t* base, int64_t Kp, int rows, int box_rows, int nslices) {
  auto enc = oz_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return TNPY_ECUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)nslices};
  cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * rows};
  cuuint32_t box[3] = {(cuuint32_t)kOzBK, (cuuint32_t)box_rows, 4};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ozaki: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TNPY_ECUDA;
  }
  return TNPY_OK;
}

int64_t oz_kp(int K) { return ((int64_t)K + kOzBK - 1) / kOzBK * kOzBK; }

size_t oz_operand_bytes(int cols, int K) {
  return Workspace::need((size_t)kOzMaxSlices * cols * oz_kp(K), 1) + Workspace::need(cols) + Workspace::need(cols + 1);
}

bool oz_operand_take(Workspace& ws, int cols, int K, OzOperand* out) {
  out->cols = cols;
  out->K = K;
  out->Kp = oz_kp(K);
  out->slices = ws.take<int8_t>((size_t)kOzMaxSlices * cols * out->Kp);
  out->scale = ws.take<double>(cols);
  out->colmax = ws.take<unsigned long long>(cols + 1);  // [cols] is the sum of squared scales
  if (!out->slices || !out->scale || !out->colmax) return false;
  out->sumsq = reinterpret_cast<double*>(out->colmax + cols);
  return true;
}

int oz_slice_operand(const double* P, int64_t ld, OzRowMap rows, const OzOperand& op, cudaStream_t stream) {
  const int K = op.K, MN = op.cols;
  TNPY_CUDA_OK(cudaMemsetAsync(op.colmax, 0, sizeof(unsigned long long) * (size_t)(MN + 1), stream));
  const int k_chunk = K > 4096 ? 256 : 64;
  dim3 g1(ceil_div(MN, 256), ceil_div(K, k_chunk));
  oz_colmax_kernel<<<g1, 256, 0, stream>>>(P, ld, rows, K, MN, k_chunk, op.colmax);
  TNPY_LAUNCH_OK();
  dim3 grid(ceil_div(MN, 32), (unsigned)((op.Kp + 127) / 128));  // 8 chunks of 16 k per block
  oz_slice_kernel<<<grid, 256, 0, stream>>>(P, ld, rows, K, MN, op.colmax, op.scale, op.sumsq, op.slices, op.Kp,
                                            (int64_t)MN * op.Kp);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

static bool premix_dims_ok(int r, int wl, int wr, int d) {
  return d <= kPmMaxD && wl - 1 <= kPmMaxCh && wr <= kPmMaxCh && wl >= 2 && wr >= 2;
}
// padded row length of the x row staged by oz_premix_a_kernel (one extra double per 16: conflict-free 16-strided reads)
static size_t premix_a_smem(int r, int d) { return sizeof(double) * (size_t)d * (r + r / 16 + 1); }

bool oz_premix_applicable(int l, int r, int wl, int wr, int d) {
  return premix_dims_ok(r, wl, wr, d) && premix_a_smem(r, d) <= 200 * 1024 && l >= 1;
}

int oz_premix_a(const double* x, const double* W, int l, int r, int wl, int wr, int d, const OzOperand& op, double* y0,
                const double* shift_dev, bool skip_zero_pieces, cudaStream_t stream) {
  TNPY_CHECK_ARG(oz_premix_applicable(l, r, wl, wr, d), "dimensions outside the direct path's limits");
  TNPY_CHECK_ARG(op.cols == l * d && op.K == (wr - 1) * r, "operand shape mismatch");
  TNPY_TRY(set_max_dynamic_smem(oz_premix_a_kernel, 200 * 1024));
  TNPY_CUDA_OK(cudaMemsetAsync(op.sumsq, 0, sizeof(double), stream));
  oz_premix_a_kernel<<<l, 256, premix_a_smem(r, d), stream>>>(x, W, r, wr, d, op.scale, op.sumsq, op.slices, op.Kp,
                                                             (int64_t)op.cols * op.Kp, y0, shift_dev, skip_zero_pieces ? 1 : 0);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

int oz_premix_b(const double* x, const double* W, int l, int r, int wl, int wr, int d, const OzOperand& op,
                bool skip_zero_pieces, cudaStream_t stream) {
  TNPY_CHECK_ARG(oz_premix_applicable(l, r, wl, wr, d), "dimensions outside the direct path's limits");
  TNPY_CHECK_ARG(op.cols == d * r && op.K == (wl - 1) * l, "operand shape mismatch");
  TNPY_CUDA_OK(cudaMemsetAsync(op.colmax, 0, sizeof(unsigned long long) * (size_t)(op.cols + 1), stream));
  const int l_chunk = l > 4096 ? 128 : 32;
  oz_premix_b_colmax_kernel<<<dim3(ceil_div(r, 256), ceil_div(l, l_chunk)), 256, 0, stream>>>(x, W, l, r, wl, wr, d,
                                                                                             l_chunk, op.colmax);
  TNPY_LAUNCH_OK();
  const dim3 grid(ceil_div(r, 32), ceil_div(l, 128));
  const int64_t stride = (int64_t)op.cols * op.Kp;
  const int skip = skip_zero_pieces ? 1 : 0;
  switch (d) {
    case 1: oz_premix_b_kernel<1><<<grid, 256, 0, stream>>>(x, W, l, r, wl, wr, op.colmax, op.scale, op.sumsq, op.slices, op.Kp, stride, skip); break;
    case 2: oz_premix_b_kernel<2><<<grid, 256, 0, stream>>>(x, W, l, r, wl, wr, op.colmax, op.scale, op.sumsq, op.slices, op.Kp, stride, skip); break;
    case 3: oz_premix_b_kernel<3><<<grid, 256, 0, stream>>>(x, W, l, r, wl, wr, op.colmax, op.scale, op.sumsq, op.slices, op.Kp, stride, skip); break;
    default: oz_premix_b_kernel<4><<<grid, 256, 0, stream>>>(x, W, l, r, wl, wr, op.colmax, op.scale, op.sumsq, op.slices, op.Kp, stride, skip); break;
  }
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// C tile += scratch tiles of the K ranges ks = 1 .. splits-1, in that fixed order (deterministic).
__global__ void __launch_bounds__(256) oz2_tail_combine_kernel(GemmOut out, int M, int N, int tiles_m, int tiles_n, Oz2Tail tail) {
  const int idx = blockIdx.x;  // split tile
  const int tile = tail.n_full + idx;
  constexpr int GROUP = 4;
  const int per_group = GROUP * tiles_n;
  const int gid = tile / per_group, first_m = gid * GROUP;
  const int gsz = min(tiles_m - first_m, GROUP), rem = tile - gid * per_group;
  const int m0 = (first_m + rem % gsz) * 2 * kOzBM, n0 = (rem / gsz) * kOz2TileN;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < 2 * kOzBM * kOz2TileN; e += gridDim.y * blockDim.x) {
    const int r = e / kOz2TileN, c = e % kOz2TileN;
    const int m = m0 + r, n = n0 + c;
    if (m >= M || n >= N) continue;
    double* dst = out.C + (int64_t)(m / out.m_inner) * out.c_outer + (int64_t)(m % out.m_inner) * out.c_inner + n;
    double acc = *dst;
    for (int ks = 1; ks < tail.splits; ++ks) acc += tail.scratch[(int64_t)((ks - 1) * tail.rem + idx) * (2 * kOzBM * kOz2TileN) + e];
    *dst = acc;
  }
}

// One CTA pair per SM pair at a time (shared memory): the last wave holds tiles % pairs tiles.  Cut those along K
// so that the last wave is as wide as the machine; parts ks > 0 go to scratch tiles and are added back in a fixed
// order.  Worth it only when the last wave is a sizeable part of the run (each part pays its own two epilogues).
static int oz_min_chunks(const OzKSkip& skip, int KT) {
  if (skip.mode == 0) return KT;
  int best = KT;
  for (int b = 0; b < skip.nblocks; ++b) {
    int n = 0;
    for (int i = 0; i < skip.nranges[b]; ++i) n += skip.hi[b][i] - skip.lo[b][i];
    if (n < best) best = n;
  }
  return best;
}

static Oz2Tail oz2_plan_tail(int M, int N, int KT) {
  const int tiles = ceil_div(M, 2 * kOzBM) * ceil_div(N, kOz2TileN);
  const int pairs = sm_count() / 2;
  Oz2Tail tail{tiles, 0, 1, nullptr};
  const int rem = tiles % pairs;
  if (rem > 0) {
    int splits = pairs / rem;
    while (splits > 1 && KT / splits < 8) --splits;
    if (splits > 4) splits = 4;
    const double full = (double)(tiles / pairs);
    if (splits > 1 && (full + 1.0 / splits + 0.04) / (full + 1.0) < 0.93) tail = Oz2Tail{tiles - rem, rem, splits, nullptr};
  }
  return tail;
}

size_t oz_mma_scratch_bytes(int M, int N) {
  // upper bound over K: at most 3 extra parts of fewer than `pairs` tiles
  const int pairs = sm_count() / 2;
  const int tiles = ceil_div(M, 2 * kOzBM) * ceil_div(N, kOz2TileN);
  const int rem = tiles % pairs;
  return rem == 0 ? 256 : Workspace::need((size_t)3 * rem * 2 * kOzBM * kOz2TileN) + 256;
}

template <int S>
static int oz2_launch(const OzOperand& A, const OzOperand& B, GemmOut out, int M, int N, int accumulate, Workspace& ws,
                      const OzKSkip& skip, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  TNPY_TRY(oz_make_map(&tmA, A.slices, A.Kp, M, kOzBM, S));
  TNPY_TRY(oz_make_map(&tmB, B.slices, B.Kp, N, kOzBN, S));
  constexpr int smem = kOz2Slots * kOz2Slot + 256 + 1024;
  TNPY_TRY(set_max_dynamic_smem(oz2_mma_kernel<S>, smem));
  const int tiles_m = ceil_div(M, 2 * kOzBM), tiles_n = ceil_div(N, kOz2TileN);
  const int KT = (int)(A.Kp / kOzBK);
  Oz2Tail tail = oz2_plan_tail(M, N, oz_min_chunks(skip, KT));
  if (tail.splits > 1) {
    tail.scratch = ws.take<double>((size_t)(tail.splits - 1) * tail.rem * 2 * kOzBM * kOz2TileN);
    if (!tail.scratch) tail = Oz2Tail{tiles_m * tiles_n, 0, 1, nullptr};  // no room: run the tail unsplit
  }
  const int items = tail.n_full + tail.rem * tail.splits;
  oz2_mma_kernel<S><<<2 * items, kOz2Threads, smem, stream>>>(tmA, tmB, A.scale, B.scale, out, M, N, KT, accumulate, tail, skip);
  TNPY_LAUNCH_OK();
  if (tail.splits > 1) {
    oz2_tail_combine_kernel<<<dim3(tail.rem, 16), 256, 0, stream>>>(out, M, N, tiles_m, tiles_n, tail);
    TNPY_LAUNCH_OK();
  }
  return TNPY_OK;
}

int oz_mma(const OzOperand& A, const OzOperand& B, GemmOut out, int M, int N, int S, int accumulate, Workspace& ws,
           double* bound_dev, cudaStream_t stream, const OzKSkip* skip_in) {
  OzKSkip skip{};  // mode 0: every tile runs all chunks
  if (skip_in != nullptr && skip_in->mode != 0) {
    skip = *skip_in;
    bool ok = skip.block_tiles > 0 && skip.nblocks > 0 && skip.nblocks <= kOzSkipBlocks;
    for (int b = 0; ok && b < skip.nblocks; ++b) ok = skip.nranges[b] > 0 && skip.nranges[b] <= kOzSkipRanges;
    TNPY_CHECK_ARG(ok, "malformed K-skip description");
  }
  TNPY_CHECK_ARG(A.cols == M && B.cols == N && A.Kp == B.Kp && A.K == B.K, "operand shapes do not match the product");
  TNPY_CHECK_ARG(A.K <= 65536, "K too large for exact int32 accumulation");
  TNPY_CHECK_ARG(S >= 5 && S <= kOzMaxSlices, "slices must be 5 .. 8");
  if (bound_dev) {
    const double e_s = (S + 2) / 4.0 * ldexp(1.0, -7 * S);
    oz_bound_kernel<<<1, 1, 0, stream>>>(A.sumsq, B.sumsq, (double)A.K * e_s, bound_dev);
    TNPY_LAUNCH_OK();
  }
  switch (S) {
    case 5: return oz2_launch<5>(A, B, out, M, N, accumulate, ws, skip, stream);
    case 6: return oz2_launch<6>(A, B, out, M, N, accumulate, ws, skip, stream);
    case 7: return oz2_launch<7>(A, B, out, M, N, accumulate, ws, skip, stream);
    default: return oz2_launch<8>(A, B, out, M, N, accumulate, ws, skip, stream);
  }
}

static std::atomic<int> g_oz_slices{8};
int ozaki_slices() { return g_oz_slices.load(); }

// smallest product (M N K) sent to the int8 path; env TNPY_OZAKI_MIN_WORK overrides it (experiments)
static double oz_min_work() {
  static const double v = [] {
    const char* e = getenv("TNPY_OZAKI_MIN_WORK");
    const double x = e ? atof(e) : 0.0;
    return x > 0.0 ? x : 6.0e9;
  }();
  return v;
}

bool ozaki_applicable(int M, int N, int K) {
  // below ~chi = 1024 the slicing passes and extra launches cost more than the faster MMA saves (measured at chi = 512)
  return K <= 65536 && K >= 64 && M >= 128 && N >= 64 && (double)M * N * K >= oz_min_work();
}

}  // namespace tnpy

using namespace tnpy;

extern "C" int tnpy_set_ozaki_slices(int slices) {
  if (slices < 6 || slices > kOzMaxSlices) {
    set_error("tnpy_set_ozaki_slices: slices must be 6, 7 or 8");
    return TNPY_EINVAL;
  }
  g_oz_slices.store(slices);
  return TNPY_OK;
}

extern "C" size_t tnpy_ozaki_workspace_bytes(int M, int N, int K, int /*slices*/) {
  return oz_operand_bytes(M, K) + oz_operand_bytes(N, K) + oz_mma_scratch_bytes(M, N) + Workspace::need(1) + 1024;
}

// C[m,n] (+)= sum_k A[k,m] B[k,n] in FP64 accuracy on the int8 tensor cores (slices in 5..8; 5 = 35 bits, what the
// eigensolver's inexact-Krylov schedule uses for the late steps of a solve).
// phase: 0 = slice both operands and multiply, 1 = slice only (fills the workspace), 2 = multiply only
// (workspace already holds the slices of these operands) -- lets a caller time / reuse the parts.
extern "C" int tnpy_ozaki_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                  int M, int N, int K, int slices, int accumulate, int phase, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "bad argument");
  TNPY_CHECK_ARG(slices >= 5 && slices <= kOzMaxSlices, "slices must be 5 .. 8");
  TNPY_CHECK_ARG(K <= 65536, "K too large for exact int32 accumulation");
  Workspace ws(workspace, workspace_bytes);
  OzOperand a, b;
  double* bound = ws.take<double>(1);
  if (!bound || !oz_operand_take(ws, M, K, &a) || !oz_operand_take(ws, N, K, &b)) {
    set_error("tnpy_ozaki_gemm_tn: workspace too small");
    return TNPY_EWORKSPACE;
  }
  if (phase != 2) {
    TNPY_TRY(oz_slice_operand(A, lda, oz_plain_rows(K), a, stream));
    TNPY_TRY(oz_slice_operand(B, ldb, oz_plain_rows(K), b, stream));
  }
  if (phase != 1) return oz_mma(a, b, plain_out(C, ldc, M), M, N, slices, accumulate, ws, nullptr, stream);
  return TNPY_OK;
}

// The rigorous error bound of the same product (header comment of this file), for tests and callers that want to
// choose `slices` themselves: K e_S ||sa||_2 ||sb||_2 from the scales left in the workspace by a phase-0/1 call.
extern "C" int tnpy_ozaki_error_bound(int M, int N, int K, int slices, void* workspace, size_t workspace_bytes,
                                      double* bound_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(bound_dev && M > 0 && N > 0 && K > 0 && slices >= 5 && slices <= kOzMaxSlices, "bad argument");
  Workspace ws(workspace, workspace_bytes);
  OzOperand a, b;
  double* scratch = ws.take<double>(1);
  if (!scratch || !oz_operand_take(ws, M, K, &a) || !oz_operand_take(ws, N, K, &b)) {
    set_error("tnpy_ozaki_error_bound: workspace too small");
    return TNPY_EWORKSPACE;
  }
  TNPY_CUDA_OK(cudaMemsetAsync(bound_dev, 0, sizeof(double), stream));
  const double e_s = (slices + 2) / 4.0 * ldexp(1.0, -7 * slices);
  oz_bound_kernel<<<1, 1, 0, stream>>>(a.sumsq, b.sumsq, (double)K * e_s, bound_dev);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}
