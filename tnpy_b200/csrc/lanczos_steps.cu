// Whole Lanczos steps of a small site in one cooperative launch (the local solve of
// finite_dmrg.py:97-133 at the sizes of BASELINE configs[0], XXZ n=100 chi=60).
//
// At chi = 60 a Lanczos vector has 7200 elements and a step of the general eigensolver is 15 kernel launches of
// 3-20 us each, every one of them latency: 146 us per step, 20 k steps per run.  Here the grid stays resident and
// the phases of a step are separated by grid barriers instead of kernel boundaries:
//
//   P1  t1[(p r'), (a m)] = sum_l x[l, (p r')] L[l, (a m)]                       32 x 32 tiles, FP64 FMA
//   P2  y[(m q), s]       = sum_{(r' b)} A[(r' b), (q m)] R[(r' b), s],           tiles x K pieces -> partial sums
//       A[(r' b), (q m)]  = sum_{a p} W[a, b, p, q] t1[(p r'), (a m)]             formed while the tile is loaded
//                                                                                 (the non-zeros of W only)
//   P3  w = sum of the partial sums;  partial h = V^T w over the CTA's own elements
//   P4  h = sum of the partials;  w -= V h;  partial h' = V^T w                   (classical Gram-Schmidt, twice)
//   P5  w -= V h';  partial ||w||^2;  w published for the next step's P1
//   P6  beta = ||w||;  V[j + 1] = w / beta;  T[:, j] = T[j, :] = h + h'
//
// Every CTA owns a fixed range of vector elements (one per thread) through P3-P6, so w lives in a register, and its
// elements of all basis vectors stay in shared memory for the whole launch: the Gram-Schmidt phases never wait for L2
// (a grid barrier invalidates L1).  All sums run in a fixed order: results are bit-reproducible from run to run and
// from stream to stream.  One extra CTA, the watcher, follows the convergence of the lowest Ritz pair while the others
// work (csrc/ritz_watch.cuh) and makes the launch return by itself; the host then runs the Ritz kernel on the full T.
// Measured at (60, 2, 60), w = 5: 20-28 us per step against 120 us (profiles/r02_small_site_steps_chi60.json).
//
// Second kernel, for vectors of up to 2^20 elements: lanczos_gs_kernel, the Gram-Schmidt half of a step (P3-P6) behind
// a matvec that stays on the GEMM kernels.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "lanczos_steps.cuh"
#include "ritz_watch.cuh"

namespace cg = cooperative_groups;

namespace tnpy {

namespace {
constexpr int kTile = 32;        // output tile edge
constexpr int kSlab = 32;        // K rows per shared-memory stage
constexpr int kThreads = 256;
static_assert(kThreads == kWatchThreads, "the watcher's Sturm rounds use one shift per thread");
constexpr int kMaxTerms = 1024;  // wl * wr * d * d of the MPO tensor (its non-zeros are a fraction of that)
constexpr int kMaxGroups = 128;  // wr * d
constexpr int64_t kMaxVector = 32768;
constexpr int kMaxKSplit = 8;
constexpr int64_t kGsMaxVector = 1 << 20;  // fused Gram-Schmidt: longest vector (8192 elements of w per CTA in shared memory)
constexpr int kMaxGatherLoads = 19;  // ceil(148 work CTAs / 8 threads per coefficient)

struct StepsArgs {
  const double* L;
  const double* W;
  const double* R;
  double* V;
  int64_t ldv;
  double* T;
  double* status;
  double* t1;
  double* ypart;
  double* wbuf;
  double* part;  // [2][grid][kStepsMaxNcv]
  double* nrm;   // [grid]
  int* stop;     // [2], by step parity: set by the watcher CTA when the lowest Ritz pair has converged
  int beta_slot, steps_slot;
  int l, r, wl, wr, d;
  int j0, nsteps;
  int ksplit, kchunk, chunk;
  int ldc;            // leading dimension of a CTA's shared-memory copy of its basis columns (>= chunk, = 8 mod 32)
  unsigned long long* trace;  // diagnostics: CTA 0's and the watcher's %globaltimer at each phase boundary (or null)
  int arrow;          // T = diag(arrow kept Ritz values) + their coupling to row `arrow` + a tridiagonal tail
  double tol, anorm;  // the launch stops on its own once |beta z_last| <= 0.7 tol max(anorm, |theta|) (tol <= 0: never)
};

__device__ __forceinline__ void tile_mma(const double (*As)[kTile + 1], const double (*Bs)[kTile + 1], int tx, int ty,
                                         double (&acc)[2][2]) {
  // two independent chains per accumulator (even / odd rows of the slab): the loop is bound by the latency of
  // dependent FP64 FMAs, not by their throughput
  double odd[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 4
  for (int kk = 0; kk < kSlab; kk += 2) {
    const double a0 = As[kk][ty], a1 = As[kk][ty + 16];
    const double b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
    const double c0 = As[kk + 1][ty], c1 = As[kk + 1][ty + 16];
    const double d0 = Bs[kk + 1][tx], d1 = Bs[kk + 1][tx + 16];
    acc[0][0] = fma(a0, b0, acc[0][0]);
    acc[0][1] = fma(a0, b1, acc[0][1]);
    acc[1][0] = fma(a1, b0, acc[1][0]);
    acc[1][1] = fma(a1, b1, acc[1][1]);
    odd[0][0] = fma(c0, d0, odd[0][0]);
    odd[0][1] = fma(c0, d1, odd[0][1]);
    odd[1][0] = fma(c1, d0, odd[1][0]);
    odd[1][1] = fma(c1, d1, odd[1][1]);
  }
  acc[0][0] += odd[0][0];
  acc[0][1] += odd[0][1];
  acc[1][0] += odd[1][0];
  acc[1][1] += odd[1][1];
}

// partial[k] = sum over this CTA's elements of V[k][.] * w[.] for k < m, from the CTA's shared-memory copy of its
// basis columns (vs[k * ldc + e]): eight threads per k, each over every eighth element, then three shuffles.
// ldc % 32 == 8, so the four k of a warp sit in disjoint banks.
__device__ __forceinline__ void partial_dots(const double* vs, int ldc, int m, int count, const double* wsm, double* out) {
  const int sub = threadIdx.x & 7;
  for (int k = threadIdx.x >> 3; k < ((m + 31) & ~31); k += kThreads / 8) {  // whole warps stay in the loop
    double s = 0.0;
    if (k < m) {
      const double* vk = vs + k * ldc;
      for (int e = sub; e < count; e += 8) s = fma(vk[e], wsm[e], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (sub == 0 && k < m) out[k] = s;
  }
}

// h[k] = sum over the CTAs of their partials (fixed order), k < m: eight threads per k, every load of a thread in
// flight at once (the partials come from L2: one round trip per pass, not one per k)
__device__ __forceinline__ void gather_partials(const double* part, int ctas, int m, double* h) {
  const int sub = threadIdx.x & 7;
  for (int k = threadIdx.x >> 3; k < ((m + 31) & ~31); k += kThreads / 8) {
    double s = 0.0;
    if (k < m) {
      double v[kMaxGatherLoads];
#pragma unroll
      for (int i = 0; i < kMaxGatherLoads; ++i) {
        const int c = sub + 8 * i;
        v[i] = c < ctas ? part[(int64_t)c * kStepsMaxNcv + k] : 0.0;
      }
#pragma unroll
      for (int i = 0; i < kMaxGatherLoads; ++i) s += v[i];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (sub == 0 && k < m) h[k] = s;
  }
}

// ---- convergence estimate on the device: csrc/ritz_watch.cuh; this is the step kernel's verdict on it -----------
// |beta z_last| <= 0.7 tol max(anorm, |theta|) for the lowest pair (theta, z) of T~ ?
__device__ bool watch_finish(int m, int k, double tol, double anorm, double* e, int* ei) {
  WatchMem w(e);
  const int tid = threadIdx.x;
  watch_rounds(8, m, k, e, ei);  // whatever the intervals of the step left over
  watch_vector(m, k, e);
  if (tid == 0) {
    const double* z = w.z;
    const double resid = fabs(w.box[4] * z[m - 1]);
    const double theta = 0.5 * (w.box[0] + w.box[1]);
    ei[15] = (resid <= 0.7 * tol * fmax(anorm, fabs(theta))) ? 1 : 0;
    // Residuals fall geometrically: the last two estimates give the rate and with it the number of steps left;
    // the next estimate is made after half of them (at most eight), so most steps run without one
    const double thr = 0.7 * tol * fmax(anorm, fabs(theta));
    int skip = 1;
    if (w.box[11] != 0.0 && w.box[10] > resid && resid > thr && thr > 0.0 && (double)m > w.box[13]) {
      const double rate = log(w.box[10] / resid) / ((double)m - w.box[13]);
      const double left = log(resid / thr) / rate;
      skip = left < 4.0 ? 1 : (left > 16.0 ? 8 : (int)(0.5 * left));
    }
    w.box[12] = (double)(m + skip);
    w.box[13] = (double)m;
    w.box[8] = w.box[0];
    w.box[9] = w.box[1];
    w.box[10] = resid;
    w.box[11] = 1.0;
  }
  __syncthreads();
  const bool done = ei[15] != 0;
  __syncthreads();
  return done;
}

__device__ __forceinline__ void trace_mark(const StepsArgs& a, int step, int slot, bool watcher) {
  // 16 slots per step: 0-7 CTA 0, 8-15 the watcher
  if (a.trace != nullptr && threadIdx.x == 0 && (blockIdx.x == 0 || watcher) && step < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[step * 16 + (watcher ? 8 : 0) + slot] = t;
  }
}

__global__ void __launch_bounds__(kThreads, 1) lanczos_steps_kernel(const StepsArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double As[kSlab][kTile + 1];
  __shared__ double Bs[kSlab][kTile + 1];
  __shared__ double coef[kMaxTerms];
  __shared__ int toff[kMaxTerms];
  __shared__ int goff[kMaxGroups + 1];
  __shared__ double wsm[kThreads];
  __shared__ double hs[kStepsMaxNcv], h2s[kStepsMaxNcv];
  __shared__ double red[32];
  __shared__ double beta_sh;

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int l = a.l, r = a.r, wl = a.wl, wr = a.wr, d = a.d;
  const int n = l * d * r;
  const int M1 = d * r, N1 = wl * l, K1 = l;  // t1 is M1 x N1, leading dimension N1
  const int M2 = d * l, N2 = r, K2 = r * wr;  // rows c = (q, m), m fastest
  // the last CTA takes no tiles and owns no vector elements: it watches the convergence of the Ritz pair
  const int ctas = gridDim.x - 1;
  const bool watcher = (int)blockIdx.x == ctas;
  const int cta = watcher ? (1 << 29) : (int)blockIdx.x;
  __shared__ int est_i[16];
  if (watcher && threadIdx.x == 0) {
    a.stop[0] = a.stop[1] = 0;
    (&As[0][0])[kWatchBox + 11] = 0.0;  // no previous estimate yet (the watcher keeps its state in As)
    (&As[0][0])[kWatchBox + 12] = 0.0;
    (&As[0][0])[kWatchBox + 13] = 0.0;
  }

  // the non-zeros of W, grouped by the output pair (b, q), in (a, p) order
  const int ngroups = wr * d;
  if (tid < ngroups) {
    const int b = tid / d, q = tid % d;
    int count = 0;
    for (int ap = 0; ap < wl * d; ++ap) {
      const int aa = ap / d, p = ap % d;
      if (a.W[((aa * wr + b) * d + p) * d + q] != 0.0) ++count;
    }
    goff[tid + 1] = count;
  }
  __syncthreads();
  if (tid == 0) {
    goff[0] = 0;
    for (int g = 0; g < ngroups; ++g) goff[g + 1] += goff[g];
  }
  __syncthreads();
  if (tid < ngroups) {
    const int b = tid / d, q = tid % d;
    int pos = goff[tid];
    for (int ap = 0; ap < wl * d; ++ap) {
      const int aa = ap / d, p = ap % d;
      const double c = a.W[((aa * wr + b) * d + p) * d + q];
      if (c != 0.0) {
        coef[pos] = c;
        toff[pos] = p * r * N1 + aa * l;
        ++pos;
      }
    }
  }
  __syncthreads();

  const int64_t first = watcher ? (int64_t)n : (int64_t)cta * a.chunk;
  const int count = first >= n ? 0 : (int)min((int64_t)a.chunk, n - first);
  const bool own = tid < count;
  // this CTA's elements of every basis vector stay in shared memory for the whole launch: the Gram-Schmidt phases
  // then never wait for L2 (the barriers invalidate L1)
  extern __shared__ double vs[];
  const int ldc = a.ldc;
  if (own)
    for (int k = 0; k <= a.j0; ++k) vs[k * ldc + tid] = a.V[(int64_t)k * a.ldv + first + tid];
  double w = 0.0;
  const double* xsrc = a.V + (int64_t)a.j0 * a.ldv;
  double xscale = 1.0;
  double* part0 = a.part;
  double* part1 = a.part + (int64_t)ctas * kStepsMaxNcv;

  for (int step = 0; step < a.nsteps; ++step) {
    const int j = a.j0 + step, m = j + 1;
    trace_mark(a, step, 0, watcher);

    // ---- P1 ----
    {
      const int tiles_n = (N1 + kTile - 1) / kTile, tiles = ((M1 + kTile - 1) / kTile) * tiles_n;
      constexpr int kPer = kSlab * kTile / kThreads;  // slab elements per thread and operand
      for (int tile = cta; tile < tiles; tile += ctas) {
        const int m0 = (tile / tiles_n) * kTile, n0 = (tile % tiles_n) * kTile;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        double pa[kPer], pb[kPer];
        auto fetch = [&](int k0) {  // the slab's loads, all in flight together
#pragma unroll
          for (int it = 0; it < kPer; ++it) {
            const int idx = tid + it * kThreads, kk = idx >> 5, c = idx & 31, k = k0 + kk;
            pa[it] = (k < K1 && m0 + c < M1) ? xsrc[(int64_t)k * M1 + m0 + c] : 0.0;
            pb[it] = (k < K1 && n0 + c < N1) ? a.L[(int64_t)k * N1 + n0 + c] : 0.0;
          }
        };
        fetch(0);
        for (int k0 = 0; k0 < K1; k0 += kSlab) {
#pragma unroll
          for (int it = 0; it < kPer; ++it) {
            const int idx = tid + it * kThreads, kk = idx >> 5, c = idx & 31;
            As[kk][c] = pa[it] * xscale;
            Bs[kk][c] = pb[it];
          }
          __syncthreads();
          if (k0 + kSlab < K1) fetch(k0 + kSlab);  // the next slab travels while this one is multiplied
          tile_mma(As, Bs, tx, ty, acc);
          __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int row = m0 + ty + 16 * i, col = n0 + tx + 16 * jj;
            if (row < M1 && col < N1) a.t1[(int64_t)row * N1 + col] = acc[i][jj];
          }
      }
    }
    trace_mark(a, step, 1, watcher);
    grid.sync();
    trace_mark(a, step, 2, watcher);
    // The watcher examines T as the steps before this one left it (j columns; status holds the norm the last one
    // found) while the others run this step, and publishes its verdict before the next step's first barrier.
    double* const wmem = &As[0][0];
    const bool watching = watcher && step >= 1 && a.tol > 0.0 && a.arrow <= j - 1 && (double)j >= wmem[kWatchBox + 12];
    if (watching) {
      watch_begin(a.T, j, a.arrow, a.status[a.beta_slot], wmem);
      watch_rounds(1, j, a.arrow, wmem, est_i);
    }

    // ---- P2 ----
    {
      const int tiles_n = (N2 + kTile - 1) / kTile, tiles = ((M2 + kTile - 1) / kTile) * tiles_n;
      const int items = tiles * a.ksplit;
      constexpr int kPer = kSlab * kTile / kThreads;
      for (int item = cta; item < items; item += ctas) {
        const int ks = item % a.ksplit, tile = item / a.ksplit;
        const int m0 = (tile / tiles_n) * kTile, n0 = (tile % tiles_n) * kTile;
        const int kbeg = ks * a.kchunk, kend = min(K2, kbeg + a.kchunk);
        // the row (q, m) of this thread's slab elements does not change from slab to slab
        const int cc = m0 + (tid & 31);
        const int q_of = cc / l, mm_of = cc - q_of * l;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        double pa[kPer], pb[kPer];
        auto fetch = [&](int k0) {
          // A[(r' b), (q m)] = sum over the non-zeros W[a, b, p, q] of W * t1[(p r'), (a m)]: the loads of up to four
          // terms of all the thread's elements are issued before anything is summed (one L2 round trip per slab)
          double x[kPer][4], cf[kPer][4];
          int more[kPer];
#pragma unroll
          for (int it = 0; it < kPer; ++it) {
            const int kk = (tid >> 5) + it * (kThreads / 32), k = k0 + kk;
            const bool live = k < kend && cc < M2;
            const int rp = k / wr, b = k - rp * wr;
            const int g = live ? b * d + q_of : 0;
            const int t0 = goff[g], t1e = live ? goff[g + 1] : t0;
            const double* src = a.t1 + (int64_t)rp * N1 + mm_of;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool on = t0 + u < t1e;
              cf[it][u] = on ? coef[t0 + u] : 0.0;
              x[it][u] = on ? src[toff[t0 + u]] : 0.0;
            }
            more[it] = t1e - t0 > 4 ? t1e : 0;
            pb[it] = (k < kend && n0 + (tid & 31) < N2) ? a.R[(int64_t)k * r + n0 + (tid & 31)] : 0.0;
          }
#pragma unroll
          for (int it = 0; it < kPer; ++it) {
            double v = 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) v = fma(cf[it][u], x[it][u], v);
            if (more[it]) {  // a denser MPO tensor: the remaining terms one by one
              const int kk = (tid >> 5) + it * (kThreads / 32), k = k0 + kk;
              const int rp = k / wr, b = k - rp * wr, g = b * d + q_of;
              const double* src = a.t1 + (int64_t)rp * N1 + mm_of;
              for (int t = goff[g] + 4; t < more[it]; ++t) v = fma(coef[t], src[toff[t]], v);
            }
            pa[it] = v;
          }
        };
        fetch(kbeg);
        for (int k0 = kbeg; k0 < kend; k0 += kSlab) {
#pragma unroll
          for (int it = 0; it < kPer; ++it) {
            const int kk = (tid >> 5) + it * (kThreads / 32);
            As[kk][tid & 31] = pa[it];
            Bs[kk][tid & 31] = pb[it];
          }
          __syncthreads();
          if (k0 + kSlab < kend) fetch(k0 + kSlab);
          tile_mma(As, Bs, tx, ty, acc);
          __syncthreads();
        }
        double* yp = a.ypart + (int64_t)ks * n;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int cr = m0 + ty + 16 * i, col = n0 + tx + 16 * jj;
            if (cr < M2 && col < N2) {
              const int q = cr / l, mm = cr - q * l;
              yp[((int64_t)mm * d + q) * r + col] = acc[i][jj];
            }
          }
      }
    }
    trace_mark(a, step, 3, watcher);
    grid.sync();
    if (watching) watch_rounds(1, j, a.arrow, wmem, est_i);

    // ---- P3: w and the first pass' partial coefficients ----
    w = 0.0;
    if (own)
      for (int ks = 0; ks < a.ksplit; ++ks) w += a.ypart[(int64_t)ks * n + first + tid];
    wsm[tid] = w;
    __syncthreads();
    if (!watcher) partial_dots(vs, ldc, m, count, wsm, part0 + (int64_t)cta * kStepsMaxNcv);
    trace_mark(a, step, 4, watcher);
    grid.sync();
    if (watching) watch_rounds(1, j, a.arrow, wmem, est_i);

    // ---- P4: first pass applied, second pass' partial coefficients ----
    gather_partials(part0, ctas, m, hs);
    __syncthreads();
    if (own)
      for (int k = 0; k < m; ++k) w = fma(-hs[k], vs[k * ldc + tid], w);
    wsm[tid] = w;
    __syncthreads();
    if (!watcher) partial_dots(vs, ldc, m, count, wsm, part1 + (int64_t)cta * kStepsMaxNcv);
    trace_mark(a, step, 5, watcher);
    grid.sync();
    if (watching) watch_rounds(1, j, a.arrow, wmem, est_i);

    // ---- P5: second pass applied, partial norm, w published ----
    gather_partials(part1, ctas, m, h2s);
    __syncthreads();
    if (own) {
      for (int k = 0; k < m; ++k) w = fma(-h2s[k], vs[k * ldc + tid], w);
      a.wbuf[first + tid] = w;
    }
    {
      const double sq = block_sum(own ? w * w : 0.0, red);
      if (tid == 0 && !watcher) a.nrm[cta] = sq;
    }
    trace_mark(a, step, 6, watcher);
    grid.sync();
    trace_mark(a, step, 7, watcher);
    // the verdict on the previous step's T, written before this step's first barrier (slots alternate: the watcher
    // may be writing this step's verdict right now)
    const int stop = step >= 2 ? *(volatile int*)(a.stop + ((step - 1) & 1)) : 0;
    if (watching) {
      const bool converged = watch_finish(j, a.arrow, a.tol, a.anorm, wmem, est_i);
      if (tid == 0) a.stop[step & 1] = converged ? 1 : 0;
    } else if (watcher && tid == 0) {
      a.stop[step & 1] = 0;
    }

    // ---- P6: normalise, new column of T ----
    if (tid < 32) {
      double s = 0.0;
      for (int c = tid; c < ctas; c += 32) s += a.nrm[c];
      s = warp_sum(s);
      if (tid == 0) beta_sh = sqrt(s);
    }
    __syncthreads();
    const double beta = beta_sh;
    const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
    if (own) {
      a.V[(int64_t)(j + 1) * a.ldv + first + tid] = w * inv;
      vs[(j + 1) * ldc + tid] = w * inv;
    }
    if (cta == 0) {
      if (tid < m) {
        const double t = hs[tid] + h2s[tid];
        a.T[tid * kStepsMaxNcv + j] = t;
        a.T[j * kStepsMaxNcv + tid] = t;
      }
      if (tid == 0) {
        a.status[a.beta_slot] = beta;
        a.status[a.steps_slot] = (double)(step + 1);
      }
    }
    xsrc = a.wbuf;
    xscale = inv;
    if (!(beta > 0.0) || stop) break;  // exact breakdown / converged: every CTA sees the same values
    __syncthreads();           // beta_sh, hs, h2s are rewritten in the next step
  }
}

// ---- mid-size sites: the Gram-Schmidt half of a step in one cooperative launch -------------------------------------
// Between 2^15 and 2^20 unknowns the matvec belongs on the tensor-pipe GEMM kernels, but the rest of a step -- two
// classical Gram-Schmidt passes against the basis, the norm, the normalised copy, the new column of T -- is nine
// launches of 4-13 us on vectors that live in L2.  Here it is one: every CTA owns a contiguous range of elements, keeps
// its part of w in shared memory through three grid barriers, reads its part of the basis from L2 / HBM four times
// (dot, apply, dot, apply -- the same traffic as the separate kernels), and all sums run in a fixed order.
struct GsArgs {
  double* V;
  int64_t ldv;
  int64_t n;
  int j;
  double* T;
  double* status;
  int beta_slot;
  double* part;  // [2][grid][kStepsMaxNcv]
  double* nrm;   // [grid]
  int chunk;
};

// partial[k] = sum over this CTA's elements of V[k][first + .] * w[.]: one warp per k, eight k in flight
__device__ __forceinline__ void partial_dots_global(const double* V, int64_t ldv, int m, int count, const double* wsm, double* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < m; k += kThreads / 32) {
    const double* vk = V + (int64_t)k * ldv;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int e = lane;
    for (; e + 96 < count; e += 128) {
      const double a0 = vk[e], a1 = vk[e + 32], a2 = vk[e + 64], a3 = vk[e + 96];
      s0 = fma(a0, wsm[e], s0);
      s1 = fma(a1, wsm[e + 32], s1);
      s2 = fma(a2, wsm[e + 64], s2);
      s3 = fma(a3, wsm[e + 96], s3);
    }
    for (; e < count; e += 32) s0 = fma(vk[e], wsm[e], s0);
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) out[k] = s;
  }
}

__global__ void __launch_bounds__(kThreads, 1) lanczos_gs_kernel(const GsArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double wsm[];  // this CTA's elements of w
  __shared__ double hs[kStepsMaxNcv], h2s[kStepsMaxNcv];
  __shared__ double red[32];
  __shared__ double beta_sh;
  const int tid = threadIdx.x, ctas = gridDim.x, cta = blockIdx.x;
  const int j = a.j, m = j + 1;
  const int64_t first = (int64_t)cta * a.chunk;
  const int count = first >= a.n ? 0 : (int)min((int64_t)a.chunk, a.n - first);
  const double* Vc = a.V + first;                         // column k of the basis, this CTA's part: Vc + k * ldv
  double* wg = a.V + (int64_t)(j + 1) * a.ldv + first;    // H v_j as the matvec left it
  double* part0 = a.part;
  double* part1 = a.part + (int64_t)ctas * kStepsMaxNcv;

  for (int e = tid; e < count; e += kThreads) wsm[e] = wg[e];
  __syncthreads();
  partial_dots_global(Vc, a.ldv, m, count, wsm, part0 + (int64_t)cta * kStepsMaxNcv);
  grid.sync();

  gather_partials(part0, ctas, m, hs);
  __syncthreads();
  for (int e = tid; e < count; e += kThreads) {
    double w = wsm[e];
    for (int k = 0; k < m; ++k) w = fma(-hs[k], Vc[(int64_t)k * a.ldv + e], w);
    wsm[e] = w;
  }
  __syncthreads();
  partial_dots_global(Vc, a.ldv, m, count, wsm, part1 + (int64_t)cta * kStepsMaxNcv);
  grid.sync();

  gather_partials(part1, ctas, m, h2s);
  __syncthreads();
  double sq = 0.0;
  for (int e = tid; e < count; e += kThreads) {
    double w = wsm[e];
    for (int k = 0; k < m; ++k) w = fma(-h2s[k], Vc[(int64_t)k * a.ldv + e], w);
    wsm[e] = w;
    sq = fma(w, w, sq);
  }
  sq = block_sum(sq, red);
  if (tid == 0) a.nrm[cta] = sq;
  grid.sync();

  if (tid < 32) {
    double s = 0.0;
    for (int c = tid; c < ctas; c += 32) s += a.nrm[c];
    s = warp_sum(s);
    if (tid == 0) beta_sh = sqrt(s);
  }
  __syncthreads();
  const double beta = beta_sh;
  const double inv = beta > 0.0 ? 1.0 / beta : 0.0;
  for (int e = tid; e < count; e += kThreads) wg[e] = wsm[e] * inv;
  if (cta == 0) {
    if (tid < m) {
      const double t = hs[tid] + h2s[tid];
      a.T[tid * kStepsMaxNcv + j] = t;
      a.T[j * kStepsMaxNcv + tid] = t;
    }
    if (tid == 0) a.status[a.beta_slot] = beta;
  }
}

std::atomic<unsigned long long*>& trace_buffer() {
  static std::atomic<unsigned long long*> p(nullptr);
  return p;
}

std::atomic<int>& fused_steps_switch() {
  static std::atomic<int> on([] {
    const char* e = getenv("TNPY_FUSED_STEPS");
    return (e && e[0] == '0') ? 0 : 1;
  }());
  return on;
}
}  // namespace

bool lanczos_steps_supported(int l, int r, int wl, int wr, int d) {
  if (!fused_steps_switch().load(std::memory_order_relaxed)) return false;
  const int64_t n = (int64_t)l * d * r;
  return n <= kMaxVector && (n + kThreads - 1) / kThreads <= sm_count() - 1 && (int64_t)wl * wr * d * d <= kMaxTerms &&
         wr * d <= kMaxGroups && l >= 1 && r >= 1;
}

LanczosStepsPlan lanczos_steps_plan(int l, int r, int wl, int wr, int d) {
  LanczosStepsPlan p;
  const int64_t n = (int64_t)l * d * r;
  const int sms = sm_count();
  const int tiles1 = ceil_div(d * r, kTile) * ceil_div(wl * l, kTile);
  const int tiles2 = ceil_div(d * l, kTile) * ceil_div(r, kTile);
  int want = (int)((n + kThreads - 1) / kThreads);
  if (tiles1 > want) want = tiles1;
  if (tiles2 > want) want = tiles2;
  if (want < 8) want = 8;
  p.grid = want < sms - 1 ? want : sms - 1;  // work CTAs; the launch adds the one that watches convergence
  p.chunk = (int)((n + p.grid - 1) / p.grid);  // <= 256: grid >= n / 256 (n <= 32768 needs 128 CTAs, a B200 has 148)
  const int K2 = r * wr;
  int ksplit = p.grid / tiles2;
  if (ksplit > kMaxKSplit) ksplit = kMaxKSplit;
  if (ksplit > ceil_div(K2, kSlab)) ksplit = ceil_div(K2, kSlab);
  if (ksplit < 1) ksplit = 1;
  p.kchunk = ceil_div(ceil_div(K2, ksplit), kSlab) * kSlab;
  p.ksplit = ceil_div(K2, p.kchunk);
  p.bytes = Workspace::need((size_t)d * r * wl * l) + Workspace::need((size_t)p.ksplit * n) + Workspace::need((size_t)n) +
            Workspace::need((size_t)2 * p.grid * kStepsMaxNcv) + Workspace::need((size_t)p.grid) + Workspace::need(64, 1) + 256;
  return p;
}

bool lanczos_gs_supported(int64_t n) {
  return fused_steps_switch().load(std::memory_order_relaxed) != 0 && n > 0 && n <= kGsMaxVector;
}

size_t lanczos_gs_bytes() { return Workspace::need((size_t)2 * sm_count() * kStepsMaxNcv) + Workspace::need((size_t)sm_count()) + 256; }

int lanczos_gs_launch(double* V, int64_t ldv, int64_t n, int j, double* T, double* status, int beta_slot, void* scratch,
                      cudaStream_t stream) {
  TNPY_CHECK_ARG(V && T && status && scratch && n > 0 && j >= 0 && j < kStepsMaxNcv, "bad argument");
  const int sms = sm_count();
  int64_t want = (n + kThreads - 1) / kThreads;
  int ctas = (int)(want < sms ? want : sms);
  if (ctas > 8 * kMaxGatherLoads) ctas = 8 * kMaxGatherLoads;  // gather_partials reads that many partials per coefficient
  GsArgs a;
  a.V = V;
  a.ldv = ldv;
  a.n = n;
  a.j = j;
  a.T = T;
  a.status = status;
  a.beta_slot = beta_slot;
  Workspace ws(scratch, lanczos_gs_bytes());
  a.part = ws.take<double>((size_t)2 * sms * kStepsMaxNcv);
  a.nrm = ws.take<double>((size_t)sms);
  a.chunk = (int)((n + ctas - 1) / ctas);
  const int smem = a.chunk * (int)sizeof(double);
  TNPY_TRY(set_max_dynamic_smem(lanczos_gs_kernel, (int)((kGsMaxVector / 128 + 64) * sizeof(double))));
  void* args[] = {&a};
  TNPY_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lanczos_gs_kernel), dim3(ctas), dim3(kThreads), args, smem,
                                           stream));
  count_launch();
  return TNPY_OK;
}

int lanczos_steps_launch(const LanczosStepsPlan& plan, const double* L, const double* W, const double* R, double* V,
                         int64_t ldv, double* T, double* status, int beta_slot, int steps_slot, int l, int r, int wl,
                         int wr, int d, int j0, int nsteps, int ncv, int arrow, double tol, double anorm, void* scratch,
                         cudaStream_t stream) {
  TNPY_CHECK_ARG(plan.chunk <= kThreads && plan.grid >= 1, "vector too long for the fused path");
  TNPY_CHECK_ARG(ncv <= kStepsMaxNcv && j0 + nsteps <= ncv, "steps beyond the basis");
  const int64_t n = (int64_t)l * d * r;
  Workspace ws(scratch, plan.bytes);
  StepsArgs a;
  a.L = L;
  a.W = W;
  a.R = R;
  a.V = V;
  a.ldv = ldv;
  a.T = T;
  a.status = status;
  a.t1 = ws.take<double>((size_t)d * r * wl * l);
  a.ypart = ws.take<double>((size_t)plan.ksplit * n);
  a.wbuf = ws.take<double>((size_t)n);
  a.part = ws.take<double>((size_t)2 * plan.grid * kStepsMaxNcv);
  a.nrm = ws.take<double>((size_t)plan.grid);
  a.stop = ws.take<int>(16);
  if (!a.t1 || !a.ypart || !a.wbuf || !a.part || !a.nrm || !a.stop) {
    set_error("lanczos_steps: workspace too small");
    return TNPY_EWORKSPACE;
  }
  a.beta_slot = beta_slot;
  a.steps_slot = steps_slot;
  a.l = l;
  a.r = r;
  a.wl = wl;
  a.wr = wr;
  a.d = d;
  a.j0 = j0;
  a.nsteps = nsteps;
  a.ksplit = plan.ksplit;
  a.kchunk = plan.kchunk;
  a.chunk = plan.chunk;
  a.ldc = ((plan.chunk + 23) / 32) * 32 + 8;  // smallest value >= chunk that is 8 mod 32
  const int smem = (ncv + 1) * a.ldc * (int)sizeof(double);
  TNPY_TRY(set_max_dynamic_smem(lanczos_steps_kernel, (kStepsMaxNcv + 1) * (((kThreads + 23) / 32) * 32 + 8) * (int)sizeof(double)));
  a.trace = trace_buffer().load(std::memory_order_relaxed);
  a.arrow = arrow;
  a.tol = tol;
  a.anorm = anorm;
  void* args[] = {&a};
  TNPY_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lanczos_steps_kernel), dim3(plan.grid + 1), dim3(kThreads),
                                           args, smem, stream));
  count_launch();
  return TNPY_OK;
}

}  // namespace tnpy

extern "C" int tnpy_set_fused_steps(int on) { return tnpy::fused_steps_switch().exchange(on ? 1 : 0); }
extern "C" int tnpy_steps_trace(void* device_buffer) {
  tnpy::trace_buffer().store(static_cast<unsigned long long*>(device_buffer));
  return TNPY_OK;
}
