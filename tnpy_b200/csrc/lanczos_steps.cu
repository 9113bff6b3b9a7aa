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
// Every CTA owns a fixed range of vector elements (one per thread) through P3-P6, so w lives in a register and the
// basis columns a CTA reads in the Gram-Schmidt phases are the ones it wrote itself.  All sums run in a fixed order:
// results are bit-reproducible from run to run and from stream to stream.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "lanczos_steps.cuh"

namespace cg = cooperative_groups;

namespace tnpy {

namespace {
constexpr int kTile = 32;        // output tile edge
constexpr int kSlab = 32;        // K rows per shared-memory stage
constexpr int kThreads = 256;
constexpr int kMaxTerms = 1024;  // wl * wr * d * d of the MPO tensor (its non-zeros are a fraction of that)
constexpr int kMaxGroups = 128;  // wr * d
constexpr int64_t kMaxVector = 32768;
constexpr int kMaxKSplit = 8;

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
  int beta_slot, steps_slot;
  int l, r, wl, wr, d;
  int j0, nsteps;
  int ksplit, kchunk, chunk;
};

__device__ __forceinline__ void tile_mma(const double (*As)[kTile + 1], const double (*Bs)[kTile + 1], int tx, int ty,
                                         double (&acc)[2][2]) {
#pragma unroll 8
  for (int kk = 0; kk < kSlab; ++kk) {
    const double a0 = As[kk][ty], a1 = As[kk][ty + 16];
    const double b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
    acc[0][0] = fma(a0, b0, acc[0][0]);
    acc[0][1] = fma(a0, b1, acc[0][1]);
    acc[1][0] = fma(a1, b0, acc[1][0]);
    acc[1][1] = fma(a1, b1, acc[1][1]);
  }
}

// partial[k] = sum over this CTA's elements of V[k][.] * w[.] for k < m; warp `wv` takes k = wv, wv + 8, ...
__device__ __forceinline__ void partial_dots(const double* V, int64_t ldv, int m, int64_t first, int count,
                                             const double* wsm, double* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < m; k += kThreads / 32) {
    const double* vk = V + (int64_t)k * ldv + first;
    double s = 0.0;
    for (int e = lane; e < count; e += 32) s = fma(vk[e], wsm[e], s);
    s = warp_sum(s);
    if (lane == 0) out[k] = s;
  }
}

// h[k] = sum over the CTAs of their partials (fixed order), k < m
__device__ __forceinline__ void gather_partials(const double* part, int ctas, int m, double* h) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < m; k += kThreads / 32) {
    double s = 0.0;
    for (int c = lane; c < ctas; c += 32) s += part[(int64_t)c * kStepsMaxNcv + k];
    s = warp_sum(s);
    if (lane == 0) h[k] = s;
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
  const int ctas = gridDim.x, cta = blockIdx.x;

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

  const int64_t first = (int64_t)cta * a.chunk;
  const int count = first >= n ? 0 : (int)min((int64_t)a.chunk, n - first);
  const bool own = tid < count;
  double w = 0.0;
  const double* xsrc = a.V + (int64_t)a.j0 * a.ldv;
  double xscale = 1.0;
  double* part0 = a.part;
  double* part1 = a.part + (int64_t)ctas * kStepsMaxNcv;

  for (int step = 0; step < a.nsteps; ++step) {
    const int j = a.j0 + step, m = j + 1;

    // ---- P1 ----
    {
      const int tiles_n = (N1 + kTile - 1) / kTile, tiles = ((M1 + kTile - 1) / kTile) * tiles_n;
      for (int tile = cta; tile < tiles; tile += ctas) {
        const int m0 = (tile / tiles_n) * kTile, n0 = (tile % tiles_n) * kTile;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int k0 = 0; k0 < K1; k0 += kSlab) {
#pragma unroll
          for (int it = 0; it < kSlab * kTile / kThreads; ++it) {
            const int idx = tid + it * kThreads, kk = idx >> 5, c = idx & 31, k = k0 + kk;
            As[kk][c] = (k < K1 && m0 + c < M1) ? xsrc[(int64_t)k * M1 + m0 + c] * xscale : 0.0;
            Bs[kk][c] = (k < K1 && n0 + c < N1) ? a.L[(int64_t)k * N1 + n0 + c] : 0.0;
          }
          __syncthreads();
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
    grid.sync();

    // ---- P2 ----
    {
      const int tiles_n = (N2 + kTile - 1) / kTile, tiles = ((M2 + kTile - 1) / kTile) * tiles_n;
      const int items = tiles * a.ksplit;
      for (int item = cta; item < items; item += ctas) {
        const int ks = item % a.ksplit, tile = item / a.ksplit;
        const int m0 = (tile / tiles_n) * kTile, n0 = (tile % tiles_n) * kTile;
        const int kbeg = ks * a.kchunk, kend = min(K2, kbeg + a.kchunk);
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int k0 = kbeg; k0 < kend; k0 += kSlab) {
#pragma unroll
          for (int it = 0; it < kSlab * kTile / kThreads; ++it) {
            const int idx = tid + it * kThreads, kk = idx >> 5, c = idx & 31, k = k0 + kk, cc = m0 + c;
            double v = 0.0;
            if (k < kend && cc < M2) {
              const int rp = k / wr, b = k - rp * wr, q = cc / l, mm = cc - q * l;
              const int g = b * d + q;
              const double* src = a.t1 + (int64_t)rp * N1 + mm;
              for (int t = goff[g]; t < goff[g + 1]; ++t) v = fma(coef[t], src[toff[t]], v);
            }
            As[kk][c] = v;
            Bs[kk][c] = (k < kend && n0 + c < N2) ? a.R[(int64_t)k * r + n0 + c] : 0.0;
          }
          __syncthreads();
          tile_mma(As, Bs, tx, ty, acc);
          __syncthreads();
        }
        double* yp = a.ypart + (int64_t)ks * n;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int cc = m0 + ty + 16 * i, col = n0 + tx + 16 * jj;
            if (cc < M2 && col < N2) {
              const int q = cc / l, mm = cc - q * l;
              yp[((int64_t)mm * d + q) * r + col] = acc[i][jj];
            }
          }
      }
    }
    grid.sync();

    // ---- P3: w and the first pass' partial coefficients ----
    w = 0.0;
    if (own)
      for (int ks = 0; ks < a.ksplit; ++ks) w += a.ypart[(int64_t)ks * n + first + tid];
    wsm[tid] = w;
    __syncthreads();
    partial_dots(a.V, a.ldv, m, first, count, wsm, part0 + (int64_t)cta * kStepsMaxNcv);
    grid.sync();

    // ---- P4: first pass applied, second pass' partial coefficients ----
    gather_partials(part0, ctas, m, hs);
    __syncthreads();
    if (own) {
      const double* v = a.V + first + tid;
      for (int k = 0; k < m; ++k) w = fma(-hs[k], v[(int64_t)k * a.ldv], w);
    }
    wsm[tid] = w;
    __syncthreads();
    partial_dots(a.V, a.ldv, m, first, count, wsm, part1 + (int64_t)cta * kStepsMaxNcv);
    grid.sync();

    // ---- P5: second pass applied, partial norm, w published ----
    gather_partials(part1, ctas, m, h2s);
    __syncthreads();
    if (own) {
      const double* v = a.V + first + tid;
      for (int k = 0; k < m; ++k) w = fma(-h2s[k], v[(int64_t)k * a.ldv], w);
      a.wbuf[first + tid] = w;
    }
    {
      const double sq = block_sum(own ? w * w : 0.0, red);
      if (tid == 0) a.nrm[cta] = sq;
    }
    grid.sync();

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
    if (own) a.V[(int64_t)(j + 1) * a.ldv + first + tid] = w * inv;
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
    if (!(beta > 0.0)) break;  // exact breakdown: every CTA sees the same beta
    __syncthreads();           // beta_sh, hs, h2s are rewritten in the next step
  }
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
  return n <= kMaxVector && (n + kThreads - 1) / kThreads <= sm_count() && (int64_t)wl * wr * d * d <= kMaxTerms &&
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
  p.grid = want < sms ? want : sms;
  p.chunk = (int)((n + p.grid - 1) / p.grid);  // <= 256: grid >= n / 256 (n <= 32768 needs 128 CTAs, a B200 has 148)
  const int K2 = r * wr;
  int ksplit = p.grid / tiles2;
  if (ksplit > kMaxKSplit) ksplit = kMaxKSplit;
  if (ksplit > ceil_div(K2, kSlab)) ksplit = ceil_div(K2, kSlab);
  if (ksplit < 1) ksplit = 1;
  p.kchunk = ceil_div(ceil_div(K2, ksplit), kSlab) * kSlab;
  p.ksplit = ceil_div(K2, p.kchunk);
  p.bytes = Workspace::need((size_t)d * r * wl * l) + Workspace::need((size_t)p.ksplit * n) + Workspace::need((size_t)n) +
            Workspace::need((size_t)2 * p.grid * kStepsMaxNcv) + Workspace::need((size_t)p.grid) + 256;
  return p;
}

int lanczos_steps_launch(const LanczosStepsPlan& plan, const double* L, const double* W, const double* R, double* V,
                         int64_t ldv, double* T, double* status, int beta_slot, int steps_slot, int l, int r, int wl,
                         int wr, int d, int j0, int nsteps, void* scratch, cudaStream_t stream) {
  TNPY_CHECK_ARG(plan.chunk <= kThreads && plan.grid >= 1, "vector too long for the fused path");
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
  if (!a.t1 || !a.ypart || !a.wbuf || !a.part || !a.nrm) {
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
  void* args[] = {&a};
  TNPY_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lanczos_steps_kernel), dim3(plan.grid), dim3(kThreads),
                                           args, 0, stream));
  count_launch();
  return TNPY_OK;
}

}  // namespace tnpy

extern "C" int tnpy_set_fused_steps(int on) { return tnpy::fused_steps_switch().exchange(on ? 1 : 0); }
