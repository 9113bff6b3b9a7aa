// Lowest eigenpair of the projected matrix of thick-restart Lanczos in O(m) per evaluation.
#pragma once
#include "common.cuh"

namespace tnpy {

constexpr int kWatchThreads = 256;  // threads of the CTA that runs these (one Sturm count per thread and round)
constexpr int kWatchLd = 48;        // leading dimension of T (== kMaxNcv of lanczos.cu, kStepsMaxNcv)

// Between restarts the projected matrix is, up to rounding-level fill, T~ = diag(theta_0 .. theta_{k-1}) coupled to row k
// (the thick-restart arrow) followed by a tridiagonal tail: a tree (star + path), so every factorisation below is O(m)
// without fill.  One CTA finds the lowest eigenvalue of T~ by 256-section on Sturm counts (all threads, a handful of rounds)
// and the last component of its eigenvector from a twisted factorisation rooted where the eigenvector is largest
// (Parlett-Dhillon: the root with the smallest |gamma|), which stays accurate when the pair has converged and the top-down
// pivots are noise.  Two users: the watcher CTA of the fused small-site steps (csrc/lanczos_steps.cu), which only decides when a launch
// returns, and ritz_kernel (csrc/lanczos.cu) at looks that need nothing but the lowest pair, which checks the vector
// against the full T before it trusts it.
__device__ inline int sturm_count(const double* dg, const double* cp2, int k, int m, double s, double tiny) {
  // cp2 = squared couplings.  Negative pivots of T~ - s in the elimination order kept vectors, row k, tail: the kept
  // vectors' pivots directly, then sign changes of the tail's polynomial recurrence (no division on that chain)
  int cnt = 0;
  double hub = dg[k] - s;
  for (int i = 0; i < k; ++i) {
    double dd = dg[i] - s;
    if (dd == 0.0) dd = -tiny;
    cnt += dd < 0.0;
    hub -= cp2[i] / dd;
  }
  if (hub == 0.0) hub = -tiny;
  cnt += hub < 0.0;
  double qp = 1.0, q = hub;
  int i = k + 1;
  while (i < m) {
    const int end = min(m, i + 8);
    for (; i < end; ++i) {
      double qn = fma(dg[i] - s, q, -cp2[i - 1] * qp);
      if (qn == 0.0) qn = q < 0.0 ? tiny : -tiny;  // a zero takes the sign opposite to its predecessor
      cnt += (qn < 0.0) != (q < 0.0);
      qp = q;
      q = qn;
    }
    const double aq = fabs(q);  // rescaled every eight rows: off the dependency chain of the recurrence
    if (aq > 1e100) {
      q *= 1e-100;
      qp *= 1e-100;
    } else if (aq < 1e-100) {
      q *= 1e100;
      qp *= 1e100;
    }
  }
  return cnt;
}

// The estimate is spread over the barrier intervals of one step so that it rarely holds the other CTAs up:
// watch_begin + 1 round, then one round per interval, the rest and the eigenvector in watch_finish.  State lives in
// the watcher's shared memory (e: >= 640 doubles, ei: >= 16 ints); every thread of the CTA calls each stage.
struct WatchMem {
  double *dg, *cp, *cp2, *dsp, *rdsp, *dm, *rdm, *dp, *rdp, *gam, *dki, *z, *box;
  __device__ explicit WatchMem(double* e)
      : dg(e), cp(e + 48), cp2(e + 96), dsp(e + 144), rdsp(e + 192), dm(e + 240), rdm(e + 290), dp(e + 340),
        rdp(e + 388), gam(e + 436), dki(e + 484), z(e + 532), box(e + 580) {}
  // dg: diagonal of T~;  cp[i]: i < k coupling of kept vector i to row k, i >= k sub-diagonal T[i + 1][i];
  // dsp: pivots of the kept vectors as leaves;  dm: bottom-up pivots of the tail (49 entries);  dp: top-down pivots
  // from row k;  r*: their reciprocals;  gam: twist values;  dki: pivot of row k when kept vector i is the root;
  // box: 0 lo, 1 hi, 2 pivot floor, 3 spoke sum, 4 beta, 5 bracket tolerance, 6 / 7 Gershgorin lo / hi,
  //      8 / 9 / 10 the previous estimate's lo / hi / residual, 11 whether there is one, 12 size of T at which the
  //      next estimate is due, 13 size at the previous one
};
constexpr int kWatchBox = 580;   // offset of `box` in the watcher's memory
constexpr int kWatchDoubles = 600;  // shared memory the watcher needs (doubles); plus 16 ints

__device__ inline void watch_begin(const double* T, int m, int k, double beta, double* e) {
  WatchMem w(e);
  const int tid = threadIdx.x;
  if (tid < m) {
    w.dg[tid] = T[tid * kWatchLd + tid];
    const double c = tid < m - 1 ? (tid < k ? T[k * kWatchLd + tid] : T[(tid + 1) * kWatchLd + tid]) : 0.0;
    w.cp[tid] = c;
    w.cp2[tid] = c * c;
  }
  __syncthreads();
  if (tid == 0) {
    double lo = 1e300, hi = -1e300, big = 0.0, hubrad = 0.0;
    for (int i = 0; i < k; ++i) hubrad += fabs(w.cp[i]);
    for (int i = 0; i < m; ++i) {
      double rad;
      if (i < k) rad = fabs(w.cp[i]);
      else if (i == k) rad = hubrad + (k < m - 1 ? fabs(w.cp[k]) : 0.0);
      else rad = fabs(w.cp[i - 1]) + (i < m - 1 ? fabs(w.cp[i]) : 0.0);
      lo = fmin(lo, w.dg[i] - rad);
      hi = fmax(hi, w.dg[i] + rad);
      big = fmax(big, fabs(w.dg[i]));
    }
    lo -= 1e-3 * (hi - lo) + 1e-300;
    hi += 1e-3 * (hi - lo);
    const double scale = fmax(fmax(fabs(lo), fabs(hi)), fmax(big, 1e-280));
    w.box[6] = lo;
    w.box[7] = hi;
    if (w.box[11] != 0.0) {
      // the lowest Ritz value only falls as the basis grows, and by no more than the residual: start from the last
      // estimate (the rounds check the ends of the bracket and fall back to the Gershgorin interval)
      lo = fmax(lo, w.box[8] - fmax(4.0 * w.box[10], 1e-9 * scale));
      hi = fmin(hi, w.box[9] + 1e-12 * scale);
    }
    w.box[0] = lo;
    w.box[1] = hi;
    w.box[2] = 1e-18 * scale;
    w.box[4] = beta;
    w.box[5] = 1e-13 * scale;
  }
  __syncthreads();
}

// `target`: 1 brackets the lowest eigenvalue (first shift with an eigenvalue below it), m the highest (first shift
// with all of them below it)
__device__ inline void watch_rounds(int rounds, int m, int k, double* e, int* ei, int target = 1) {
  WatchMem w(e);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double tiny = w.box[2] * 1e-12;
  for (int round = 0; round < rounds; ++round) {
    const double lo = w.box[0], hi = w.box[1];
    if (hi - lo <= w.box[5]) break;
    // 256 shifts from lo to hi inclusive; the lowest eigenvalue lies between the last shift with no eigenvalue
    // below it and the first with one
    const double sigma = lo + (hi - lo) * (double)tid / (double)(kWatchThreads - 1);
    const int cnt = sturm_count(w.dg, w.cp2, k, m, sigma, tiny);
    const unsigned ball = __ballot_sync(0xffffffffu, cnt >= target);
    if (lane == 0) ei[warp] = (int)ball;
    __syncthreads();
    if (tid == 0) {
      int f = kWatchThreads;
      for (int wv = 0; wv < kWatchThreads / 32; ++wv)
        if (ei[wv] != 0) {
          f = wv * 32 + __ffs(ei[wv]) - 1;
          break;
        }
      if (f == 0) {  // already an eigenvalue below lo: the warm start was too high
        w.box[1] = lo;
        w.box[0] = w.box[6];
      } else if (f == kWatchThreads) {  // none below hi
        w.box[0] = hi;
        w.box[1] = w.box[7];
      } else {
        w.box[0] = lo + (hi - lo) * (double)(f - 1) / (double)(kWatchThreads - 1);
        w.box[1] = lo + (hi - lo) * (double)f / (double)(kWatchThreads - 1);
      }
    }
    __syncthreads();
  }
}

// z (w.z, normalised) = eigenvector of T~ for the eigenvalue bracketed in box[0], box[1]; every thread calls it
__device__ inline void watch_vector(int m, int k, double* e) {
  WatchMem w(e);
  const int tid = threadIdx.x;
  const double s = w.box[0];  // just below the lowest eigenvalue: T~ - s is positive definite, every pivot positive
  const double floor_ = w.box[2];
  auto guard = [&](double v) { return v > floor_ ? v : floor_; };
  if (tid < k) {
    const double dd = guard(w.dg[tid] - s);
    w.dsp[tid] = dd;
    w.rdsp[tid] = 1.0 / dd;
  }
  __syncthreads();
  if (tid == 0) {  // bottom-up along the tail
    double rnext = 0.0;
    for (int j = m - 1; j > k; --j) {
      const double dd = guard(w.dg[j] - s - w.cp2[j] * rnext);  // cp2[m - 1] == 0
      rnext = 1.0 / dd;
      w.dm[j] = dd;
      w.rdm[j] = rnext;
    }
  } else if (tid == 32) {  // top-down from row k
    double sum = 0.0;
    for (int i = 0; i < k; ++i) sum += w.cp2[i] * w.rdsp[i];
    w.box[3] = sum;
    double dd = guard(w.dg[k] - s - sum), rr = 1.0 / dd;
    w.dp[k] = dd;
    w.rdp[k] = rr;
    for (int j = k + 1; j < m; ++j) {
      dd = guard(w.dg[j] - s - w.cp2[j - 1] * rr);
      rr = 1.0 / dd;
      w.dp[j] = dd;
      w.rdp[j] = rr;
    }
  }
  __syncthreads();
  if (tid == 0) w.gam[k] = w.dg[k] - s - w.box[3] - (k < m - 1 ? w.cp2[k] * w.rdm[k + 1] : 0.0);
  __syncthreads();
  if (tid < m && tid != k) {
    if (tid < k) {
      w.dki[tid] = guard(w.gam[k] + w.cp2[tid] * w.rdsp[tid]);
      w.gam[tid] = w.dg[tid] - s - w.cp2[tid] / w.dki[tid];
    } else {
      w.gam[tid] = w.dg[tid] - s - w.cp2[tid - 1] * w.rdp[tid - 1] - (tid < m - 1 ? w.cp2[tid] * w.rdm[tid + 1] : 0.0);
    }
  }
  __syncthreads();
  if (tid == 0) {
    double* z = w.z;
    int root = 0;
    for (int i = 1; i < m; ++i)
      if (fabs(w.gam[i]) < fabs(w.gam[root])) root = i;
    for (int i = 0; i < m; ++i) z[i] = 0.0;
    z[root] = 1.0;
    int from = k;  // the tail is walked downwards from here
    if (root == k) {
      for (int i = 0; i < k; ++i) z[i] = -w.cp[i] * w.rdsp[i];
    } else if (root < k) {
      z[k] = -w.cp[root] / w.dki[root];
      for (int i = 0; i < k; ++i)
        if (i != root) z[i] = -w.cp[i] * z[k] * w.rdsp[i];
    } else {
      for (int j = root; j > k; --j) z[j - 1] = -w.cp[j - 1] * z[j] * w.rdp[j - 1];
      for (int i = 0; i < k; ++i) z[i] = -w.cp[i] * z[k] * w.rdsp[i];
      from = root;
    }
    for (int j = from; j < m - 1; ++j) z[j + 1] = -w.cp[j] * z[j] * w.rdm[j + 1];
    double nn = 0.0;
    for (int i = 0; i < m; ++i) nn = fma(z[i], z[i], nn);
    const double inv = 1.0 / sqrt(nn);
    for (int i = 0; i < m; ++i) z[i] *= inv;
  }
  __syncthreads();
}

}  // namespace tnpy
