// On-device lowest-eigenpair solver for H_eff (replaces primme.eigsh behind linalg.eigshmv,
// reference linalg.py:64-87, called from finite_dmrg.py:111).
//
// Thick-restart Lanczos with full reorthogonalisation (a local classical Gram-Schmidt pass against the vectors the
// three-term recurrence couples to, then a pass against the whole basis, a further one when the DGKS test asks) and
// an explicitly projected matrix T = V^T H V.  Everything -- the matvec chain, the vector kernels,
// the Rayleigh-Ritz solve (parallel cyclic Jacobi on T in one CTA) and the convergence test -- runs
// on the device; the host only reads one 64-byte status record per Lanczos step to decide whether to
// stop or restart.  Basis vectors never leave HBM.
#include <math.h>
#include <stdlib.h>

#include "comm.cuh"
#include "heff.cuh"
#include "lanczos_steps.cuh"
#include "ritz_watch.cuh"

namespace tnpy {

int multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h, int mode,
              cudaStream_t stream, const int* skip = nullptr);
int multi_axpy(const double* V, int64_t ldv, int m, const double* h, double* w, int64_t n, double* nrm_out,
               cudaStream_t stream, const int* skip = nullptr);
int scale_copy(const double* x, double* out, int64_t n, double alpha, const double* s_dev, int inv,
               cudaStream_t stream);
int axpy(double alpha, const double* a_dev, const double* x, double* y, int64_t n, cudaStream_t stream);
int combine(const double* V, int64_t ldv, int m, const double* c, int64_t ldc_, int nout, double* out, int64_t ldo,
            int64_t n, cudaStream_t stream);

constexpr int kMaxNcv = 48;
static_assert(kMaxNcv == kStepsMaxNcv && kMaxNcv == kWatchLd, "the fused small-site steps and the O(m) Ritz code read T with this leading dimension");
// leading dimension of the shared-memory matrices of the Ritz solve: odd, so that a column walk (the A <- A J phase
// of the Jacobi rounds, consecutive threads on consecutive rows) is spread over the banks.  With the natural 48 every
// row of a column sits in the same bank and a 30 x 30 Ritz problem took 1.5 ms (profiles/r02_launches_*).
constexpr int kRitzLd = kMaxNcv + 1;

// status record (device and pinned host mirror)
enum { ST_THETA = 0, ST_RESID = 1, ST_ANORM = 2, ST_DONE = 3, ST_BETA = 4, ST_RCOEF = 5, ST_BOUND = 6, ST_EXTRA = 7, ST_STEPS = 8, ST_TRUE = 9, ST_SIZE = 10 };

// Symmetric eigen-decomposition of the m x m matrix held in shared memory `a` (leading dim kMaxNcv)
// by parallel cyclic Jacobi (round-robin pair ordering).  Eigenvectors accumulate in `z` (columns).
__device__ void jacobi_eig_smem(double (*a)[kRitzLd], double (*z)[kRitzLd], int m, double* cs, double* sn, int* pp,
                                int* qq, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int idx = tid; idx < m * m; idx += nt) z[idx / m][idx % m] = (idx / m == idx % m) ? 1.0 : 0.0;
  __syncthreads();
  if (m == 1) return;
  const int me = (m + 1) & ~1;  // even player count (last one may be a bye)
  const int half = me / 2;
  double o_prev = 1e300;  // thread 0 only
  for (int sweep = 0; sweep < 40; ++sweep) {
    // convergence: off-diagonal mass vs diagonal mass
    double off = 0.0, dia = 0.0;
    for (int idx = tid; idx < m * m; idx += nt) {
      const int i = idx / m, j = idx % m;
      const double v = a[i][j];
      if (i == j) dia += v * v; else off += v * v;
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
      // round-robin tournament: player me-1 fixed, the others rotate
      if (tid < half) {
        int p = (tid == 0) ? me - 1 : (round + tid) % (me - 1);
        int q = (round + me - 1 - tid) % (me - 1);
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, s = 0.0;
        if (q < m) {
          const double apq = a[p][q];
          if (apq != 0.0) {
            const double app = a[p][p], aqq = a[q][q];
            if (fabs(apq) > 1e-300 && fabs(apq) >= 2.3e-16 * 1e-3 * sqrt(fabs(app * aqq)) ) {
              const double tau = (aqq - app) / (2.0 * apq);
              const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
              c = 1.0 / sqrt(1.0 + t * t);
              s = t * c;
            }
          }
        }
        cs[tid] = c; sn[tid] = s; pp[tid] = p; qq[tid] = q;
      }
      __syncthreads();
      // A <- J^T A J in one pass: work item (i, j) owns the 2 x 2 block A[{p_i, q_i}][{p_j, q_j}], which only the
      // rotations of pairs i (rows) and j (columns) touch; Z <- Z J: item (k, i) owns Z[k][{p_i, q_i}].
      for (int idx = tid; idx < half * half + m * half; idx += nt) {
        if (idx < half * half) {
          const int i = idx / half, j = idx % half;
          const int pi = pp[i], qi = qq[i], pj = pp[j], qj = qq[j];
          if (sn[i] == 0.0 && sn[j] == 0.0) continue;
          const bool ri = qi < m, rj = qj < m;  // a bye (q == m on an odd m) has no second row / column
          const double ci = cs[i], si = sn[i], cj = cs[j], sj = sn[j];
          double b00 = a[pi][pj], b01 = rj ? a[pi][qj] : 0.0, b10 = ri ? a[qi][pj] : 0.0, b11 = (ri && rj) ? a[qi][qj] : 0.0;
          // rows: J_i^T B
          const double r00 = ci * b00 - si * b10, r01 = ci * b01 - si * b11;
          const double r10 = si * b00 + ci * b10, r11 = si * b01 + ci * b11;
          // columns: (J_i^T B) J_j
          b00 = cj * r00 - sj * r01; b01 = sj * r00 + cj * r01;
          b10 = cj * r10 - sj * r11; b11 = sj * r10 + cj * r11;
          a[pi][pj] = b00;
          if (rj) a[pi][qj] = b01;
          if (ri) a[qi][pj] = b10;
          if (ri && rj) a[qi][qj] = b11;
        } else {
          const int e = idx - half * half, k = e / half, i = e % half;
          const int p = pp[i], q = qq[i];
          if (q >= m || sn[i] == 0.0) continue;
          const double c = cs[i], s = sn[i];
          const double zkp = z[k][p], zkq = z[k][q];
          z[k][p] = c * zkp - s * zkq;
          z[k][q] = s * zkp + c * zkq;
        }
      }
      __syncthreads();
    }
  }
}

// After Lanczos step j: fold h (+ h2) into column j of T, solve the (j+1) x (j+1) Ritz problem, write
// status, the sorted Ritz values `thetas` and the sorted Ritz coefficient matrix S (column i = i-th lowest).
// h == nullptr: T is complete already (the fused small-site steps write their columns themselves) and the last step
// done is j + *steps_done - 1, j being the first step of that launch.
__global__ void __launch_bounds__(256) ritz_kernel(double* __restrict__ T, const double* __restrict__ h,
                                                   const double* __restrict__ h2, const double* __restrict__ h_local,
                                                   int local_from, const double* __restrict__ beta_dev,
                                                   int j, double tol, double* __restrict__ S,
                                                   double* __restrict__ thetas, double* __restrict__ status,
                                                   int fold_only, const double* __restrict__ steps_done = nullptr,
                                                   int arrow = 0, int ncv = 0) {
  __shared__ double a[kMaxNcv][kRitzLd];
  __shared__ double z[kMaxNcv][kRitzLd];
  __shared__ double cs[kMaxNcv], sn[kMaxNcv], red[72];
  __shared__ int pp[kMaxNcv], qq[kMaxNcv], order[kMaxNcv];
  if (steps_done) j += (int)*steps_done - 1;
  const int m = j + 1;
  const int tid = threadIdx.x;
  if (h != nullptr && tid < m) {
    // column j of T = V^T H V: the local pass (vectors local_from .. j) plus the full pass(es)
    const double v = h[tid] + (h2 ? h2[tid] : 0.0) + (tid >= local_from ? h_local[tid - local_from] : 0.0);
    T[tid * kMaxNcv + j] = v;
    T[j * kMaxNcv + tid] = v;
  }
  if (fold_only) return;  // a step whose convergence is not looked at: T gets its column, nothing else
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += blockDim.x) a[idx / m][idx % m] = T[(idx / m) * kMaxNcv + idx % m];
  __syncthreads();
  // A look that is not a restart needs nothing but the lowest Ritz pair (the status record, column 0 of S).  T is, up to
  // rounding-level fill, diag(`arrow` kept Ritz values) + their coupling to row `arrow` + a tridiagonal tail, whose lowest
  // pair costs O(m) per evaluation (csrc/ritz_watch.cuh) instead of the ~8 Jacobi sweeps of 2 m barriers each below
  // (0.25 ms at m = 30).  The vector is then checked against the *full* T -- Rayleigh quotient and residual -- and only
  // trusted when it is an eigenvector of it to 1e-13 ||T|| (and 1 % of the solver's tolerance); otherwise, and on
  // restart looks (ncv > 0 && m == ncv: every kept Ritz vector is needed), the Jacobi solve runs.
  if (ncv > 0 && m < ncv && m >= 12 && arrow <= m - 1) {
    double* e = &z[0][0];
    int* ei = pp;
    WatchMem w(e);
    if (tid == 0) w.box[11] = 0.0;  // no previous estimate to start from
    __syncthreads();
    const double beta = *beta_dev;
    watch_begin(T, m, arrow, beta, e);
    if (tid == 0) w.box[5] = 4e-16 * (w.box[2] * 1e18);  // bracket to working precision (box[2] = 1e-18 scale)
    __syncthreads();
    watch_rounds(10, m, arrow, e, ei, 1);
    watch_vector(m, arrow, e);
    double ti = 0.0;
    if (tid < m)
      for (int k = 0; k < m; ++k) ti = fma(a[tid][k], w.z[k], ti);
    double s1 = block_sum(tid < m ? w.z[tid] * ti : 0.0, red);
    if (tid == 0) red[64] = s1;
    __syncthreads();
    const double theta = red[64];
    const double ri = tid < m ? ti - theta * w.z[tid] : 0.0;
    double s2 = block_sum(ri * ri, red);
    // the largest eigenvalue, coarsely (it only scales the stopping rule): three rounds from [max diagonal, Gershgorin]
    if (tid == 0) {
      red[65] = sqrt(s2);
      double top = w.dg[0];
      for (int i = 1; i < m; ++i) top = fmax(top, w.dg[i]);
      w.box[0] = top;
      w.box[1] = w.box[7];
      w.box[5] = 0.0;
    }
    __syncthreads();
    watch_rounds(3, m, arrow, e, ei, m);
    const double anorm = fmax(fabs(theta), fabs(w.box[0]));
    const double rn = red[65];
    const bool trusted = rn <= fmin(1e-13, 1e-2 * tol) * anorm;
    if (trusted) {
      const double zl = w.z[m - 1];
      const double sgn = w.z[0] < 0.0 ? -1.0 : 1.0;
      if (tid < m) S[tid * kMaxNcv + 0] = sgn * w.z[tid];
      if (tid == 0) {
        const double resid = fabs(beta * zl);
        thetas[0] = theta;
        status[ST_THETA] = theta;
        status[ST_RESID] = resid;
        status[ST_ANORM] = anorm;
        status[ST_BETA] = beta;
        status[ST_DONE] = (resid <= tol * anorm) ? 1.0 : 0.0;
        status[ST_RCOEF] = beta * sgn * zl;
      }
      return;
    }
    __syncthreads();
  }
  jacobi_eig_smem(a, z, m, cs, sn, pp, qq, red);
  if (tid == 0) {
    // sort eigenvalues ascending (stable selection; m <= 48)
    for (int i = 0; i < m; ++i) order[i] = i;
    for (int i = 0; i < m; ++i) {
      int best = i;
      for (int k = i + 1; k < m; ++k)
        if (a[order[k]][order[k]] < a[order[best]][order[best]]) best = k;
      const int t = order[i]; order[i] = order[best]; order[best] = t;
    }
    double anorm = 0.0;
    for (int i = 0; i < m; ++i) anorm = fmax(anorm, fabs(a[i][i]));
    const int lo = order[0];
    const double beta = *beta_dev;
    const double resid = fabs(beta * z[m - 1][lo]);
    status[ST_THETA] = a[lo][lo];
    status[ST_RESID] = resid;
    status[ST_ANORM] = anorm;
    status[ST_BETA] = beta;
    status[ST_DONE] = (resid <= tol * anorm) ? 1.0 : 0.0;
    // signed coefficient of v_{m+1} in H x - theta x for the sign convention S is written in below
    status[ST_RCOEF] = beta * (z[0][lo] < 0.0 ? -z[m - 1][lo] : z[m - 1][lo]);
  }
  __syncthreads();
  for (int idx = tid; idx < m * m; idx += blockDim.x) {
    const int k = idx / m, i = idx % m;
    const int col = order[i];
    // fix the sign of every Ritz vector so that its first basis component is non-negative
    const double sgn = z[0][col] < 0.0 ? -1.0 : 1.0;
    S[k * kMaxNcv + i] = sgn * z[k][col];
  }
  if (tid < m) thetas[tid] = a[order[tid]][order[tid]];
}

__global__ void restart_T_kernel(double* __restrict__ T, const double* __restrict__ thetas, int keep) {
  for (int idx = threadIdx.x; idx < kMaxNcv * kMaxNcv; idx += blockDim.x) {
    const int i = idx / kMaxNcv, k = idx % kMaxNcv;
    T[idx] = (i == k && i < keep) ? thetas[i] : 0.0;
  }
}

// process-wide switch of the inexact-Krylov slice schedule (tnpy_set_inexact_slices / TNPY_INEXACT_SLICES=0)
static std::atomic<int>& inexact_switch() {
  static std::atomic<int> on([] {
    const char* env = getenv("TNPY_INEXACT_SLICES");
    return (env && env[0] == '0') ? 0 : 1;
  }());
  return on;
}

// safety factor of the schedule: bound(S) <= factor tol ||A|| / (||r|| / ||A||); TNPY_INEXACT_FACTOR overrides it (experiments)
static double inexact_factor() {
  static const double f = [] {
    const char* env = getenv("TNPY_INEXACT_FACTOR");
    const double v = env ? atof(env) : 0.0;
    return v > 0.0 ? v : 0.00125;
  }();
  return f;
}

static double* pinned_status() { return static_cast<double*>(thread_pinned_scratch()); }

static size_t eig_ws_layout(int64_t n, int ncv, int keep, size_t chain) {
  const int64_t ldv = n + (n & 1);
  size_t total = 0;
  total += Workspace::need((size_t)(ncv + 1) * ldv);  // V
  total += Workspace::need((size_t)keep * ldv);       // Y
  total += Workspace::need(kMaxNcv * kMaxNcv) * 2;    // T, S
  total += Workspace::need(64) * 5;                   // thetas, h, h2, h_local, status
  total += lanczos_gs_bytes() + 256;                  // partials of the fused Gram-Schmidt launch
  total += chain + 512;
  return total;
}

// skip = 1 when the first Gram-Schmidt pass did not cancel heavily (||w'|| >= eta ||h||); h2 is then zero.
__global__ void reorth_decision_kernel(const double* __restrict__ h, int m, const double* __restrict__ nrm_after,
                                       double eta, double* __restrict__ h2, int* __restrict__ skip,
                                       double* __restrict__ extra_passes) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int j = threadIdx.x; j < m; j += blockDim.x) a = fma(h[j], h[j], a);
  a = block_sum(a, sh);
  __shared__ int decision;
  if (threadIdx.x == 0) {
    const double hn = sqrt(a), wn = *nrm_after;
    decision = (wn >= eta * hn && wn > 0.0) ? 1 : 0;
    *skip = decision;
    if (!decision) *extra_passes += 1.0;  // diagnostics: how often the extra full pass really ran
  }
  __syncthreads();
  if (decision)
    for (int j = threadIdx.x; j < m; j += blockDim.x) h2[j] = 0.0;
}

// Row-sharded solve: a norm is the square root of a sum over the ranks.  pre: *sq = the local norm squared -- or, when
// this step's pass was skipped on the device (*skip != 0) and *norm is therefore already the global norm, *norm^2 /
// world so that the all-reduce reproduces it; post: *norm = sqrt(*sq).
__global__ void norm_pre_reduce_kernel(const double* __restrict__ local_norm, const double* __restrict__ norm,
                                       const int* __restrict__ skip, int world, double* __restrict__ sq) {
  const bool skipped = skip != nullptr && *skip != 0;
  *sq = skipped ? (*norm * *norm) / world : (*local_norm * *local_norm);
}
__global__ void norm_post_reduce_kernel(const double* __restrict__ sq, double* __restrict__ norm) { *norm = sqrt(*sq); }

static void pick_sizes(int64_t n, int ncv_in, int& ncv, int& keep) {
  // default basis of 32: on hard local problems (first sweeps from a random MPS) thick restart with
  // (32, 10) needs ~20 % fewer matvecs than (20, 6) and is within ~10 % of unrestarted Lanczos, while the
  // extra reorthogonalisation traffic stays well below the cost of one matvec
  ncv = ncv_in <= 0 ? 32 : ncv_in;
  if (ncv > kMaxNcv) ncv = kMaxNcv;
  if (ncv < 3) ncv = 3;
  if (ncv > n) ncv = (int)n;
  keep = ncv / 3;
  if (keep < 1) keep = 1;
  if (keep > 12) keep = 12;
}

}  // namespace tnpy

using namespace tnpy;

extern "C" size_t tnpy_eig_workspace_bytes(int l, int r, int wl, int wr, int d, int ncv_in) {
  int ncv, keep;
  const int64_t n = (int64_t)l * d * r;
  pick_sizes(n, ncv_in, ncv, keep);
  size_t apply = heff_apply_bytes(l, l, r, wl, wr, d);
  if (lanczos_steps_supported(l, r, wl, wr, d)) {
    const size_t fused = lanczos_steps_plan(l, r, wl, wr, d).bytes;
    if (fused > apply) apply = fused;  // the fused small-site steps use the chain's part of the workspace instead
  }
  return eig_ws_layout(n, ncv, keep, heff_plan_bytes(l, l, r, wl, wr, d) + apply + 1024);
}

// comm == nullptr: the whole problem on this GPU (lo == l, row0 == 0).  Otherwise rank g of the communicator holds
// the bra rows [row0, row0 + lo) of the left bond: L = L_full[:, :, rows] as (l, wl, lo), the same rows of psi / hpsi
// and of every Lanczos vector, and full copies of W and R.  Per step the ranks all-gather the current vector (the
// one exchange of data: (G - 1) / G of 8 N bytes per rank over NVLink), run their row block of the matvec, and
// all-reduce the Gram-Schmidt coefficients and norms (a few dozen doubles); the small Ritz problem is solved
// redundantly on every rank from identical inputs, so all ranks take identical decisions.
// diagnostics of the calling thread's last solve: matvecs, looks (status read-backs), extra full Gram-Schmidt passes,
// restarts, matvecs that ran with fewer int8 slices than the solve's base count, failed true-residual checks, matvecs
// with five slices
static thread_local long long g_last_counters[7] = {0, 0, 0, 0, 0, 0, 0};

static int eig_lowest_impl(const tnpy_comm* comm, const double* L, const double* W, const double* R, double* psi,
                           double* hpsi, int l, int row0, int lo, int r, int wl, int wr, int d, int flags, double tol,
                           int max_matvec, int ncv_in, double* stats_host, void* workspace, size_t workspace_bytes,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(psi && W, "null pointer");
  TNPY_CHECK_ARG(l > 0 && r > 0 && wl > 0 && wr > 0 && d > 0, "non-positive dimension");
  TNPY_CHECK_ARG(lo > 0 && row0 >= 0 && row0 + lo <= l, "row block outside the left bond");
  TNPY_CHECK_ARG(comm != nullptr || lo == l, "a row block needs a communicator");
  TNPY_CHECK_ARG(comm == nullptr || (lo * comm->world == l && row0 == comm->rank * lo), "rows must be split evenly, rank g holding block g");
  const int64_t n = (int64_t)lo * d * r;        // local vector length
  const int64_t n_full = (int64_t)l * d * r;    // global vector length
  const int64_t ldv = n + (n & 1);
  int ncv, keep;
  pick_sizes(n_full, ncv_in, ncv, keep);
  if (tol <= 0.0) tol = 2.220446049250313e-16 * 1e4;
  if (max_matvec <= 0) max_matvec = 1000;
  // sum over the ranks of `count` doubles in place / of a squared norm (no-ops on a single GPU)
  auto reduce = [&](double* buf, int count) -> int {
    return comm ? comm_allreduce_sum(comm, buf, (size_t)count, stream) : TNPY_OK;
  };

  Workspace ws(workspace, workspace_bytes);
  double* x_full = comm ? ws.take<double>((size_t)n_full) : nullptr;  // the all-gathered current vector
  double* V = ws.take<double>((size_t)(ncv + 1) * ldv);
  double* Y = ws.take<double>((size_t)keep * ldv);
  double* T = ws.take<double>(kMaxNcv * kMaxNcv);
  double* S = ws.take<double>(kMaxNcv * kMaxNcv);
  double* thetas = ws.take<double>(64);
  double* h = ws.take<double>(64);
  double* h2 = ws.take<double>(64);
  double* h_local = ws.take<double>(64);
  double* status = ws.take<double>(64);
  int* skip2 = reinterpret_cast<int*>(status + 48);  // device flag: skip the second Gram-Schmidt pass of this step
  char* gs_mem = ws.take<char>(lanczos_gs_bytes());
  if (!V || !Y || !T || !S || !thetas || !h || !h2 || !h_local || !status || !gs_mem || (comm && !x_full)) {
    set_error("tnpy_eig_lowest: workspace too small (%zu bytes given)", workspace_bytes);
    return TNPY_EWORKSPACE;
  }
  double* hst = pinned_status();
  if (!hst) {
    set_error("tnpy_eig_lowest: pinned status allocation failed");
    return TNPY_ECUDA;
  }
  // L, W, R are constant for the whole solve: everything that depends only on them (tcgen05 path: the int8 slices
  // of the environments) is prepared once, in this call's workspace.  A matvec whose rigorous error bound with 7
  // slices is orders of magnitude below the residual threshold does not need the eighth (28 instead of 36 slice
  // GEMMs); the bound is read back with every status record and the slice count raised -- or the solve moved to
  // the native FP64 chain -- if it ever comes within 1 % of tol * ||A||.
  const size_t plan_bytes = heff_plan_bytes(l, lo, r, wl, wr, d);  // enough for any mode: a later re-plan fits too
  char* plan_mem = ws.take<char>(plan_bytes);
  if (!plan_mem) {
    set_error("tnpy_eig_lowest: workspace too small (%zu bytes given)", workspace_bytes);
    return TNPY_EWORKSPACE;
  }
  HeffPlan plan;
  {
    Workspace mem(plan_mem, plan_bytes);
    TNPY_TRY(heff_plan_init(&plan, L, W, R, nullptr, l, lo, row0, r, wl, wr, d, flags, TNPY_GEMM_AUTO, mem, stream));
  }
  int slices = (tol >= 1e-10 && ozaki_slices() == 8) ? 7 : ozaki_slices();
  // Inexact Krylov (Simoncini & Szyld; Bouras & Fraysse): the matvec error a Lanczos step tolerates grows like
  // 1 / ||r|| of the current Ritz pair, because later basis vectors enter the converged Ritz vector with ever smaller
  // weights.  On the tcgen05 path that is fewer int8 slices for the later steps of a solve -- 21 or 15 slice-pair
  // GEMMs instead of 28 -- chosen at every look from the *rigorous* bound of the products (it scales with e_S):
  //     bound(S) <= (0.01 / 8) tol ||A|| / (||r|| / ||A||).
  // A solve that used fewer slices than `slices` is not trusted on its Lanczos residual: the true residual
  // H psi - theta psi is formed with one matvec at full accuracy (it is also the image the sweep wants next), and a
  // solve that fails it continues from psi with the schedule switched off.  TNPY_INEXACT_SLICES=0 switches it off.
  bool inexact = inexact_switch().load(std::memory_order_relaxed) != 0 && !comm && tol >= 1e-10;
  bool used_inexact = false;
  long long n_reduced = 0, n_five = 0, n_failed_checks = 0;
  int cur_slices = slices, last_used_slices = slices;
  double unit_bound = 0.0;  // measured bound / e_S: the scale-free part
  auto e_of = [](int S) { return (S + 2) / 4.0 * ldexp(1.0, -7 * S); };
  const size_t chain_off = ws.used;
  // Small sites: whole steps in one cooperative launch (csrc/lanczos_steps.cu); the host then only launches the
  // Ritz solve and reads the status record every `stride` steps.
  bool fused = !comm && plan.mode == HEFF_FP64_CHAIN && lanczos_steps_supported(l, r, wl, wr, d);
  // Mid-size sites keep the GEMM kernels for the matvec and run the rest of a step -- both Gram-Schmidt passes, the
  // norm, the normalised copy, the new column of T -- in one cooperative launch instead of nine small ones.
  bool gs_fused = !comm && !fused && lanczos_gs_supported(n_full);
  LanczosStepsPlan steps_plan{};
  if (fused) {
    steps_plan = lanczos_steps_plan(l, r, wl, wr, d);
    // a workspace sized while the fused path was switched off only holds the general solver's scratch
    if (workspace_bytes < chain_off + steps_plan.bytes) fused = false;
    if (!fused) gs_fused = lanczos_gs_supported(n_full);
  }
  // global norm of the vector whose local norm multi_dot / multi_axpy just left in *local (see the kernels above)
  double* sq = status + 40;
  auto reduce_norm = [&](const double* local, double* norm, const int* skip) -> int {
    if (!comm) return TNPY_OK;
    norm_pre_reduce_kernel<<<1, 1, 0, stream>>>(local, norm, skip, comm->world, sq);
    TNPY_LAUNCH_OK();
    TNPY_TRY(comm_allreduce_sum(comm, sq, 1, stream));
    norm_post_reduce_kernel<<<1, 1, 0, stream>>>(sq, norm);
    TNPY_LAUNCH_OK();
    return TNPY_OK;
  };
  TNPY_CUDA_OK(cudaMemsetAsync(status, 0, sizeof(double) * 64, stream));
  // V[0] = v0 / ||v0||
  TNPY_TRY(multi_dot(psi, ldv, 1, psi, n, status + ST_BETA, 1, stream));
  TNPY_TRY(reduce_norm(status + ST_BETA, status + ST_BETA, nullptr));
  TNPY_TRY(scale_copy(psi, V, n, 1.0, status + ST_BETA, 1, stream));
  TNPY_CUDA_OK(cudaMemsetAsync(T, 0, sizeof(double) * kMaxNcv * kMaxNcv, stream));

  int j = 0, n_matvec = 0, n_restart = 0;
  int n_looks = 0;
  int whole_basis_step = 0;  // step whose local Gram-Schmidt set is the whole basis (the first after a restart)
  double worst_bound = 0.0, anorm_seen = 0.0;
  bool done = false;
  // Daniel-Gragg-Kaufman-Stewart: the second pass is skipped only when ||w'|| >= ||w|| / sqrt 2, i.e.
  // ||w'|| >= ||h|| (||w||^2 = ||h||^2 + ||w'||^2).  A looser, tolerance-tied threshold (1e-3) was measured to
  // derail cold-sweep solves: spurious Ritz values of order 1e3 and 1000 wasted matvecs at 5 of 34 sites of
  // XXZ n=40 chi=512 (profiles/r01_sweep_trace_cold_chi512_*.jsonl).
  const double eta = 1.0;
  // The Ritz problem (a Jacobi eigensolve of T in one CTA, 0.05-0.2 ms) and the host read-back are only needed when
  // somebody looks at the result: on restart steps, at the matvec limit, and every `stride` steps.  The stride comes
  // from the residual history: Lanczos residuals fall geometrically, so the two last looks give a rate and with it
  // the number of steps left to the threshold; the next look happens after half of those, at most `stride_cap` steps
  // on (3 where a matvec costs milliseconds -- an overshoot step is then dearer than a look -- 8 below 2^20 unknowns,
  // where a look costs as much as one to three steps: at (256, 2, 256) a step is 0.2 ms, a Ritz solve at m = 30 0.25 ms).
  // The stopping rule itself is unchanged; a local solve can overshoot by at most stride_cap - 1 matvecs.
  const int stride_cap = n_full >= (1 << 20) ? 3 : 8;
  static const bool fast_ritz = [] {  // TNPY_FAST_RITZ=0: every look runs the Jacobi solve
    const char* env = getenv("TNPY_FAST_RITZ");
    return !(env && env[0] == '0');
  }();
  double last_resid = 0.0, last_anorm = 0.0;
  int last_look_matvec = 0;
  int since_check = 0, stride = 1;
  while (true) {
    int m;  // size of the Ritz problem looked at below
    if (fused) {
      // The launch runs to the next restart and returns by itself two steps after the device-side estimate of the
      // residual has reached 0.7 tol ||A|| (||A|| >= the last look's max |Ritz value| and |theta| of the estimate
      // itself; the margin covers estimates of ||A|| that shrink across a restart); the stopping rule proper is still
      // ritz_kernel's, on the full T.
      int nsteps = ncv - j;
      if (nsteps > max_matvec - n_matvec) nsteps = max_matvec - n_matvec;
      if ((int64_t)nsteps > n_full - j) nsteps = (int)(n_full - j);
      if (nsteps < 1) nsteps = 1;
      if (lanczos_steps_launch(steps_plan, plan.L, plan.W, plan.R, V, ldv, T, status, ST_BETA, ST_STEPS, l, r, wl, wr, d, j, nsteps,
                               ncv, whole_basis_step, tol, last_anorm, static_cast<char*>(workspace) + chain_off,
                               stream) != TNPY_OK) {
        // a cooperative launch the device cannot place right now (other contexts holding SMs): nothing has run,
        // the general multi-kernel solver continues from the same state
        cudaGetLastError();
        fused = false;
        continue;
      }
      ritz_kernel<<<1, 256, 0, stream>>>(T, nullptr, nullptr, nullptr, 0, status + ST_BETA, j, tol, S, thetas, status, 0,
                                         status + ST_STEPS, whole_basis_step, fast_ritz ? ncv : 0);
      TNPY_LAUNCH_OK();
      ++n_looks;
      TNPY_CUDA_OK(cudaMemcpyAsync(hst, status, sizeof(double) * ST_SIZE, cudaMemcpyDeviceToHost, stream));
      TNPY_CUDA_OK(cudaStreamSynchronize(stream));
      const int steps_done = (int)hst[ST_STEPS];  // fewer than asked: stopped by itself, or an exact breakdown
      last_anorm = hst[ST_ANORM];
      n_matvec += steps_done;
      j += steps_done - 1;
      m = j + 1;
      goto looked;
    }
    {
    double* vj = V + (int64_t)j * ldv;
    double* w = V + (int64_t)(j + 1) * ldv;
    Workspace chain(static_cast<char*>(workspace) + chain_off, workspace_bytes - chain_off);
    if (comm) TNPY_TRY(comm_allgather(comm, vj, x_full, (size_t)n, stream));
    const int step_slices = cur_slices < slices ? cur_slices : slices;
    TNPY_TRY(heff_plan_apply(plan, comm ? x_full : vj, w, step_slices, nullptr, chain, stream));
    if (plan.mode != HEFF_FP64_CHAIN) {
      if (step_slices < slices) {
        used_inexact = true;
        ++n_reduced;
        if (step_slices == 5) ++n_five;
      }
      last_used_slices = step_slices;
    }
    ++n_matvec;
    // Gram-Schmidt in two stages.  H v_j has analytically non-zero components only on v_{j-1} and v_j
    // (three-term recurrence; on every kept Ritz vector in the first step after a thick restart), so a *local*
    // classical Gram-Schmidt pass against just those removes everything large; the pass against the whole basis
    // that follows then sees a vector whose components along V are at rounding / loss-of-orthogonality level, i.e.
    // it is the second pass of "twice is enough" at the cost of one.  The Daniel-Gragg-Kaufman-Stewart test
    // (||w'|| >= ||w|| / sqrt 2 for the full pass, decided on the device) still guards it: a third, full pass runs
    // when the full pass cancelled after all.  A plain single full pass is not enough: ||w'|| / ||w|| is ~0.6 in
    // every Lanczos step, and each unguarded step multiplies the basis' orthogonality defect by ~1.3
    // (docs/experiments/local_solver_study.py); the coefficients of all passes add up to column j of T = V^T H V.
    bool gs_done = false;
    if (gs_fused) {
      if (lanczos_gs_launch(V, ldv, n, j, T, status, ST_BETA, gs_mem, stream) == TNPY_OK) {
        gs_done = true;
      } else {  // the device cannot place the cooperative launch: the separate kernels below do the same work
        cudaGetLastError();
        gs_fused = false;
      }
    }
    const int local_from = (j == whole_basis_step) ? 0 : (j > 0 ? j - 1 : 0);
    const int n_local = j + 1 - local_from;
    double* v_local = V + (int64_t)local_from * ldv;
    double* local_norm = comm ? status + 41 : status + ST_BETA;  // sharded: multi_axpy leaves the *local* norm here
    if (!gs_done) {
    TNPY_TRY(multi_dot(v_local, ldv, n_local, w, n, h_local, 0, stream));
    TNPY_TRY(reduce(h_local, n_local));
    TNPY_TRY(multi_axpy(v_local, ldv, n_local, h_local, w, n, nullptr, stream));
    TNPY_TRY(multi_dot(V, ldv, j + 1, w, n, h, 0, stream));
    TNPY_TRY(reduce(h, j + 1));
    TNPY_TRY(multi_axpy(V, ldv, j + 1, h, w, n, local_norm, stream));
    TNPY_TRY(reduce_norm(local_norm, status + ST_BETA, nullptr));
    reorth_decision_kernel<<<1, 64, 0, stream>>>(h, j + 1, status + ST_BETA, eta, h2, skip2, status + ST_EXTRA);
    TNPY_LAUNCH_OK();
    TNPY_TRY(multi_dot(V, ldv, j + 1, w, n, h2, 0, stream, skip2));
    TNPY_TRY(reduce(h2, j + 1));  // skipped pass: the decision kernel zeroed h2 on every rank
    TNPY_TRY(multi_axpy(V, ldv, j + 1, h2, w, n, local_norm, stream, skip2));
    TNPY_TRY(reduce_norm(local_norm, status + ST_BETA, skip2));
    }
    m = j + 1;
    ++since_check;
    const bool look = m == ncv || n_matvec >= max_matvec || m >= n_full || since_check >= stride;
    if (gs_done) {  // T has its column and w is normalised already: the Ritz kernel only runs when somebody looks
      if (look) {
        ritz_kernel<<<1, 256, 0, stream>>>(T, nullptr, nullptr, nullptr, 0, status + ST_BETA, j, tol, S, thetas, status, 0, nullptr,
                                           whole_basis_step, fast_ritz ? ncv : 0);
        TNPY_LAUNCH_OK();
      }
    } else {
      ritz_kernel<<<1, 256, 0, stream>>>(T, h, h2, h_local, local_from, status + ST_BETA, j, tol, S, thetas, status, look ? 0 : 1,
                                         nullptr, whole_basis_step, fast_ritz ? ncv : 0);
      TNPY_LAUNCH_OK();
      TNPY_TRY(scale_copy(w, w, n, 1.0, status + ST_BETA, 1, stream));
    }
    if (!look) {
      ++j;
      continue;
    }
    since_check = 0;
    ++n_looks;
    if (plan.bound)
      TNPY_CUDA_OK(cudaMemcpyAsync(status + ST_BOUND, plan.bound, sizeof(double), cudaMemcpyDeviceToDevice, stream));
    TNPY_CUDA_OK(cudaMemcpyAsync(hst, status, sizeof(double) * ST_SIZE, cudaMemcpyDeviceToHost, stream));
    TNPY_CUDA_OK(cudaStreamSynchronize(stream));
    }
  looked:
    // ||A|| for this test: the largest |Ritz value| or Lanczos coefficient seen so far (a lower bound of ||A|| that
    // is already tight after a few steps; the very first Ritz values of a random start can be near zero)
    anorm_seen = fmax(anorm_seen, fmax(hst[ST_ANORM], hst[ST_BETA]));
    // the measured bound belongs to the slice count of the products since the last look; its scale-free part
    // predicts the bound of any other count
    if (plan.mode != HEFF_FP64_CHAIN && hst[ST_BOUND] > 0.0) unit_bound = hst[ST_BOUND] / e_of(last_used_slices);
    if (plan.mode != HEFF_FP64_CHAIN && n_matvec >= 3 && unit_bound * e_of(slices) > 0.01 * tol * anorm_seen) {
      // the int8 products are no longer safely below the residual threshold: spend the eighth slice, then leave
      // the tcgen05 path altogether (the basis built so far stays valid: its vectors are exact matvecs to within
      // the bound, and T is the explicit projection)
      worst_bound = hst[ST_BOUND];
      if (slices < kOzMaxSlices) {
        slices = kOzMaxSlices;
      } else {
        Workspace again(plan_mem, plan_bytes);
        TNPY_TRY(heff_plan_init(&plan, L, W, R, nullptr, l, lo, row0, r, wl, wr, d, flags, TNPY_GEMM_FP64, again, stream));
      }
      if (plan.bound) TNPY_CUDA_OK(cudaMemsetAsync(plan.bound, 0, sizeof(double), stream));
      unit_bound = 0.0;
      cur_slices = slices;
    }
    if (inexact && plan.mode != HEFF_FP64_CHAIN && n_matvec >= 3 && unit_bound > 0.0 && anorm_seen > 0.0) {
      const double rel = fmin(1.0, fmax(hst[ST_RESID] / anorm_seen, tol));
      const double allowed = inexact_factor() * tol * anorm_seen / rel;
      int pick = slices;
      while (pick > 5 && unit_bound * e_of(pick - 1) <= allowed) --pick;
      if (pick != cur_slices) {
        cur_slices = pick;
        if (plan.bound) TNPY_CUDA_OK(cudaMemsetAsync(plan.bound, 0, sizeof(double), stream));  // next reading: this count only
      }
    }
    {
      const double thr = tol * hst[ST_ANORM], res = hst[ST_RESID];
      int guess = res > 1e4 * thr ? 3 : (res > 1e2 * thr ? 2 : 1);  // no history yet: by the distance alone
      if (last_resid > 0.0 && res > 0.0 && res < last_resid && thr > 0.0 && res > thr) {
        const double rate = log(last_resid / res) / (double)(n_matvec - last_look_matvec);  // decades (e-folds) per step
        const double left = log(res / thr) / rate;
        guess = left < 2.0 ? 1 : (int)(0.5 * left);
      }
      stride = guess < 1 ? 1 : (guess > stride_cap ? stride_cap : guess);
      last_resid = res;
      last_look_matvec = n_matvec;
    }
    done = hst[ST_DONE] != 0.0 || !(hst[ST_BETA] > 0.0) || m >= n_full;
    if (done || n_matvec >= max_matvec) {
      // psi = V[0..m-1] . S[:, 0]
      TNPY_TRY(combine(V, ldv, m, S, kMaxNcv, 1, psi, ldv, n, stream));
      if (used_inexact && done && hst[ST_BETA] > 0.0 && m < n_full) {
        // some steps ran with fewer slices: the Lanczos residual is an estimate, the true one decides.  Y = H psi at
        // full accuracy, V[0] (the basis is spent either way) = Y - theta psi.
        Workspace chain(static_cast<char*>(workspace) + chain_off, workspace_bytes - chain_off);
        TNPY_TRY(heff_plan_apply(plan, psi, Y, slices, nullptr, chain, stream));
        ++n_matvec;
        TNPY_TRY(scale_copy(Y, V, n, 1.0, nullptr, 0, stream));
        TNPY_TRY(axpy(-hst[ST_THETA], nullptr, psi, V, n, stream));
        TNPY_TRY(multi_dot(V, ldv, 1, V, n, status + ST_TRUE, 1, stream));
        TNPY_CUDA_OK(cudaMemcpyAsync(hst + ST_TRUE, status + ST_TRUE, sizeof(double), cudaMemcpyDeviceToHost, stream));
        TNPY_CUDA_OK(cudaStreamSynchronize(stream));
        ++n_looks;
        if (hst[ST_TRUE] <= tol * hst[ST_ANORM]) {
          hst[ST_RESID] = hst[ST_TRUE];
          if (hpsi) TNPY_TRY(scale_copy(Y, hpsi, n, 1.0, nullptr, 0, stream));
          break;
        }
        // not there yet: carry on from psi with every product at full accuracy
        ++n_failed_checks;
        inexact = false;
        used_inexact = false;
        cur_slices = slices;
        done = false;
        if (n_matvec >= max_matvec) {
          if (hpsi) TNPY_TRY(scale_copy(Y, hpsi, n, 1.0, nullptr, 0, stream));
          hst[ST_RESID] = hst[ST_TRUE];
          break;
        }
        TNPY_TRY(scale_copy(psi, V, n, 1.0, nullptr, 0, stream));
        TNPY_CUDA_OK(cudaMemsetAsync(T, 0, sizeof(double) * kMaxNcv * kMaxNcv, stream));
        j = 0;
        whole_basis_step = 0;
        since_check = 0;
        stride = 1;
        last_resid = 0.0;
        ++n_restart;
        continue;
      }
      if (hpsi) {
        // H psi from the Lanczos relation H V_m = V_m T + beta v_{m+1} e_m^T (exact to rounding here, T being the
        // explicit projection): H psi = theta psi + (beta s_m) v_{m+1}; v_{m+1} = V[m] was normalised above
        TNPY_TRY(scale_copy(psi, hpsi, n, hst[ST_THETA], nullptr, 0, stream));
        if (hst[ST_BETA] > 0.0 && m < n_full) TNPY_TRY(axpy(hst[ST_RCOEF], nullptr, V + (int64_t)m * ldv, hpsi, n, stream));
      }
      break;
    }
    if (m == ncv) {
      // thick restart: keep the `keep` lowest Ritz vectors plus the residual direction V[m]
      const int k = keep < m ? keep : m;
      TNPY_TRY(combine(V, ldv, m, S, kMaxNcv, k, Y, ldv, n, stream));
      TNPY_CUDA_OK(cudaMemcpyAsync(V + (int64_t)k * ldv, V + (int64_t)m * ldv, sizeof(double) * n,
                                   cudaMemcpyDeviceToDevice, stream));
      TNPY_CUDA_OK(cudaMemcpyAsync(V, Y, sizeof(double) * (size_t)k * ldv, cudaMemcpyDeviceToDevice, stream));
      restart_T_kernel<<<1, 256, 0, stream>>>(T, thetas, k);
      TNPY_LAUNCH_OK();
      j = k;
      whole_basis_step = k;  // H v_k has a component on every kept Ritz vector
      ++n_restart;
    } else {
      ++j;
    }
  }
  TNPY_CUDA_OK(cudaStreamSynchronize(stream));
  if (stats_host) {
    stats_host[0] = hst[ST_THETA];
    stats_host[1] = hst[ST_RESID];
    stats_host[2] = (double)n_matvec;
    stats_host[3] = (double)n_restart;
    stats_host[4] = done ? 1.0 : 0.0;
    stats_host[5] = hst[ST_ANORM];
    stats_host[6] = worst_bound > hst[ST_BOUND] ? worst_bound : hst[ST_BOUND];  // rigorous bound on the int8 products' error
    stats_host[7] = (double)(plan.mode * 10 + (plan.mode == HEFF_FP64_CHAIN ? 0 : slices));
  }
  g_last_counters[0] = n_matvec;
  g_last_counters[1] = n_looks;
  g_last_counters[2] = (long long)hst[ST_EXTRA];
  g_last_counters[3] = n_restart;
  g_last_counters[4] = n_reduced;
  g_last_counters[6] = n_five;
  g_last_counters[5] = n_failed_checks;
  if (!done) {
    set_error("tnpy_eig_lowest: not converged after %d matvecs (resid %.3e, tol*|A| %.3e)", n_matvec, hst[ST_RESID],
              tol * hst[ST_ANORM]);
    return TNPY_ENOCONV;
  }
  return TNPY_OK;
}

extern "C" int tnpy_eig_lowest(const double* L, const double* W, const double* R, double* psi, int l, int r, int wl,
                               int wr, int d, int flags, double tol, int max_matvec, int ncv_in, double* stats_host,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  return eig_lowest_impl(nullptr, L, W, R, psi, nullptr, l, 0, l, r, wl, wr, d, flags, tol, max_matvec, ncv_in, stats_host,
                         workspace, workspace_bytes, stream_);
}

extern "C" int tnpy_eig_lowest_image(const double* L, const double* W, const double* R, double* psi, double* hpsi,
                                     int l, int r, int wl, int wr, int d, int flags, double tol, int max_matvec,
                                     int ncv_in, double* stats_host, void* workspace, size_t workspace_bytes,
                                     void* stream_) {
  TNPY_CHECK_ARG(hpsi != nullptr, "null hpsi");
  return eig_lowest_impl(nullptr, L, W, R, psi, hpsi, l, 0, l, r, wl, wr, d, flags, tol, max_matvec, ncv_in, stats_host,
                         workspace, workspace_bytes, stream_);
}

extern "C" size_t tnpy_eig_rows_workspace_bytes(int l, int l_rows, int r, int wl, int wr, int d, int ncv_in) {
  int ncv, keep;
  pick_sizes((int64_t)l * d * r, ncv_in, ncv, keep);
  return eig_ws_layout((int64_t)l_rows * d * r, ncv, keep,
                       heff_plan_bytes(l, l_rows, r, wl, wr, d) + heff_apply_bytes(l, l_rows, r, wl, wr, d) + 1024) +
         Workspace::need((size_t)l * d * r);
}

extern "C" int tnpy_eig_lowest_rows(const tnpy_comm* comm, const double* L_rows, const double* W, const double* R,
                                    double* psi_rows, double* hpsi_rows, int l, int row0, int l_rows, int r, int wl, int wr,
                                    int d, int flags, double tol, int max_matvec, int ncv_in, double* stats_host,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  TNPY_CHECK_ARG(comm != nullptr, "null communicator");
  return eig_lowest_impl(comm, L_rows, W, R, psi_rows, hpsi_rows, l, row0, l_rows, r, wl, wr, d, flags, tol, max_matvec,
                         ncv_in, stats_host, workspace, workspace_bytes, stream_);
}

extern "C" int tnpy_last_eig_counters(int64_t* out, int n) {
  int k = 0;
  for (; k < n && k < 7; ++k) out[k] = g_last_counters[k];
  return k;
}

extern "C" int tnpy_set_inexact_slices(int on) { return inexact_switch().exchange(on ? 1 : 0); }
