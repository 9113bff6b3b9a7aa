// Small sites: several whole Lanczos steps (matvec chain + two Gram-Schmidt passes + normalisation + the new column
// of T) in ONE cooperative launch (csrc/lanczos_steps.cu).  Used by the eigensolver (csrc/lanczos.cu) when the vector
// is short enough that a step is bound by kernel latencies, not by arithmetic or bandwidth.
#pragma once
#include "common.cuh"

namespace tnpy {

constexpr int kStepsMaxNcv = 48;  // == kMaxNcv of lanczos.cu: leading dimension of T, longest basis

struct LanczosStepsPlan {
  int grid;     // work CTAs of the cooperative launch (the launch adds one that watches convergence; < SM count)
  int chunk;    // vector elements owned by one CTA (<= 256)
  int ksplit;   // the second GEMM's K range is cut into `ksplit` pieces of `kchunk` rows
  int kchunk;
  size_t bytes; // scratch the launch needs from the caller's workspace
};

// true when this site runs on the fused path: full (unsharded) problem, FP64, short vectors, small MPO tensor
bool lanczos_steps_supported(int l, int r, int wl, int wr, int d);
LanczosStepsPlan lanczos_steps_plan(int l, int r, int wl, int wr, int d);

// Runs Lanczos steps j0 .. j0 + nsteps - 1 on the basis V (column j at V + j * ldv; V[j0] normalised): for each step
// w = H_eff v_j, two classical Gram-Schmidt passes against V[0..j], T[:, j] = T[j, :] = the summed coefficients,
// V[j + 1] = w / ||w|| (j + 1 <= ncv, the capacity of the basis).  status[beta_slot] = the last ||w||, status[steps_slot] = steps done (fewer than nsteps only
// after an exact breakdown ||w|| == 0, or when the launch stopped itself).  `scratch` holds plan.bytes.
// tol > 0: one extra CTA follows the lowest Ritz pair (theta, z) of T (assumed to be diag(`arrow` kept Ritz values) +
// their coupling to row `arrow` + a tridiagonal tail, i.e. what thick-restart Lanczos produces up to rounding-level
// fill) and the launch returns two steps after |beta z_last| <= 0.7 tol max(anorm, |theta|) first held.
int lanczos_steps_launch(const LanczosStepsPlan& plan, const double* L, const double* W, const double* R, double* V,
                         int64_t ldv, double* T, double* status, int beta_slot, int steps_slot, int l, int r, int wl,
                         int wr, int d, int j0, int nsteps, int ncv, int arrow, double tol, double anorm, void* scratch,
                         cudaStream_t stream);

// Mid-size sites (up to 2^20 unknowns): the Gram-Schmidt half of step j in one cooperative launch.  On entry V[j + 1]
// holds H v_j as the matvec left it; on return V[j + 1] is orthogonalised twice against V[0..j] and normalised,
// T[:, j] = T[j, :] holds the summed coefficients and status[beta_slot] the norm.  scratch: lanczos_gs_bytes().
bool lanczos_gs_supported(int64_t n);
size_t lanczos_gs_bytes();
int lanczos_gs_launch(double* V, int64_t ldv, int64_t n, int j, double* T, double* status, int beta_slot, void* scratch,
                      cudaStream_t stream);

}  // namespace tnpy
