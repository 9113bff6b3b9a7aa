/*
 * tnpy_cuda.h -- C ABI of the B200-native finite-DMRG local-update hot path.
 *
 * This is the boundary a maintainer of tanlin2013/tnpy would bind (ctypes stub in INTEGRATION.md).
 * The reference has no FFI of its own -- the path is pure Python on top of quimb / primme /
 * numpy -- so every entry point names the reference *Python* callable it replaces (file:line under
 * /root/reference/tnpy).
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers to C-order (row-major) float64 unless the name ends
 *     in `_host`.  The caller (PyTorch on the Python side) allocates inputs, outputs and workspaces; the library
 *     only borrows them for the duration of the call (a tnpy_heff_plan borrows its plan memory for the life of the
 *     handle).  The one piece of device memory the library owns is a 0.6 MB reduction scratch per (device, stream),
 *     allocated on first use by the vector entry points, which take no workspace argument.
 *   - One process per GPU, calls from one host thread at a time per stream.  Reduction scratch, the pinned status
 *     word and kernel attributes are kept per device / per stream, so several streams of one device may be used
 *     concurrently from different host threads.
 *   - Every call takes the CUDA stream to enqueue on (as a `void*` == cudaStream_t) and is
 *     asynchronous with respect to the host unless documented otherwise.
 *   - Return value: 0 on success, a negative TNPY_E* code otherwise.  Never throws.
 *     `tnpy_last_error()` returns a thread-local human-readable message for the last failure.
 *   - Layouts (reference: matrix_product_state.py:40-45, :411-440):
 *       site tensor x / A : (l, d, r)          "lpr"
 *       MPO tensor W      : (wl, wr, d, d)     "lrud"  (u = ket index, d = bra index)
 *       left env  L       : (l, wl, l)         (ket bond, MPO bond, bra bond)
 *       right env R       : (r, wr, r)         (ket bond, MPO bond, bra bond)
 *     Edge sites are expressed with unit bonds: pass l == 1, wl == 1 and L == NULL for site 0
 *     (r == 1, wr == 1, R == NULL for the last site); NULL stands for the 1x1x1 tensor [1.0].
 */
#ifndef TNPY_CUDA_H_
#define TNPY_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNPY_OK 0
#define TNPY_EINVAL (-1)    /* bad argument (null pointer, non-positive dimension, ...) */
#define TNPY_EWORKSPACE (-2) /* caller-supplied workspace too small */
#define TNPY_ECUDA (-3)     /* a CUDA runtime / driver call failed */
#define TNPY_ENOCONV (-4)   /* iterative routine hit its iteration limit (result still written) */

/* GEMM kernel selection (tnpy_set_gemm_algo, the `algo` arguments, env TNPY_GEMM_ALGO=auto|generic|dmma|ozaki|fp64). */
#define TNPY_GEMM_AUTO 0    /* the default: large products of the contraction chains run FP64-accurately on the
                               tcgen05 int8 tensor cores (Ozaki scheme, csrc/ozaki.cu), everything else on the
                               FP64 tensor pipe (TMA + DMMA) or, for tiny / unaligned operands, the generic kernel */
#define TNPY_GEMM_GENERIC 1 /* generic shared-memory tiled DFMA kernel (any shape / stride)       */
#define TNPY_GEMM_DMMA 2    /* force the TMA + mbarrier + FP64 tensor-core (DMMA) kernel           */
#define TNPY_GEMM_OZAKI 3   /* same selection as AUTO, spelled out                                  */
#define TNPY_GEMM_FP64 4    /* native FP64 arithmetic only: DMMA when TMA-describable, else generic; never the
                               int8 tensor-core path                                                   */

/* Canonical-gauge shortcuts for the contraction chains (`flags` arguments).  With tnpy's
 * upper-triangular MPOs (model/utils.py:25-28: row 0 / last column are the boundary vectors) and a
 * mixed-canonical MPS, L[:, 0, :] and R[:, w_r - 1, :] are identity matrices.  The caller vouches for a
 * flag (tnpy_identity_defect measures it); a set flag replaces that channel's GEMM slice by a
 * transpose/copy.  Results change at the level of the measured defect; executed flops drop by up to
 * 2/w while the algorithmic flop count F_mv quoted everywhere is unchanged. */
#define TNPY_LEFT_IDENTITY 1
#define TNPY_RIGHT_IDENTITY 2

/* ---- library ---------------------------------------------------------------------------- */
int tnpy_version(void);
const char* tnpy_last_error(void);
/* Number of kernels this library has launched from the calling process so far (for bench.py's
 * `gpu_launches`). */
int64_t tnpy_launch_count(void);
/* Override the process-wide GEMM selection of the contraction chains (default TNPY_GEMM_AUTO). */
int tnpy_set_gemm_algo(int algo);

/* Force the DMMA tile configuration (benchmarking): -1 auto, 0 = 128x128, 1 = 128x64, 2 = 64x64. */
int tnpy_set_gemm_tile(int cfg);
/* Register-resident FP64 throughput probe (fixes the FP64 roofline denominator on the box):
 * kind 0 = DMMA m8n8k4 tensor pipe, kind 1 = DFMA.  Synchronous; writes achieved TFLOP/s. */
int tnpy_probe_fp64(int kind, int threads_per_block, int blocks_per_sm, int ilp, int iters,
                    double* tflops_host, void* scratch_dev);

/* ---- dense building block ----------------------------------------------------------------
 * C[m, n] (+)= sum_k A[k, m] * B[k, n]   (C = A^T B, all three row-major, contraction index slow)
 * A: K x M with row stride lda, B: K x N with row stride ldb, C: M x N with row stride ldc.
 * This is the shape every big contraction of the path takes once the contracted bond is the
 * slowest index of both operands; replaces the numpy.tensordot -> dgemm lowering that
 * quimb/opt_einsum produce for matrix_product_state.py:315, :336, :438.
 * accumulate != 0 adds into C. */
int tnpy_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                 int M, int N, int K, int accumulate, int algo, void* stream);

/* ---- the same GEMM in FP64 accuracy on the tcgen05 int8 tensor cores (Ozaki scheme: eight error-free 7-bit
 * slices per operand column, the first `slices` of them multiplied as exact int8 x int8 -> int32 slice GEMMs with
 * the accumulators in TMEM, FP64 recombination in the epilogue).  This is what the contraction chains run for their
 * large products; the entry point exposes it for tests and benchmarks.
 * phase: 0 = slice + multiply, 1 = slice only into the workspace, 2 = multiply from the workspace.
 * Error (rigorous): |C - A^T B|[m][n] <= K e_S sa[m] sb[n], e_S = (S+2)/4 2^(-7S), sa / sb = the power-of-two column
 * scales (> 2 max|column|); tnpy_ozaki_error_bound writes K e_S ||sa||_2 ||sb||_2 >= ||C - A^T B||_F for the
 * operands last sliced into `workspace` to *bound_dev. */
size_t tnpy_ozaki_workspace_bytes(int M, int N, int K, int slices);
/* Slices used by chain products nobody declared a tolerance for: 8 (default, componentwise error ~4e-16, same as
 * DMMA), 7 (~2e-14) or 6 (~3e-12).  tnpy_eig_lowest picks 7 by itself when its tolerance is >= 1e-10 and the
 * rigorous bound of every product stays below 1 % of tol * ||A||. */
int tnpy_set_ozaki_slices(int slices);
int tnpy_ozaki_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                       int M, int N, int K, int slices, int accumulate, int phase, void* workspace,
                       size_t workspace_bytes, void* stream);
int tnpy_ozaki_error_bound(int M, int N, int K, int slices, void* workspace, size_t workspace_bytes,
                           double* bound_dev, void* stream);

/* ---- a1: Environment.one_site_matvec(site).matvec(x)  (matrix_product_state.py:411-440) ----
 * y[m,q,s] = sum_{l,a,p,b,r} L[l,a,m] W[a,b,p,q] R[r,b,s] x[l,p,r]
 * x, y: (l, d, r).  Workspace: tnpy_heff_workspace_bytes(). */
size_t tnpy_heff_workspace_bytes(int l, int r, int wl, int wr, int d);
int tnpy_heff_apply(const double* L, const double* W, const double* R, const double* x, double* y,
                    int l, int r, int wl, int wr, int d, int flags, void* workspace,
                    size_t workspace_bytes, void* stream);
/* Prepared H_eff (what `Environment.one_site_matvec(site)` returns in the reference is an operator object that is
 * applied 10-30 times per site, matrix_product_state.py:411-440): everything that depends only on (L, W, R) is
 * computed once -- on the tcgen05 path the int8 slices of the environments -- into caller-owned `plan_memory`
 * (tnpy_heff_plan_bytes), which must stay untouched while the handle lives, as must L, W and R.  The handle itself
 * is a small host object; destroy it with tnpy_heff_plan_destroy.  W_host: host copy of W or NULL (the library then
 * reads W back, synchronising the stream once, when that decides the mode).  algo: TNPY_GEMM_*.
 * Modes (tnpy_heff_plan_mode): 0 = FP64 chain, 1 = chain with its large GEMMs on tcgen05, 2 = the direct path of the
 * mixed-canonical gauge: with both identity flags and an MPO tensor whose non-zero blocks all have a = 0 or
 * b = wr - 1,  y = sum_{b<wr-1} (W_0b x) R_b + sum_{a>0} L_a^T (W_{a,wr-1} x) + W_{0,wr-1} x  is two independent
 * tcgen05 GEMMs whose x-side operands are mixed and sliced straight from x -- no FP64 intermediate exists.
 * tnpy_heff_plan_apply: y = H_eff x; slices = 0 (default) or 5 .. 8; workspace: tnpy_heff_workspace_bytes().
 * tnpy_heff_plan_error_bound: the largest rigorous Frobenius-norm bound of any int8 product issued through the
 * plan so far (0 on the FP64 chain), copied device to device. */
typedef struct tnpy_heff_plan tnpy_heff_plan;
size_t tnpy_heff_plan_bytes(int l, int r, int wl, int wr, int d);
int tnpy_heff_plan_create(tnpy_heff_plan** handle, const double* L, const double* W, const double* R,
                          const double* W_host, int l, int r, int wl, int wr, int d, int flags, int algo,
                          void* plan_memory, size_t plan_bytes, void* stream);
int tnpy_heff_plan_mode(const tnpy_heff_plan* handle);
int tnpy_heff_plan_apply(const tnpy_heff_plan* handle, const double* x, double* y, int slices, void* workspace,
                         size_t workspace_bytes, void* stream);
int tnpy_heff_plan_error_bound(const tnpy_heff_plan* handle, double* bound_dev_out, void* stream);
int tnpy_heff_plan_destroy(tnpy_heff_plan* handle);

/* max_ij |E[i, channel, j] - delta_ij| of an environment E (dim, w, dim), written to device memory.
 * Workspace: 8 KB. */
int tnpy_identity_defect(const double* E, int dim, int w, int channel, double* defect_dev,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Row block of the same matvec for the chi-sharded multi-GPU layout: the caller holds
 * L_rows = L[:, :, row0 : row0 + l_rows] stored contiguously as (l, wl, l_rows), the full x and R, and gets
 * y_rows = y[row0 : row0 + l_rows] as (l_rows, d, r).  No reduction across ranks is needed; the ranks
 * all-gather the next x.  flags as for tnpy_heff_apply (TNPY_LEFT_IDENTITY then means
 * L[li, 0, row0 + m] = delta(li, row0 + m)); in the mixed-canonical gauge the R-side term of the direct path only
 * reads the caller's own rows of x.  Workspace: tnpy_heff_workspace_bytes() of the full problem is enough. */
int tnpy_heff_apply_rows(const double* L_rows, const double* W, const double* R, const double* x,
                         double* y_rows, int l, int row0, int l_rows, int r, int wl, int wr, int d, int flags,
                         void* workspace, size_t workspace_bytes, void* stream);
/* Prepared form of the row block: same handle type and tnpy_heff_plan_apply (x full (l, d, r) in, y_rows (l_rows, d, r)
 * out); plan memory tnpy_heff_plan_bytes() of the full problem is enough. */
int tnpy_heff_plan_create_rows(tnpy_heff_plan** handle, const double* L_rows, const double* W, const double* R,
                               const double* W_host, int l, int row0, int l_rows, int r, int wl, int wr, int d,
                               int flags, int algo, void* plan_memory, size_t plan_bytes, void* stream);

/* ---- a7: Environment.update_left / update_right  (matrix_product_state.py:296-336) ---------
 * left : Lout[r,b,s] = sum L[l,a,m] A[l,p,r] W[a,b,p,q] A[m,q,s]      Lout: (r, wr, r)
 * right: Rout[l,a,m] = sum R[r,b,s] A[l,p,r] W[a,b,p,q] A[m,q,s]      Rout: (l, wl, l)
 * A: (l, d, r).  Workspace: tnpy_env_workspace_bytes(). */
size_t tnpy_env_workspace_bytes(int l, int r, int wl, int wr, int d);
int tnpy_env_update_left(const double* L, const double* A, const double* W, double* Lout,
                         int l, int r, int wl, int wr, int d, int flags, void* workspace,
                         size_t workspace_bytes, void* stream);
int tnpy_env_update_right(const double* R, const double* A, const double* W, double* Rout,
                          int l, int r, int wl, int wr, int d, int flags, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Row block of the left update for the chi-sharded sweep: L_rows = L[:, :, row0 : row0 + l_rows] as (l, wl, l_rows), A in
 * full; Lout_partial (r, wr, r) receives this block's contribution (the sum over the bra index runs over the block
 * only) -- the ranks' contributions add up to the update (one all-reduce).  Workspace: tnpy_env_workspace_bytes(). */
int tnpy_env_update_left_rows(const double* L_rows, const double* A, const double* W, double* Lout_partial,
                              int l, int row0, int l_rows, int r, int wl, int wr, int d, int flags, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---- a6: Environment.one_site_full_matrix  (matrix_product_state.py:372-409) ---------------
 * H[(l,p,r),(m,q,s)] = sum_{a,b} L[l,a,m] W[a,b,p,q] R[r,b,s] dense, N = l*d*r <= 32768; H is N x N
 * row-major.  With the workspace of the query (d^2 l^2 wr doubles) the matrix is built in two passes -- L W first,
 * then wr multiply-adds per entry --, without one (NULL, 0) or for tiny sites in one pass of wl * wr per entry. */
size_t tnpy_heff_dense_workspace_bytes(int l, int r, int wl, int wr, int d);
int tnpy_heff_dense(const double* L, const double* W, const double* R, double* H,
                    int l, int r, int wl, int wr, int d, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ---- vector kernels used by the on-device eigensolver (replace primme's host BLAS-1/2) -----
 * Results of reductions are written to DEVICE memory (result pointers are device pointers). */
int tnpy_dot(const double* x, const double* y, int64_t n, double* result, void* stream);
int tnpy_nrm2(const double* x, int64_t n, double* result, void* stream);
/* y += alpha * x   (alpha passed by value) */
int tnpy_axpy(double alpha, const double* x, double* y, int64_t n, void* stream);
/* y += (*alpha_dev * alpha_scale) * x   (alpha read from device memory: no host round trip) */
int tnpy_axpy_dev(const double* alpha_dev, double alpha_scale, const double* x, double* y, int64_t n,
                  void* stream);
int tnpy_scal(double alpha, double* x, int64_t n, void* stream);
/* h[j] = sum_i V[j, i] * w[i], j < m; V is m x n row-major with row stride ldv. */
int tnpy_multi_dot(const double* V, int64_t ldv, int m, const double* w, int64_t n, double* h,
                   void* stream);
/* w[i] -= sum_j h[j] * V[j, i] */
int tnpy_multi_axpy(const double* V, int64_t ldv, int m, const double* h, double* w, int64_t n,
                    void* stream);

/* ---- a2: linalg.eigshmv -> primme.eigsh(A, v0, k=1, which="SA", tol)  (linalg.py:64-87) ------
 * Lowest eigenpair of H_eff(L, W, R) by thick-restart Lanczos with full reorthogonalisation, run
 * entirely on the device (matvec chain + vector kernels + on-device Ritz solve + convergence flag).
 * psi: in = start vector v0, out = normalised eigenvector, N = l*d*r doubles.
 * Stops when ||H x - theta x|| <= tol * max|Ritz| (primme's rule); tol <= 0 means 1e4 * eps.
 * stats_host (>= 8 doubles, host memory): [0] = theta, [1] = residual norm, [2] = matvec count,
 * [3] = restart count, [4] = converged (1/0), [5] = max |Ritz value| (the ||A|| estimate), [6] = largest rigorous
 * error bound of an int8 product during the solve (0 on the FP64 chain), [7] = 10 * plan mode + slices in use at the
 * end.  The call synchronises the stream before returning (one host read-back per convergence check, never per
 * matvec).
 * Workspace: tnpy_eig_workspace_bytes(). */
size_t tnpy_eig_workspace_bytes(int l, int r, int wl, int wr, int d, int ncv);
int tnpy_eig_lowest(const double* L, const double* W, const double* R, double* psi,
                    int l, int r, int wl, int wr, int d, int flags, double tol, int max_matvec, int ncv,
                    double* stats_host, void* workspace, size_t workspace_bytes, void* stream);
/* Same solve, and hpsi (N doubles) = H_eff psi for the returned psi at no extra matvec, from the Lanczos
 * relation H V_m = V_m T + beta v_{m+1} e_m^T:  H psi = theta psi + (beta s_m) v_{m+1}  (equal to a fresh
 * tnpy_heff_apply(psi) to rounding).  FiniteDMRG.perturb_wave_function (finite_dmrg.py:116-141), which the
 * sweep calls right after the solve, needs exactly that vector. */
int tnpy_eig_lowest_image(const double* L, const double* W, const double* R, double* psi, double* hpsi,
                          int l, int r, int wl, int wr, int d, int flags, double tol, int max_matvec, int ncv,
                          double* stats_host, void* workspace, size_t workspace_bytes, void* stream);
/* Diagnostics of the calling thread's last tnpy_eig_lowest* call: matvecs, looks (status read-backs = stream
 * synchronisations), extra full Gram-Schmidt passes the DGKS test asked for, thick restarts, matvecs that ran with
 * fewer int8 slices than the solve's base count (inexact-Krylov schedule), true-residual checks that failed and sent the
 * solve on at full accuracy, matvecs with five slices.  Returns how many were written (at most 7). */
int tnpy_last_eig_counters(int64_t* out, int n);
/* On the tcgen05 path tnpy_eig_lowest runs the later steps of a solve with fewer int8 slices (inexact Krylov: the
 * matvec error a Lanczos step tolerates grows like 1 / ||r|| of the current Ritz pair), chosen at every look from the
 * rigorous bound of the products, bound(S) <= (0.01 / 8) tol ||A||^2 / ||r||, S >= 5; a solve that did so is accepted only
 * on its *true* residual ||H psi - theta psi|| <= tol ||A||, formed with one matvec at full accuracy (which is also the
 * image tnpy_eig_lowest_image returns), and otherwise continues from psi with the schedule off.  on = 0 switches the
 * schedule off for the process (also: environment TNPY_INEXACT_SLICES=0).  Returns the previous setting. */
int tnpy_set_inexact_slices(int on);
/* Small sites (vectors of <= 32768 elements on the FP64 chain) run whole Lanczos steps -- matvec, two Gram-Schmidt
 * passes, normalisation, the new column of T -- in one cooperative launch, `stride` steps per launch, instead of
 * ~15 launches per step (csrc/lanczos_steps.cu).  on = 0 keeps every site on the general multi-kernel solver (also:
 * environment TNPY_FUSED_STEPS=0).  Returns the previous setting. */
int tnpy_set_fused_steps(int on);
/* Diagnostics: with a device buffer of 64 * 16 uint64 set, every fused launch records %globaltimer (ns) at the phase
 * boundaries of its first 64 steps -- slots 0-7 of a step by CTA 0 (step start, P1 done, after barrier 1, P2 done,
 * P3 done, P4 done, P5 done, after barrier 5), slots 8-15 the same by the convergence-watching CTA.  NULL switches it
 * off (the default). */
int tnpy_steps_trace(void* device_buffer);

/* ---- chi-sharded local solve over the GPUs of one box (BASELINE configs[4]) -------------------
 * One process per GPU.  A communicator wraps an NCCL communicator that the library creates itself (libnccl.so.2 is
 * resolved at run time: the copy PyTorch has loaded, else the system's; nothing links against it): rank 0 calls
 * tnpy_comm_unique_id, the TNPY_COMM_ID_BYTES bytes travel to every rank by any means (tnpy_b200/parallel.py uses a
 * torch.distributed broadcast), every rank calls tnpy_comm_create -- a collective -- on its own device.
 * tnpy_comm_allgather / tnpy_comm_allreduce_sum: the two collectives the solve uses, on device doubles. */
#define TNPY_COMM_ID_BYTES 128
typedef struct tnpy_comm tnpy_comm;
int tnpy_comm_unique_id(char* id_out);
int tnpy_comm_create(tnpy_comm** comm, const char* id, int world, int rank);
int tnpy_comm_destroy(tnpy_comm* comm);
int tnpy_comm_world(const tnpy_comm* comm);
int tnpy_comm_rank(const tnpy_comm* comm);
int tnpy_comm_allgather(const tnpy_comm* comm, const double* send, double* recv, int64_t count, void* stream);
int tnpy_comm_allreduce_sum(const tnpy_comm* comm, double* buf, int64_t count, void* stream);
/* tnpy_eig_lowest with the left bond's bra rows split evenly over the ranks: rank g holds rows [row0, row0 + l_rows),
 * row0 = g * l_rows, l_rows * world == l -- L_rows = L[:, :, rows] contiguous as (l, wl, l_rows), the same rows of psi
 * (in: start vector, out: eigenvector), of hpsi (may be NULL; H_eff psi as in tnpy_eig_lowest_image) and of every
 * Lanczos vector -- plus full copies of W and R.  Per Lanczos step the ranks all-gather the current vector (the
 * only exchange of data, (G-1)/G of 8 N bytes per rank over NVLink), run their row block of the matvec
 * (tnpy_heff_apply_rows; in the mixed-canonical gauge its R-side term needs no remote data at all) and all-reduce
 * the Gram-Schmidt coefficients and norms; the small Ritz problem is solved redundantly from identical inputs, so
 * every rank takes the same decisions and returns the same stats.  flags as for tnpy_heff_apply_rows.
 * Collective: every rank of the communicator must call it with the same l, r, wl, wr, d, flags, tol, max_matvec, ncv. */
size_t tnpy_eig_rows_workspace_bytes(int l, int l_rows, int r, int wl, int wr, int d, int ncv);
int tnpy_eig_lowest_rows(const tnpy_comm* comm, const double* L_rows, const double* W, const double* R,
                         double* psi_rows, double* hpsi_rows, int l, int row0, int l_rows, int r, int wl, int wr,
                         int d, int flags, double tol, int max_matvec, int ncv, double* stats_host, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- f2: ShiftInvertDMRG.one_site_solver -> primme.eigsh(A, M=M, k=1, which="SA")  (finite_dmrg.py:341-355)
 * Lowest eigenpair of the symmetric-definite pencil  A x = lambda M x,  A = H_eff(LA, WA, RA) (MPO of
 * H - eps), M = H_eff(LM, WM, RM) (its square, w^2 channels), both at the same site (same l, r, d).
 * Generalised Davidson on the device; psi: in = start vector, out = eigenvector normalised to
 * x^T M x = 1 (the convention of primme.eigsh(A, M=M) and scipy.linalg.eigh(a, b)).  Stops when ||A x - theta M x|| <= tol (||A x|| + |theta| ||M x||).
 * stats_host as for tnpy_eig_lowest ([2] counts iterations = one A and one M matvec each). */
size_t tnpy_geig_workspace_bytes(int l, int r, int wl_a, int wr_a, int wl_m, int wr_m, int d, int ncv);
int tnpy_geig_lowest(const double* LA, const double* WA, const double* RA, const double* LM,
                     const double* WM, const double* RM, double* psi, int l, int r, int wl_a, int wr_a,
                     int wl_m, int wr_m, int d, int flags_a, double tol, int max_iter, int ncv,
                     double* stats_host, void* workspace, size_t workspace_bytes, void* stream);

/* Dense route for the same pencil (reference: scipy.linalg.eigh(a, b), finite_dmrg.py:344-348): a, b are
 * n x n row-major symmetric (b positive definite), both destroyed.  x is normalised to x^T b x = 1. */
size_t tnpy_geig_dense_workspace_bytes(int n);
int tnpy_geig_dense_lowest(double* a, double* b, int n, double* theta_dev, double* x, void* workspace,
                           size_t workspace_bytes, void* stream);
/* The same pencil for large n (10^3 .. 3 x 10^4 unknowns): b = D C C^T D by a blocked Cholesky factorisation (D =
 * sqrt(diag b)), S = X^T a X with X = D^-1 C^-T on the FP64 tensor pipe, lowest eigenpair of S by the on-device
 * Lanczos solver of tnpy_eig_lowest (tol, max_matvec and stats_host as there; stats_host may be NULL), x = X z with
 * x^T b x = 1.  a and b are left intact.  TNPY_ENOCONV with "not positive definite" in tnpy_last_error() when the
 * factorisation breaks down, or when the Lanczos iteration did not reach tol (x, theta then hold the best pair).
 * Workspace: tnpy_geig_chol_workspace_bytes(n) (about 4 padded n x n matrices). */
size_t tnpy_geig_chol_workspace_bytes(int n);
int tnpy_geig_chol_lowest(const double* a, const double* b, int n, double tol, int max_matvec, double* theta_dev,
                          double* x, double* stats_host, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a5: linalg.eigh(matrix)  (linalg.py:42-61), k = 1 --------------------------------------
 * Lowest eigenpair of a dense symmetric N x N matrix (row-major, destroyed) by cyclic Jacobi
 * on the device.  evec: N doubles, eval_dev: 1 double (device). */
size_t tnpy_eigh_workspace_bytes(int n);
int tnpy_eigh_lowest(double* H, int n, double* eval_dev, double* evec, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- a8: linalg.svd(matrix, cutoff)  (linalg.py:9-23) ---------------------------------------
 * Thin SVD of a row-major rows x cols matrix A (destroyed): A = U diag(s) Vt with singular values
 * sorted descending (ties broken by original column index => deterministic).  k = min(rows, cols).
 * U: rows x k (ldu = k), s: k, Vt: k x cols (ldvt = cols).  One-sided Jacobi (Hestenes) on the short
 * side: a single-CTA shared-memory kernel for small problems, block Jacobi on the FP64 tensor pipe
 * otherwise.  Exactly-zero singular values are reported as 0 with an orthonormal
 * completion of U / Vt, as LAPACK does.  Synchronises the stream once per Jacobi sweep.
 * Workspace: tnpy_svd_workspace_bytes(). */
size_t tnpy_svd_workspace_bytes(int rows, int cols);
/* Jacobi sweeps the last tnpy_svd call on this process needed (-1: single-CTA shared-memory path). */
int tnpy_last_svd_sweeps(void);
/* Significant (not yet orthogonal) pair counts per Jacobi sweep of that call; returns how many were written. */
int tnpy_last_svd_trace(unsigned int* counts, int max_counts);
int tnpy_svd(double* A, int rows, int cols, double* U, double* s, double* Vt, void* workspace,
             size_t workspace_bytes, void* stream);

/* ---- a9: MatrixProductState.split_tensor neighbour absorb  (matrix_product_state.py:207-223) --
 * right: out[k, j] = sum_i s[k] * Vt[k, i] * Nb[i, j]      Nb: (k, cols_nb)   (diag(s) Vt . A[site+1])
 * left : out[i, k] = sum_j Nb[i, j] * U[j, k] * s[k]       Nb: (rows_nb, k)   (A[site-1] . U diag(s)) */
int tnpy_absorb_right(const double* s, const double* Vt, int k, int n, const double* Nb, int cols_nb,
                      double* out, void* workspace, size_t workspace_bytes, void* stream);
int tnpy_absorb_left(const double* U, const double* s, int n, int k, const double* Nb, int rows_nb,
                     double* out, void* workspace, size_t workspace_bytes, void* stream);
/* nb = cols_nb (right) or rows_nb (left) */
size_t tnpy_absorb_workspace_bytes(int k, int n, int nb);

/* ---- a8/a9 without the SVD: orthogonal split ("QR-then-small-SVD", the small SVD deferred) -----
 * linalg.svd is only ever called with cutoff = current bond (matrix_product_state.py:206, :218), i.e. as an
 * orthogonalisation: any A = Q T with orthonormal Q leaves the state, the environments and every later local
 * problem unchanged (a bond gauge), and the singular values of the bond are those of the small square T.
 *   rows > cols:  A = Q T,  Q: rows x cols with orthonormal columns, T: cols x cols   (the U, diag(s) Vt slots)
 *   rows < cols:  A = T Q,  T: rows x rows, Q: rows x cols with orthonormal rows      (the U diag(s), Vt slots)
 *   rows == cols: A = Q T, or A = T Q when flags has TNPY_QR_T_FIRST (a leftward split of a square site tensor)
 * Cholesky-QR applied twice to the norm-scaled vectors, all big products on the FP64 tensor pipe (csrc/qr.cu).
 * A is not modified.  *defect_dev (device double) receives max|Q^T Q - I| as measured on the device, +inf if a
 * Cholesky pivot broke down or a vector is exactly zero: the caller must check it (<= ~1e-13) and fall back to
 * tnpy_svd otherwise.  flags = TNPY_QR_SHIFTED runs the shifted three-pass variant (first factorisation on
 * G + sigma I, sigma = 100 u n): for vectors too ill-conditioned for two passes (cond of the normalised vectors
 * between ~1e7 and ~1e12, cold sweeps), at 1.5x the cost.  Never synchronises the stream.
 * Workspace: tnpy_qr_split_workspace_bytes(). */
#define TNPY_QR_SHIFTED 1
#define TNPY_QR_T_FIRST 2 /* square input only: factorise as T Q instead of Q T */
size_t tnpy_qr_split_workspace_bytes(int rows, int cols);
int tnpy_qr_split(const double* A, int rows, int cols, double* Q, double* T, double* defect_dev, int flags,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- layout helper: out[r, p, l] = in[l, p, r]  (mirror of a site tensor, used by the right
 * environment update so that the contracted bond is the slowest index) */
int tnpy_mirror_lpr(const double* in, double* out, int l, int d, int r, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TNPY_CUDA_H_ */
