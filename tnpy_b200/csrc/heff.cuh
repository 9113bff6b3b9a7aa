// Internal interface of the contraction chains (csrc/contract.cu).
#pragma once
#include "ozaki.cuh"

namespace tnpy {

// How a prepared H_eff evaluates y = H_eff x (chosen once per plan from the sizes, the gauge flags, the MPO
// tensor's block structure and the GEMM selection):
//   FP64_CHAIN  GEMM (DMMA / generic) -> W-mix -> GEMM, native FP64                           (any operands)
//   OZ_CHAIN    the same chain with its big GEMMs on the tcgen05 int8 path; the environments' slices are made
//               once per plan, x and the mixed intermediate are sliced per call               (any operands)
//   OZ_DIRECT   mixed-canonical gauge (L[:,0,:] = R[:,wr-1,:] = I) and an MPO tensor whose non-zero blocks all
//               have a = 0 or b = wr - 1 (every nearest-neighbour Hamiltonian in tnpy.model without a penalty
//               term):  y = sum_{b<wr-1} (W_0b x) R_b + sum_{a>0} L_a^T (W_{a,wr-1} x) + W_{0,wr-1} x
//               -- two independent tcgen05 GEMMs whose x-side operands are premixed and sliced straight from x;
//               no FP64 intermediate exists, nothing of size w N touches HBM except the int8 slices.
enum HeffMode { HEFF_FP64_CHAIN = 0, HEFF_OZ_CHAIN = 1, HEFF_OZ_DIRECT = 2 };

struct HeffPlan {
  const double *L, *W, *R;
  int l, lo, row0, r, wl, wr, d, flags;  // lo rows of the bra bond starting at row0 (lo == l, row0 == 0: all of them)
  int mode;
  bool g1_oz, g3_oz;   // OZ_CHAIN: which of the two GEMMs run on the tcgen05 path
  OzOperand envL, envR;  // sliced constant operands (plan memory)
  OzKSkip skipR, skipL;  // direct path: K ranges the R-side / L-side GEMM skips (zero pieces of the premixed operands)
  OzKSkip skip3;         // tcgen05 chain: K ranges the second GEMM skips (zero pieces of the mixed intermediate)
  double* r2;            // FP64 chain with the right identity flag: R channel-major without its last channel
  double* bound;         // device scalar: largest rigorous error bound of a tcgen05 product issued through this plan
};

// bytes of plan memory / of per-call workspace, both upper bounds over the modes the arguments allow
size_t heff_plan_bytes(int l, int lo, int r, int wl, int wr, int d);
size_t heff_apply_bytes(int l, int lo, int r, int wl, int wr, int d);
// W_host: host copy of W (wl, wr, d, d) or NULL (then the library reads W back when the direct path is otherwise
// possible, synchronising the stream once).  algo: TNPY_GEMM_* (AUTO = the process-wide selection).
int heff_plan_init(HeffPlan* plan, const double* L, const double* W, const double* R, const double* W_host, int l,
                   int lo, int row0, int r, int wl, int wr, int d, int flags, int algo, Workspace& mem,
                   cudaStream_t stream);
// y = H_eff x - shift x.  slices: 7-bit slices per operand on the tcgen05 modes (0 = library default);
// shift_dev: device scalar or NULL.
int heff_plan_apply(const HeffPlan& plan, const double* x, double* y, int slices, const double* shift_dev,
                    Workspace& ws, cudaStream_t stream);

// one-shot convenience: ephemeral plan carved from `ws`, then apply
int heff_apply(const double* L, const double* W, const double* R, const double* x, double* y, int l, int r, int wl,
               int wr, int d, int flags, Workspace& ws, cudaStream_t stream);

// A^T B with the GEMM selection of the chains: tcgen05 when `algo` allows it, the shape is worth it and `ws` has
// room for the slices, native FP64 otherwise.
int chain_gemm(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
               int accumulate, int algo, Workspace& ws, cudaStream_t stream);
size_t chain_gemm_bytes(int M, int N, int K);

}  // namespace tnpy
