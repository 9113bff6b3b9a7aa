// Internal interface of the tcgen05 path (csrc/ozaki.cu): FP64-accurate GEMMs as exact int8 slice GEMMs.
// Every buffer lives in a caller-supplied Workspace -- the library allocates nothing on the device.
#pragma once
#include "common.cuh"

namespace tnpy {

constexpr int kOzMaxSlices = 8;  // slices kept in memory per operand (7 bits each: 56 bits >= the 53 of a double)

// Row gather of a sliced operand: GEMM row k of the operand is source row
//   (k % kin) * kmul + k / kin + koff        of the row-major matrix P (row stride ld),
// which lets an environment tensor (bond, MPO channel, bond) be sliced MPO-channel-slowest and with leading /
// trailing channels dropped without a copy.  {K, 1, 0} is the plain matrix.
struct OzRowMap {
  int kin, kmul, koff;
};
inline OzRowMap oz_plain_rows(int K) { return OzRowMap{K, 1, 0}; }

// int8 slices of one GEMM operand, K-major: slices[s][col][k] (Kp bytes per row, Kp = K rounded up to 64),
// scale[col] = the power of two every entry of the column was divided by, colmax = scratch of the slicer,
// sumsq = sum_col scale[col]^2 (for the error bound of a product).
struct OzOperand {
  int8_t* slices;
  double* scale;
  unsigned long long* colmax;
  double* sumsq;
  int cols;
  int K;
  int64_t Kp;
};

int64_t oz_kp(int K);
size_t oz_operand_bytes(int cols, int K);
// carve an operand out of a workspace (false: workspace too small)
bool oz_operand_take(Workspace& ws, int cols, int K, OzOperand* out);
// slice P (K x cols after the row gather) into `op` with kOzMaxSlices slices
int oz_slice_operand(const double* P, int64_t ld, OzRowMap rows, const OzOperand& op, cudaStream_t stream);

// K chunks (64 k each) that whole blocks of tiles may skip because one operand is identically zero there -- the direct
// path's premixed operands have a zero (channel, q) piece wherever column q of that channel's d x d MPO block vanishes
// (S+ / S- blocks: half of them).  mode 1: block = m-tile / block_tiles, mode 2: block = n-tile / block_tiles; a
// block's tiles run the concatenation of its ranges [lo, hi) (in chunks) instead of [0, Kp / 64).
constexpr int kOzSkipBlocks = 4;
constexpr int kOzSkipRanges = 8;
struct OzKSkip {
  int mode, block_tiles, nblocks;
  int nranges[kOzSkipBlocks];
  int lo[kOzSkipBlocks][kOzSkipRanges], hi[kOzSkipBlocks][kOzSkipRanges];
};

// scratch the MMA kernel may need for the K-split tiles of its last wave
size_t oz_mma_scratch_bytes(int M, int N);
// out (+)= A^T B from sliced operands, using the first S slices of each (S in 6..8); *bound_dev, if given, is
// raised to max(*bound_dev, the rigorous normwise error bound of this product, see ozaki.cu)
int oz_mma(const OzOperand& A, const OzOperand& B, GemmOut out, int M, int N, int S, int accumulate, Workspace& ws,
           double* bound_dev, cudaStream_t stream, const OzKSkip* skip = nullptr);

// is a GEMM of this shape worth the tcgen05 path?  (below ~chi = 1024 the slicing passes cost more than they save)
bool ozaki_applicable(int M, int N, int K);
// slices used by products that nobody declared a tolerance for (8 unless tnpy_set_ozaki_slices changed it)
int ozaki_slices();

// H_eff in the mixed-canonical gauge for MPO tensors without interior-to-interior blocks (csrc/ozaki.cu, "direct
// path"): premixed, already sliced operands straight from x.
//   A side:  Xa[(b, ri), (q, m)] = sum_p W[0, b, p, q] x[m, p, ri]            b < wr - 1     (GEMM against R)
//   B side:  Xb[(a, li), (q, s)] = sum_p W[a + 1, wr - 1, p, q] x[li, p, s]   a < wl - 1     (GEMM against L)
// and y0[m, q, s] = sum_p W[0, wr - 1, p, q] x[m, p, s] - shift * x[m, q, s]  (the term with both identities).
constexpr int kPmMaxD = 4;    // physical dimension limit of the premix kernels
constexpr int kPmMaxCh = 16;  // MPO bond limit of the premix kernels
bool oz_premix_applicable(int l, int r, int wl, int wr, int d);
// skip_zero_pieces: do not write the (channel, q) pieces whose MPO block column vanishes -- only when the GEMM that
// consumes the operand is told to skip those K ranges (OzKSkip), their bytes are then never read.
int oz_premix_a(const double* x, const double* W, int l, int r, int wl, int wr, int d, const OzOperand& op, double* y0,
                const double* shift_dev, bool skip_zero_pieces, cudaStream_t stream);
int oz_premix_b(const double* x, const double* W, int l, int r, int wl, int wr, int d, const OzOperand& op,
                bool skip_zero_pieces, cudaStream_t stream);

}  // namespace tnpy
