// Blocked Cholesky factorisation + explicit triangular inverse of a (padded) symmetric positive definite matrix and
// the helpers around it (csrc/qr.cu); shared by the QR split and the dense pencil solver of ShiftInvertDMRG.
#pragma once
#include "common.cuh"

namespace tnpy {

constexpr int kCholBlock = 64;  // == NB of qr.cu
// n rounded up to a power-of-two number of 64-blocks (the recursive-doubling inverse wants that)
int chol_padded_dim(int n);
// G (np x np, leading n x n symmetric positive definite) -> D^-1 G D^-1, D = sqrt(diag G), identity on the padding;
// dinv (np) = 1 / D (1 on the padding).  A non-positive or non-finite diagonal entry raises *fail.
int spd_scale_pad(double* G, int n, int np, double* dinv, int* fail, cudaStream_t stream);
// G (np x np, lower part = SPD matrix; destroyed, holds the factor L afterwards) -> Cinv = L^-1 (np x np, lower).
// Tmp: np x np scratch, Dk: np x 64 scratch, *fail raised when a pivot is not positive.
int cholesky_inverse(double* G, int np, double* Cinv, double* Tmp, double* Dk, int* fail, cudaStream_t stream);
// out[c][r] = in[r][c] * (scale ? scale[c] : 1);  in: rows x cols (ld_in), out: cols x rows (ld_out)
int transpose(const double* in, int rows, int cols, int64_t ld_in, double* out, int64_t ld_out, const double* scale,
              cudaStream_t stream);

}  // namespace tnpy
