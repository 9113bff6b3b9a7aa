// FP64 GEMM  C[m,n] (+)= sum_k A[k,m] * B[k,n]  for the tnpy_b200 contraction chain.
//
// Every big contraction of the fDMRG local update (H_eff.psi, environment updates; reference
// matrix_product_state.py:296-336, :411-440) is brought to this one shape by keeping the
// contracted bond the slowest index of both operands, so both operands are "MN-major" and a
// single kernel serves the whole path (see DESIGN.md "Contraction chain").
//
// Two kernels:
//   * gemm_tn_dmma   -- the product path at scale.  Persistent CTAs (one per SM), a dedicated TMA
//                       producer warp feeding a multi-stage mbarrier ring of 128B-swizzled
//                       shared-memory slabs, consumer warps issuing FP64 tensor-core MMAs
//                       (mma.sync m8n8k4 f64 == SASS DMMA.8x8x4, the only FP64 tensor shape on
//                       sm_100a; tcgen05 has no f64 kind) with register accumulators.
//   * gemm_tn_generic -- any shape / any stride shared-memory tiled DFMA kernel, used for the tiny
//                       edge-of-chain bonds and for operands TMA cannot describe (odd leading
//                       dimension, unaligned base).  Still a CUDA kernel: there is no CPU path.
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"

namespace tnpy {

int forced_gemm_tile();

// =============================================================================================
// generic kernel
// =============================================================================================
template <int TS, int KS>
__global__ void __launch_bounds__(256) gemm_tn_generic(const double* __restrict__ A, int64_t lda,
                                                       const double* __restrict__ B, int64_t ldb,
                                                       GemmOut out, int M, int N, int K, int accumulate) {
  __shared__ double As[KS][TS + 1];
  __shared__ double Bs[KS][TS + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

  for (int k0 = 0; k0 < K; k0 += KS) {
    for (int idx = threadIdx.x; idx < KS * TS; idx += 256) {
      const int kk = idx / TS, c = idx % TS;
      const int k = k0 + kk;
      As[kk][c] = (k < K && m0 + c < M) ? A[(int64_t)k * lda + m0 + c] : 0.0;
      Bs[kk][c] = (k < K && n0 + c < N) ? B[(int64_t)k * ldb + n0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
    double* crow = out.C + (int64_t)(m / out.m_inner) * out.c_outer + (int64_t)(m % out.m_inner) * out.c_inner;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n < N) crow[n] = accumulate ? crow[n] + acc[i][j] : acc[i][j];
    }
  }
}

// =============================================================================================
// TMA + mbarrier + DMMA kernel
// =============================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
  // grouped rasterisation: 8 m-tiles share each sweep over n so that the CTAs resident at the same
  // time reuse A and B panels out of L2.
  constexpr int GROUP = 8;
  const int per_group = GROUP * tiles_n;
  const int gid = tile / per_group;
  const int first_m = gid * GROUP;
  const int gsz = min(tiles_m - first_m, GROUP);
  const int rem = tile - gid * per_group;
  tm = first_m + rem % gsz;
  tn = rem / gsz;
}

constexpr int kBK = 16;              // k rows per pipeline stage
constexpr int kSlabBytes = kBK * 128;  // one slab = kBK rows x 16 doubles (128 B, one swizzle span)

template <int BM, int BN, int WM, int WN, int STAGES>
struct DmmaCfg {
  static constexpr int kConsumerWarps = (BM / WM) * (BN / WN);
  // consumers fill whole warpgroups; one extra warpgroup hosts the producer warp so that
  // setmaxnreg (warpgroup-granular) can move registers from the producer to the consumers.
  static constexpr int kThreads = (kConsumerWarps + 4) * 32;
  static constexpr int kRegsConsumer = kConsumerWarps == 8 ? 232 : 232;
  static constexpr int kRegsProducer = 40;
  static constexpr int kASlabs = BM / 16, kBSlabs = BN / 16;
  static constexpr int kStageBytes = (kASlabs + kBSlabs) * kSlabBytes;
  static constexpr int kSmemBytes = STAGES * kStageBytes + 2 * STAGES * 8 + 1024;
};

template <int BM, int BN, int WM, int WN, int STAGES>
__global__ void __launch_bounds__(DmmaCfg<BM, BN, WM, WN, STAGES>::kThreads, 1)
    gemm_tn_dmma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmOut out,
                 int M, int N, int K, int accumulate, int vec_ok, int tiles_m, int tiles_n, int ksplit,
                 double* __restrict__ partial) {
  using Cfg = DmmaCfg<BM, BN, WM, WN, STAGES>;
  constexpr int NWN = BN / WN;
  constexpr int MI = WM / 8, NI = WN / 8;
  static_assert(WM % 16 == 0 && WN % 16 == 0, "warp tile must cover whole 16-wide slabs");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], Cfg::kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int KT = (K + kBK - 1) / kBK;
  const int total_tiles = tiles_m * tiles_n;
  // split-K (few output tiles, long K: e.g. the second GEMM of H_eff at chi = 256, 32 tiles x 96 k-steps on 148 SMs):
  // work item = (tile, K piece); piece ks writes its plain M x N partial to `partial + ks M N`, summed in a fixed order
  // by splitk_reduce_kernel afterwards.  ksplit == 1: the plain kernel.
  const int kt_per = (KT + ksplit - 1) / ksplit;
  const int total_work = total_tiles * ksplit;

  if (warp >= Cfg::kConsumerWarps) {
    // ------------------------------- TMA producer (one elected lane) --------------------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::kRegsProducer));
    if (warp == Cfg::kConsumerWarps && lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int tile = work % total_tiles, ks = work / total_tiles;
        int tm, tn;
        tile_coords(tile, tiles_m, tiles_n, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        const int a_slabs = min(Cfg::kASlabs, (M - m0 + 15) / 16);
        const int b_slabs = min(Cfg::kBSlabs, (N - n0 + 15) / 16);
        const uint32_t bytes = (a_slabs + b_slabs) * kSlabBytes;
        const int kt_end = min(KT, (ks + 1) * kt_per);
        for (int kt = ks * kt_per; kt < kt_end; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], bytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kASlabs * kSlabBytes;
          for (int s = 0; s < a_slabs; ++s) tma_load_2d(sa + s * kSlabBytes, &tmA, &full_bar[stage], m0 + 16 * s, kt * kBK);
          for (int s = 0; s < b_slabs; ++s) tma_load_2d(sb + s * kSlabBytes, &tmB, &full_bar[stage], n0 + 16 * s, kt * kBK);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // --------------------------------- DMMA consumers --------------------------------------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kRegsConsumer));
  const int wm_idx = warp / NWN, wn_idx = warp % NWN;
  const int g = lane >> 2, t = lane & 3;
  // Fragment element (row m0+g or n0+g of the 8-wide sub-tile, k = 2t+j of the 8-row k block) inside
  // a 128B-swizzled slab: byte = k*128 + ((chunk ^ (k & 7)) << 4) + (elem << 3), chunk = 4*half + g/2.
  uint32_t off[2][2];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 2 * t + j;
      off[j][h] = k * 128 + (((4 * h + (g >> 1)) ^ k) << 4) + ((g & 1) << 3);
    }

  const uint32_t smem_base = smem_u32(smem);
  uint32_t stage = 0, phase = 0;
  for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
    const int tile = work % total_tiles, ks = work / total_tiles;
    int tm, tn;
    tile_coords(tile, tiles_m, tiles_n, tm, tn);
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int n = 0; n < NI; ++n) acc[i][n][0] = acc[i][n][1] = 0.0;

    const int kt_end = min(KT, (ks + 1) * kt_per);
    for (int kt = ks * kt_per; kt < kt_end; ++kt) {
      mbar_wait(&full_bar[stage], phase);
      const uint32_t sa = smem_base + stage * Cfg::kStageBytes + (wm_idx * (WM / 16)) * kSlabBytes;
      const uint32_t sb = smem_base + stage * Cfg::kStageBytes + (Cfg::kASlabs + wn_idx * (WN / 16)) * kSlabBytes;
#pragma unroll
      for (int kb = 0; kb < kBK / 8; ++kb) {
        double a[MI][2], b[NI][2];
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            a[i][j] = lds_f64(sa + (i >> 1) * kSlabBytes + kb * 1024 + off[j][i & 1]);
#pragma unroll
        for (int n = 0; n < NI; ++n)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            b[n][j] = lds_f64(sb + (n >> 1) * kSlabBytes + kb * 1024 + off[j][n & 1]);
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < MI; ++i)
#pragma unroll
            for (int n = 0; n < NI; ++n) dmma884(acc[i][n][0], acc[i][n][1], a[i][j], b[n][j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }

    // epilogue: registers -> global (16-byte stores; each quad of lanes writes 64 contiguous bytes)
    const int row_base = tm * BM + wm_idx * WM + g;
    const int col_base = tn * BN + wn_idx * WN + 2 * t;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int m = row_base + 8 * i;
      if (m >= M) continue;
      double* crow = ksplit > 1 ? partial + ((int64_t)ks * M + m) * N
                                : out.C + (int64_t)(m / out.m_inner) * out.c_outer + (int64_t)(m % out.m_inner) * out.c_inner;
#pragma unroll
      for (int n = 0; n < NI; ++n) {
        const int col = col_base + 8 * n;
        if (col + 1 < N && vec_ok) {
          double2 v = make_double2(acc[i][n][0], acc[i][n][1]);
          double2* p = reinterpret_cast<double2*>(crow + col);
          if (accumulate) {
            const double2 o = *p;
            v.x += o.x;
            v.y += o.y;
          }
          *p = v;
        } else {
          if (col < N) crow[col] = accumulate ? crow[col] + acc[i][n][0] : acc[i][n][0];
          if (col + 1 < N) crow[col + 1] = accumulate ? crow[col + 1] + acc[i][n][1] : acc[i][n][1];
        }
      }
    }
  }
}

// out(m, n) (+)= sum over the K pieces of partial[ks][m][n], pieces in ascending order (bit-reproducible)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const double* __restrict__ partial, int ksplit, GemmOut out, int M,
                                                            int N, int accumulate) {
  const int64_t total = (int64_t)M * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(e / N), n = (int)(e % N);
    double* c = out.C + (int64_t)(m / out.m_inner) * out.c_outer + (int64_t)(m % out.m_inner) * out.c_inner + n;
    double v = accumulate ? *c : 0.0;
    for (int ks = 0; ks < ksplit; ++ks) v += partial[(int64_t)ks * total + e];
    *c = v;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
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

// Operand P: K rows x MN columns (row stride ld elements), box = 16 columns x kBK rows, 128B swizzle.
static int make_operand_map(CUtensorMap* map, const double* P, int64_t ld, int MN, int K) {
  auto enc = tensor_map_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return TNPY_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)MN, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {16, (cuuint32_t)kBK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(P), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p ld=%lld MN=%d K=%d)", (int)r, (const void*)P,
              (long long)ld, MN, K);
    return TNPY_ECUDA;
  }
  return TNPY_OK;
}

template <int BM, int BN, int WM, int WN, int STAGES>
static int launch_dmma(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmOut out, int M, int N, int K, int accumulate,
                       int vec_ok, int ksplit, double* partial, cudaStream_t stream) {
  using Cfg = DmmaCfg<BM, BN, WM, WN, STAGES>;
  auto kern = gemm_tn_dmma<BM, BN, WM, WN, STAGES>;
  TNPY_TRY(set_max_dynamic_smem(kern, Cfg::kSmemBytes));
  const int tiles_m = ceil_div(M, BM), tiles_n = ceil_div(N, BN);
  const int grid = min(tiles_m * tiles_n * ksplit, sm_count());
  if (ksplit > 1) {
    const int pvec = (reinterpret_cast<uintptr_t>(partial) % 16 == 0) && (N % 2 == 0);
    kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, out, M, N, K, 0, pvec, tiles_m, tiles_n, ksplit, partial);
    TNPY_LAUNCH_OK();
    const int64_t total = (int64_t)M * N;
    const int64_t want = (total + 255) / 256, cap = (int64_t)sm_count() * 8;
    const int rgrid = (int)(want < cap ? want : cap);
    splitk_reduce_kernel<<<rgrid, 256, 0, stream>>>(partial, ksplit, out, M, N, accumulate);
    TNPY_LAUNCH_OK();
    return TNPY_OK;
  }
  kern<<<grid, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(tmA, tmB, out, M, N, K, accumulate, vec_ok, tiles_m, tiles_n, 1, nullptr);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// Tile choice: minimise (waves x tile area / relative efficiency) over the compiled configurations.
static int pick_tile(int M, int N) {
  const int sms = sm_count();
  auto cost = [&](int bm, int bn, double eff) {
    const int64_t tiles = (int64_t)ceil_div(M, bm) * ceil_div(N, bn);
    const int64_t waves = (tiles + sms - 1) / sms;
    return (double)waves * bm * bn / eff;
  };
  double c0 = cost(128, 128, 1.00), c1 = cost(128, 64, 0.93), c2 = cost(64, 64, 0.80);
  const int forced = forced_gemm_tile();
  if (forced == 0) c0 = -1.0;
  if (forced == 1) { c1 = -1.0; c0 = 1e300; }
  if (forced == 2) { c2 = -1.0; c0 = c1 = 1e300; }
  if (c0 <= c1 && c0 <= c2) return 0;
  return c1 <= c2 ? 1 : 2;
}

// K pieces for the DMMA kernel: when the output has fewer tiles than half the SMs and K is long, cut K so that the
// work items fill the machine (at most 8 pieces, each at least 128 rows of K)
static int pick_ksplit(int M, int N, int K, int tile) {
  const int bm = tile == 2 ? 64 : 128, bn = tile == 0 ? 128 : 64;
  const int64_t tiles = (int64_t)ceil_div(M, bm) * ceil_div(N, bn);
  const int sms = sm_count();
  if (K < 512 || tiles * 2 > sms) return 1;
  int ks = (int)(sms / tiles);
  if (ks > 8) ks = 8;
  if (ks > K / 128) ks = K / 128;
  return ks < 2 ? 1 : ks;
}

size_t gemm_splitk_doubles(int M, int N, int K) {
  const bool tiny = (int64_t)M * N < 64 * 64 || K < 16 || M < 32 || N < 32;
  if (tiny) return 0;
  const int ks = pick_ksplit(M, N, K, pick_tile(M, N));
  return ks > 1 ? (size_t)ks * M * N : 0;
}

static bool tma_ok(const double* P, int64_t ld) {
  return (reinterpret_cast<uintptr_t>(P) % 16 == 0) && (ld % 2 == 0) && ld > 0;
}

int gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
            int accumulate, int algo, cudaStream_t stream) {
  return gemm_tn_ws(A, lda, B, ldb, out, M, N, K, accumulate, algo, nullptr, 0, stream);
}

// The same with scratch for the split-K partials (gemm_splitk_doubles(M, N, K) doubles; fewer: no split)
int gemm_tn_ws(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
               int accumulate, int algo, double* scratch, size_t scratch_doubles, cudaStream_t stream) {
  TNPY_CHECK_ARG(A && B && out.C, "null operand");
  TNPY_CHECK_ARG(M > 0 && N > 0 && K > 0, "non-positive dimension");
  TNPY_CHECK_ARG(lda >= M && ldb >= N && out.m_inner > 0, "leading dimension too small");
  // This entry point is native FP64 (the tcgen05 path needs a workspace for its slices: chain_gemm in
  // contract.cu / tnpy_ozaki_gemm_tn); the process-wide selection only matters when it forces a kernel.
  if (algo == TNPY_GEMM_AUTO) algo = current_gemm_algo();
  if (algo == TNPY_GEMM_OZAKI || algo == TNPY_GEMM_FP64) algo = TNPY_GEMM_AUTO;
  const bool can_tma = tma_ok(A, lda) && tma_ok(B, ldb);
  if (algo == TNPY_GEMM_DMMA && !can_tma) {
    set_error("gemm_tn: TNPY_GEMM_DMMA requested but operands are not TMA-describable (16B base, even ld)");
    return TNPY_EINVAL;
  }
  // Tiny problems (edge-of-chain bonds) are latency-bound: one launch of the generic kernel.
  const bool tiny = (int64_t)M * N < 64 * 64 || K < 16 || M < 32 || N < 32;
  if (algo == TNPY_GEMM_GENERIC || (algo == TNPY_GEMM_AUTO && (!can_tma || tiny))) {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
    gemm_tn_generic<64, 16><<<grid, 256, 0, stream>>>(A, lda, B, ldb, out, M, N, K, accumulate);
    TNPY_LAUNCH_OK();
    return TNPY_OK;
  }
  CUtensorMap tmA, tmB;
  TNPY_TRY(make_operand_map(&tmA, A, lda, M, K));
  TNPY_TRY(make_operand_map(&tmB, B, ldb, N, K));
  const int vec_ok = (reinterpret_cast<uintptr_t>(out.C) % 16 == 0) && (out.c_inner % 2 == 0) && (out.c_outer % 2 == 0);
  const int tile = pick_tile(M, N);
  int ksplit = pick_ksplit(M, N, K, tile);
  if (ksplit > 1 && (scratch == nullptr || scratch_doubles < (size_t)ksplit * M * N)) ksplit = 1;
  if (tile == 0) return launch_dmma<128, 128, 64, 32, 4>(tmA, tmB, out, M, N, K, accumulate, vec_ok, ksplit, scratch, stream);
  if (tile == 1) return launch_dmma<128, 64, 32, 32, 6>(tmA, tmB, out, M, N, K, accumulate, vec_ok, ksplit, scratch, stream);
  return launch_dmma<64, 64, 32, 16, 8>(tmA, tmB, out, M, N, K, accumulate, vec_ok, ksplit, scratch, stream);
}

}  // namespace tnpy

extern "C" int tnpy_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int M,
                            int N, int K, int accumulate, int algo, void* stream) {
  TNPY_CHECK_ARG(C != nullptr && ldc >= N, "bad C / ldc");
  return tnpy::gemm_tn(A, lda, B, ldb, tnpy::plain_out(C, ldc, M), M, N, K, accumulate, algo,
                       static_cast<cudaStream_t>(stream));
}
