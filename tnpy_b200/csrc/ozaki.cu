// EXPERIMENT (reported separately, never substituted silently): FP64-accurate GEMM on the 5th-gen
// tensor cores.  tcgen05 has no f64 kind, so C = A^T B is evaluated with the Ozaki scheme: every
// operand column is scaled by a power of two and split error-free into S signed 7-bit slices,
//     A[k][m] = 2^ea[m] * sum_i a_i[k][m] * 2^(-7(i+1)),    |a_i| <= 64,
// all slice products a_i^T b_j are exact int8 x int8 -> int32 GEMMs (tcgen05.mma.kind::i8, accumulators
// in TMEM), slice pairs of equal weight i + j = d share one TMEM accumulator (exact: < 2^31 for
// K <= 65536), pairs with i + j >= S are below the target precision and skipped, and the epilogue
// recombines the S accumulators in FP64:  C[m][n] = 2^(ea[m]+eb[n]) * sum_d acc_d[m][n] * 2^(-7(d+2)).
//
// Kernel anatomy (one 128 x 64 output tile per CTA, 256 threads):
//   warp 0   TMA producer: one 3-D box (k, rows, slices) per operand per stage, 64B-swizzled, K-major
//   warp 1   MMA issuer (one lane): S(S+1)/2 slice pairs x 2 k-steps of tcgen05.mma per stage,
//            tcgen05.commit to the stage's empty barrier, final commit to the epilogue barrier
//   warp 2   TMEM allocator (512 columns = S accumulators of 64 int32 columns)
//   warps 4-7 epilogue: tcgen05.ld 32x32b, FP64 recombination, scaled store through the GEMM row map
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace tnpy {

constexpr int kOzBM = 128;   // tile rows (TMEM lanes)
constexpr int kOzBN = 64;    // tile columns per accumulator
constexpr int kOzBK = 64;    // k elements (= bytes) per stage row: one 64B swizzle span
constexpr int kOzStages = 2;  // measured alternatives: 4 stages of 32-byte rows -10 %, 32-column epilogue loads -4 %
constexpr int kOzMaxSlices = 8;  // 8 accumulators x 64 columns = all 512 TMEM columns

// ---------------------------------------------------------------------------------------------
// slicing
// ---------------------------------------------------------------------------------------------
// colmax[c] = max_k |P[k][c]| as the bit pattern of a non-negative double (integer order == value
// order), reduced over k-chunks with atomicMax; colmax must be zeroed first.
__global__ void __launch_bounds__(256) oz_colmax_kernel(const double* __restrict__ P, int64_t ld, int K, int MN,
                                                        int k_chunk, unsigned long long* __restrict__ colmax) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= MN) return;
  const int k0 = blockIdx.y * k_chunk, k1 = min(K, k0 + k_chunk);
  double mx = 0.0;
  for (int k = k0; k < k1; ++k) mx = fmax(mx, fabs(P[(int64_t)k * ld + c]));
  if (mx > 0.0) atomicMax(&colmax[c], (unsigned long long)__double_as_longlong(mx));
}

// scale = 2^e with |P[k][c]| / 2^e <= 0.5 for all k  (0 for an all-zero column)
__device__ __forceinline__ double oz_scale_of(unsigned long long bits) {
  const double mx = __longlong_as_double((long long)bits);
  return mx > 0.0 ? ldexp(1.0, ilogb(mx) + 2) : 0.0;
}

// slices[s][c][k] (k contiguous, Kp bytes per row) from P[k][c]: 32 columns x 128 k per block, digits staged
// in shared memory so that every (slice, column) row leaves as one 128-byte line.
template <int S>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ P, int64_t ld, int K, int MN,
                                                       const unsigned long long* __restrict__ colmax,
                                                       double* __restrict__ scale, int8_t* __restrict__ slices,
                                                       int64_t Kp, int64_t slice_stride) {
  __shared__ __align__(16) int8_t tile[S][32][132];
  const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 128;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  const int c = c0 + tx;
  const double sc = c < MN ? oz_scale_of(colmax[c]) : 0.0;
  const double inv = sc > 0.0 ? 1.0 / sc : 0.0;  // exact: power of two
  if (blockIdx.y == 0 && ty == 0 && c < MN) scale[c] = sc;
  // digit = rint(128 t) without conversion instructions: adding 1.5 * 2^52 leaves the rounded integer in the low
  // mantissa bits (two's complement), subtracting it again gives the rounded value; the remainder is exact.
  const double magic = 6755399441055744.0;
#pragma unroll 4
  for (int i = ty; i < 128; i += 8) {
    const int k = k0 + i;
    double t = (k < K && c < MN) ? P[(int64_t)k * ld + c] * inv : 0.0;  // |t| <= 0.5
#pragma unroll
    for (int sl = 0; sl < S; ++sl) {
      const double u = fma(t, 128.0, magic);
      tile[sl][tx][i] = (int8_t)__double2loint(u);  // |digit| <= 64
      t = fma(t, 128.0, magic - u);                 // exact remainder, |t| <= 0.5
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * 32 * 32; idx += 256) {
    const int w = idx % 32, cc = (idx / 32) % 32, sl = idx / 1024;
    const int64_t k = k0 + 4 * w;
    if (c0 + cc < MN && k < Kp)
      *reinterpret_cast<int32_t*>(slices + sl * slice_stride + (int64_t)(c0 + cc) * Kp + k) =
          *reinterpret_cast<const int32_t*>(&tile[sl][cc][4 * w]);
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = oz_smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void oz_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(oz_smem_u32(dst)), "l"(map), "r"(oz_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major operand tile, rows of kOzBK bytes, 64B swizzle: SBO = 8 rows * 64 B, LBO unused, version 1
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);             // start address
  d |= (uint64_t)0 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * kOzBK) >> 4) << 32;            // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;                             // layout type: SWIZZLE_64B
  return d;
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}\n"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar))
               : "memory");
}

template <int S>
__global__ void __launch_bounds__(256, 1)
    oz_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const double* __restrict__ scaleA, const double* __restrict__ scaleB, GemmOut out, int M, int N,
                  int KT, int accumulate) {
  constexpr int kABytes = S * kOzBM * kOzBK, kBBytes = S * kOzBN * kOzBK;
  constexpr int kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOzStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kOzStages;
  uint64_t* tmem_full = empty_bar + kOzStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grouped rasterisation (8 m-tiles per sweep over n) so that co-resident CTAs share operand panels in L2
  int tm, tn;
  {
    const int tiles_m = (M + kOzBM - 1) / kOzBM, tiles_n = (N + kOzBN - 1) / kOzBN;
    constexpr int GROUP = 8;
    const int tile = blockIdx.x, per_group = GROUP * tiles_n;
    const int gid = tile / per_group, first_m = gid * GROUP;
    const int gsz = min(tiles_m - first_m, GROUP), rem = tile - gid * per_group;
    tm = first_m + rem % gsz;
    tn = rem / gsz;
  }
  const int m0 = tm * kOzBM, n0 = tn * kOzBN;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kOzStages; ++s) {
      oz_mbar_init(&full_bar[s], 1);
      oz_mbar_init(&empty_bar[s], 1);
    }
    oz_mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kt = 0; kt < KT; ++kt) {
        oz_mbar_wait(&empty_bar[stage], phase ^ 1);
        oz_mbar_expect_tx(&full_bar[stage], kStageBytes);
        uint8_t* sa = smem + stage * kStageBytes;
        oz_tma_3d(sa, &tmA, &full_bar[stage], kt * kOzBK, m0, 0);
        oz_tma_3d(sa + kABytes, &tmB, &full_bar[stage], kt * kOzBK, n0, 0);
        if (++stage == kOzStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 64, M = 128
      constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kOzBN >> 3) << 17) | ((uint32_t)(kOzBM >> 4) << 24);
      uint32_t stage = 0, phase = 0;
      for (int kt = 0; kt < KT; ++kt) {
        oz_mbar_wait(&full_bar[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = oz_smem_u32(smem + stage * kStageBytes);
        const uint32_t sb = sa + kABytes;
#pragma unroll
        for (int kk = 0; kk < kOzBK / 32; ++kk) {
#pragma unroll
          for (int i = 0; i < S; ++i) {
            const uint64_t da = oz_smem_desc(sa + i * (kOzBM * kOzBK) + kk * 32);
#pragma unroll
            for (int j = 0; j < S - i; ++j) {
              const uint64_t db = oz_smem_desc(sb + j * (kOzBN * kOzBK) + kk * 32);
              // accumulator d = i + j; its first contribution in program order is (kt, kk, i) = (0, 0, 0)
              oz_mma_i8(tmem_base + (uint32_t)((i + j) * kOzBN), da, db, idesc, (kt | kk | i) != 0 ? 1u : 0u);
            }
          }
        }
        oz_commit(&empty_bar[stage]);  // frees the stage once these MMAs have read it
        if (++stage == kOzStages) { stage = 0; phase ^= 1; }
      }
      oz_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    oz_mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = m0 + q * 32 + lane;
    const double sa = (m < M) ? scaleA[m] : 0.0;
    double* crow = nullptr;
    if (m < M) crow = out.C + (int64_t)(m / out.m_inner) * out.c_outer + (int64_t)(m % out.m_inner) * out.c_inner;
    for (int c0 = 0; c0 < kOzBN; c0 += 8) {
      double acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.0;
#pragma unroll
      for (int d = S - 1; d >= 0; --d) {  // smallest weights first
        uint32_t v[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(d * kOzBN + c0);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const double wgt = ldexp(1.0, -7 * (d + 2));
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fma((double)(int)v[c], wgt, acc[c]);
      }
      if (m < M) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int n = n0 + c0 + c;
          if (n < N) {
            const double val = acc[c] * sa * scaleB[n];
            crow[n] = accumulate ? crow[n] + val : val;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel, second generation: CTA pair (cta_group::2), 256 x 128 tile, two accumulator passes
// ---------------------------------------------------------------------------------------------
// ncu on oz_mma_kernel (profiles/r01_oz_mma_kernel_ncu_full_raw.csv) shows the tensor cores' shared-memory
// operand pipe at 74-89 % of peak with the MMA pipe only 50-60 % busy: an M=128, N=64, K=32 int8 MMA reads
// 4 KB of A + 2 KB of B = 48 wavefronts of 128 B for 32 cycles of math.  N per instruction is capped by TMEM
// (S accumulators x N columns <= 512), so this kernel (a) runs the S diagonals in two passes of <= 4
// accumulators, which allows N = 128, and (b) pairs two CTAs (cta_group::2, M = 256): each CTA stages its own
// 128 rows of A and only half (64 rows) of B, i.e. 4 + 2 KB per 64 cycles of math = 75 % of the operand pipe.
//
//   pass 0: diagonals d = 4 .. S-1 (small weights; needs all slices: two 48 KB slots per 64-byte K chunk)
//   pass 1: diagonals d = 0 .. 3   (needs slices 0..3 of both operands: one slot per K chunk)
//
// Shared-memory ring: 4 slots of {A slices[4][128 rows][64 B], B slices[4][64 rows][64 B]} per CTA.  Both CTAs'
// TMA loads complete on the leader's full barrier; the leader's elected thread issues every MMA and frees
// slots / publishes accumulators in both CTAs with multicast commits; each CTA's epilogue warps drain their
// own 128 TMEM lanes (pass 0 writes C, pass 1 adds to it) and release TMEM to the leader between the passes.
constexpr int kOz2SlotA = 4 * kOzBM * kOzBK;  // 32 KB
constexpr int kOz2SlotB = 4 * kOzBN * kOzBK;  // 16 KB
constexpr int kOz2Slot = kOz2SlotA + kOz2SlotB;
constexpr int kOz2Slots = 4;
constexpr int kOz2TileN = 2 * kOzBN;  // 128 columns per accumulator
constexpr int kOz2EpiWarps = 8;       // two per TMEM lane quadrant
constexpr int kOz2PartDepth = 4;      // 16-column chunks of C requested ahead of use in the epilogue

// Tail splitting: tiles of the last, partly filled wave are cut along K (see oz2_mma_kernel).
struct Oz2Tail {
  int n_full;       // work items [0, n_full) are whole tiles
  int rem;          // number of split tiles (tile indices n_full .. n_full + rem - 1)
  int splits;       // K ranges per split tile
  double* scratch;  // (splits - 1) * rem plain 256 x 128 tiles receiving the K ranges ks > 0
};
constexpr int kOz2Threads = 128 + 32 * kOz2EpiWarps;

__device__ __forceinline__ uint32_t oz_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void oz_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t oz_mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void oz_tma_3d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1,
                                               int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(oz_smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void oz_tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void oz_mma_i8_pair(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n}\n"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z), "r"(z),
      "r"(z), "r"(z)
      : "memory");
}
// arrive (once the preceding MMAs have completed) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void oz_commit_pair(uint64_t* bar) {
  asm volatile(
      "{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %1;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n}\n"
      ::"r"(oz_smem_u32(bar)), "r"(3u)
      : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// Issue the MMAs of one 64-byte K chunk for pass PASS (0: diagonals 4..S-1, 1: diagonals 0..3).  lo4 / hi4 are the
// (address >> 4) fields of the slot(s) holding slice groups 0..3 / 4..7; everything else is a compile-time
// constant so the single issuing thread spends two integer adds per MMA.
template <int S, int PASS>
__device__ __forceinline__ void oz2_issue_chunk(uint32_t lo4, uint32_t hi4, uint32_t tmem_base, bool first_chunk) {
  constexpr int d_lo = PASS == 0 ? 4 : 0, d_hi = PASS == 0 ? S - 1 : 3;
  // instruction descriptor: D = S32, A = B = signed int8, both K-major, N = 128, M = 256 (128 rows per CTA)
  constexpr uint32_t idesc =
      (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kOz2TileN >> 3) << 17) | ((uint32_t)((2 * kOzBM) >> 4) << 24);
  // K-major, 64B swizzle: SBO = 8 rows * 64 B, descriptor version 1, layout type SWIZZLE_64B
  constexpr uint64_t desc_hi = ((uint64_t)((8 * kOzBK) >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
#pragma unroll
  for (int kk = 0; kk < kOzBK / 32; ++kk) {
#pragma unroll
    for (int i = 0; i < S; ++i) {
      const uint64_t da = desc_hi | (uint64_t)((i < 4 ? lo4 : hi4) + (uint32_t)(((i & 3) * (kOzBM * kOzBK) + kk * 32) >> 4));
#pragma unroll
      for (int j = 0; j < S; ++j) {
        if (i + j < d_lo || i + j > d_hi) continue;
        const uint64_t db = desc_hi | (uint64_t)((j < 4 ? lo4 : hi4) +
                                                 (uint32_t)((kOz2SlotA + (j & 3) * (kOzBN * kOzBK) + kk * 32) >> 4));
        // the first contribution to every diagonal of a pass comes from slice i = 0 at k-step 0 of chunk 0
        const uint32_t acc = (kk == 0 && i == 0) ? (first_chunk ? 0u : 1u) : 1u;
        oz_mma_i8_pair(tmem_base + (uint32_t)((i + j - d_lo) * kOz2TileN), da, db, idesc, acc);
      }
    }
  }
}

template <int S, int PASS>
__device__ __forceinline__ void oz2_issue_pass(uint32_t smem0, uint64_t* full_bar, uint64_t* empty_bar, uint64_t* tmem_full,
                                               uint32_t tmem_base, int KT, uint32_t& slot, uint32_t& phase) {
  for (int kt = 0; kt < KT; ++kt) {
    const uint32_t lo = smem0 + slot * kOz2Slot;
    oz_mbar_wait(&full_bar[slot], phase);
    if (PASS == 0) oz_mbar_wait(&full_bar[slot + 1], phase);  // slots are taken in pairs (0,1) / (2,3): same phase
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    oz2_issue_chunk<S, PASS>(lo >> 4, (lo + kOz2Slot) >> 4, tmem_base, kt == 0);
    oz_commit_pair(&empty_bar[slot]);
    if (PASS == 0) oz_commit_pair(&empty_bar[slot + 1]);
    slot += PASS == 0 ? 2 : 1;
    if (slot == kOz2Slots) { slot = 0; phase ^= 1; }
  }
  oz_commit_pair(tmem_full);
}

template <int S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kOz2Threads, 1)
    oz2_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const double* __restrict__ scaleA, const double* __restrict__ scaleB, GemmOut out, int M, int N,
                   int KT, int accumulate, Oz2Tail tail) {
  static_assert(S > 4 && S <= 8, "two passes of at most four diagonals");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOz2Slots * kOz2Slot);
  uint64_t* empty_bar = full_bar + kOz2Slots;
  uint64_t* tmem_full = empty_bar + kOz2Slots;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = oz_cluster_rank();
  int tm, tn, ks = 0, kt0 = 0, kt1 = KT;
  double* tail_tile = nullptr;
  {
    const int tiles_m = (M + 2 * kOzBM - 1) / (2 * kOzBM), tiles_n = (N + kOz2TileN - 1) / kOz2TileN;
    constexpr int GROUP = 4;
    // work items: the first tail.n_full items are whole tiles; the tiles of the last, partly filled wave are cut
    // into tail.splits K ranges each so that the wave fills the machine (part ks > 0 goes to a scratch tile)
    int tile = blockIdx.x >> 1;
    if (tile >= tail.n_full) {
      const int j = tile - tail.n_full;
      tile = tail.n_full + j % tail.rem;
      ks = j / tail.rem;
      kt0 = (int)((int64_t)KT * ks / tail.splits);
      kt1 = (int)((int64_t)KT * (ks + 1) / tail.splits);
      if (ks > 0) tail_tile = tail.scratch + (int64_t)((ks - 1) * tail.rem + (tile - tail.n_full)) * (2 * kOzBM * kOz2TileN);
    }
    const int per_group = GROUP * tiles_n;
    const int gid = tile / per_group, first_m = gid * GROUP;
    const int gsz = min(tiles_m - first_m, GROUP), rem = tile - gid * per_group;
    tm = first_m + rem % gsz;
    tn = rem / gsz;
  }
  const int m0 = tm * 2 * kOzBM + (int)rank * kOzBM;  // this CTA's rows of C / columns of A
  const int n0 = tn * kOz2TileN;                      // the pair's columns of C
  const int nb0 = n0 + (int)rank * kOzBN;             // the half of the B tile this CTA stages
  if (threadIdx.x == 0) {
    for (int s = 0; s < kOz2Slots; ++s) {
      oz_mbar_init(&full_bar[s], 1);
      oz_mbar_init(&empty_bar[s], 1);
    }
    oz_mbar_init(tmem_full, 1);
    oz_mbar_init(tmem_empty, 2 * kOz2EpiWarps);  // every epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  oz_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // load sequence: t in [0, 2 KT) = pass 0 (chunk t / 2, slice group t % 2), t in [2 KT, 3 KT) = pass 1 (group 0)
      const int nk = kt1 - kt0, T = 3 * nk;
      uint32_t slot = 0, phase = 0;
      for (int t = 0; t < T; ++t) {
        int kt, sub;
        if (t < 2 * nk) { kt = kt0 + (t >> 1); sub = t & 1; } else { kt = kt0 + t - 2 * nk; sub = 0; }
        oz_mbar_wait(&empty_bar[slot], phase ^ 1);
        if (rank == 0) oz_mbar_expect_tx(&full_bar[slot], 2 * kOz2Slot);  // both CTAs' boxes land on this barrier
        const uint32_t leader_full = oz_mapa(oz_smem_u32(&full_bar[slot]), 0);
        uint8_t* dst = smem + slot * kOz2Slot;
        oz_tma_3d_pair(dst, &tmA, leader_full, kt * kOzBK, m0, sub * 4);
        oz_tma_3d_pair(dst + kOz2SlotA, &tmB, leader_full, kt * kOzBK, nb0, sub * 4);
        if (++slot == kOz2Slots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t smem0 = oz_smem_u32(smem);
      uint32_t slot = 0, phase = 0;
      oz2_issue_pass<S, 0>(smem0, full_bar, empty_bar, tmem_full, tmem_base, kt1 - kt0, slot, phase);
      oz_mbar_wait(tmem_empty, 0);  // both CTAs' epilogues have drained the pass-0 accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      oz2_issue_pass<S, 1>(smem0, full_bar, empty_bar, tmem_full, tmem_base, kt1 - kt0, slot, phase);
    }
  } else if (warp >= 4) {
    // Epilogue: 8 warps; warp w drains TMEM lanes 32 (w % 4) + 16 ((w - 4) / 4) .. + 15 with 16x256b loads, whose
    // fragment is the mma accumulator layout: thread t holds lanes t/4 and t/4 + 8, columns 8 b + 2 (t % 4) + {0, 1}
    // of every 8-column block b -- four threads cover 64 contiguous bytes of a C row, so global accesses are
    // whole sectors without a shared-memory transpose.
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int lane_base = q * 32 + half * 16;
    const int r0 = lane_base + (lane >> 2), r1 = r0 + 8;
    const int mrow0 = m0 + r0, mrow1 = m0 + r1;
    const double sa0 = (mrow0 < M) ? scaleA[mrow0] : 0.0, sa1 = (mrow1 < M) ? scaleA[mrow1] : 0.0;
    // row pointers are biased by the tile's first column: element (row, n0 + c) lives at crow[c]
    double* crow0 = nullptr;
    double* crow1 = nullptr;
    if (tail_tile != nullptr) {  // K part ks > 0 of a split tile: plain 256 x 128 scratch tile
      if (mrow0 < M) crow0 = tail_tile + (int64_t)((int)rank * kOzBM + r0) * kOz2TileN;
      if (mrow1 < M) crow1 = tail_tile + (int64_t)((int)rank * kOzBM + r1) * kOz2TileN;
    } else {
      if (mrow0 < M)
        crow0 = out.C + (int64_t)(mrow0 / out.m_inner) * out.c_outer + (int64_t)(mrow0 % out.m_inner) * out.c_inner + n0;
      if (mrow1 < M)
        crow1 = out.C + (int64_t)(mrow1 / out.m_inner) * out.c_outer + (int64_t)(mrow1 % out.m_inner) * out.c_inner + n0;
    }
    const int cpair = 2 * (lane & 3);
    const int ncols = min(kOz2TileN, N - n0);  // valid columns of this tile
    // 16-byte accesses when every pair this thread touches is aligned and in range (uniform per thread)
    const bool vec = (ncols == kOz2TileN) && ((reinterpret_cast<uintptr_t>(crow0) | reinterpret_cast<uintptr_t>(crow1)) & 15) == 0;
    const uint32_t leader_empty = oz_mapa(oz_smem_u32(tmem_empty), 0);
    for (int pass = 0; pass < 2; ++pass) {
      const int d_lo = pass == 0 ? 4 : 0, nd = pass == 0 ? S - 4 : 4;
      const bool add = pass == 1 || ((accumulate & 1) && tail_tile == nullptr);
      double wgt[4];
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) wgt[dd] = ldexp(1.0, -7 * (d_lo + dd + 2));
      // Current values of the elements this thread owns (read-modify-write of pass 1), kept kOz2PartDepth column
      // chunks ahead of their use; the first chunks are requested while the MMAs of this pass are still running.
      double part[kOz2PartDepth][2][2][2];  // [chunk ring][row][block][column of the pair]
      auto load_part = [&](int slot, int c0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int c = c0 + 8 * b + cpair;
          if (!add) {
            part[slot][0][b][0] = part[slot][0][b][1] = part[slot][1][b][0] = part[slot][1][b][1] = 0.0;
          } else if (vec) {
            const double2 z = make_double2(0.0, 0.0);
            const double2 p0 = crow0 ? *reinterpret_cast<const double2*>(crow0 + c) : z;
            const double2 p1 = crow1 ? *reinterpret_cast<const double2*>(crow1 + c) : z;
            part[slot][0][b][0] = p0.x; part[slot][0][b][1] = p0.y;
            part[slot][1][b][0] = p1.x; part[slot][1][b][1] = p1.y;
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              part[slot][0][b][e] = (crow0 != nullptr && c + e < ncols) ? crow0[c + e] : 0.0;
              part[slot][1][b][e] = (crow1 != nullptr && c + e < ncols) ? crow1[c + e] : 0.0;
            }
          }
        }
      };
#pragma unroll
      for (int pc = 0; pc < kOz2PartDepth; ++pc) load_part(pc, 16 * pc);
      oz_mbar_wait(tmem_full, (uint32_t)pass);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int ci = 0; ci < kOz2TileN / 16; ++ci) {
        const int c0 = 16 * ci;
        uint32_t v[4][8];
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          if (dd < nd) {
            const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(dd * kOz2TileN + c0);
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[dd][0]), "=r"(v[dd][1]), "=r"(v[dd][2]), "=r"(v[dd][3]), "=r"(v[dd][4]), "=r"(v[dd][5]),
                           "=r"(v[dd][6]), "=r"(v[dd][7])
                         : "r"(taddr));
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double res[2][2][2];
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = c0 + 8 * b + cpair + e;
            const double sb = c < ncols ? scaleB[n0 + c] : 0.0;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {  // registers 4 b + 2 rr + e: lane t/4 + 8 rr, column 8 b + 2 (t % 4) + e
              double acc = 0.0;
#pragma unroll
              for (int dd = 3; dd >= 0; --dd)  // smallest weights first
                if (dd < nd) acc = fma((double)(int)v[dd][4 * b + 2 * rr + e], wgt[dd], acc);
              res[rr][b][e] = fma(acc * (rr == 0 ? sa0 : sa1), sb, part[ci % kOz2PartDepth][rr][b][e]);
            }
          }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int c = c0 + 8 * b + cpair;
          if (vec) {
            if (crow0) *reinterpret_cast<double2*>(crow0 + c) = make_double2(res[0][b][0], res[0][b][1]);
            if (crow1) *reinterpret_cast<double2*>(crow1 + c) = make_double2(res[1][b][0], res[1][b][1]);
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (crow0 != nullptr && c + e < ncols) crow0[c + e] = res[0][b][e];
              if (crow1 != nullptr && c + e < ncols) crow1[c + e] = res[1][b][e];
            }
          }
        }
        if (ci + kOz2PartDepth < kOz2TileN / 16) load_part(ci % kOz2PartDepth, c0 + 16 * kOz2PartDepth);
      }
      if (pass == 0) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) oz_mbar_arrive_remote(leader_empty);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  oz_cluster_sync();  // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 oz_encoder() {
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

// slices[s][row][k]: dims (Kp, rows, S), box (kOzBK, box_rows, S), 64B swizzle
static int oz_make_map(CUtensorMap* map, const int8_t* base, int64_t Kp, int rows, int S, int box_rows,
                       int box_slices = 0) {
  if (box_slices <= 0) box_slices = S;
  auto enc = oz_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return TNPY_ECUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * rows};
  cuuint32_t box[3] = {(cuuint32_t)kOzBK, (cuuint32_t)box_rows, (cuuint32_t)box_slices};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ozaki: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TNPY_ECUDA;
  }
  return TNPY_OK;
}

// `scale` doubles as scratch for the column maxima: the first MN 8-byte words of `colmax_scratch`.
template <int S>
static int oz_slice(const double* P, int64_t ld, int K, int MN, double* scale, unsigned long long* colmax_scratch,
                    int8_t* slices, int64_t Kp, cudaStream_t stream) {
  TNPY_CUDA_OK(cudaMemsetAsync(colmax_scratch, 0, sizeof(unsigned long long) * (size_t)MN, stream));
  const int k_chunk = K > 4096 ? 256 : 64;
  dim3 g1(ceil_div(MN, 256), ceil_div(K, k_chunk));
  oz_colmax_kernel<<<g1, 256, 0, stream>>>(P, ld, K, MN, k_chunk, colmax_scratch);
  TNPY_LAUNCH_OK();
  dim3 grid(ceil_div(MN, 32), (unsigned)((Kp + 127) / 128));
  oz_slice_kernel<S><<<grid, 256, 0, stream>>>(P, ld, K, MN, colmax_scratch, scale, slices, Kp, (int64_t)MN * Kp);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

// 1 = one CTA per 128 x 64 tile (oz_mma_kernel), 2 = CTA pair per 256 x 128 tile in two passes (oz2_mma_kernel)
static std::atomic<int> g_oz_variant{0};
static int oz_variant() {
  int v = g_oz_variant.load();
  if (v == 0) {
    const char* e = getenv("TNPY_OZAKI_VARIANT");
    v = (e && e[0] == '1') ? 1 : 2;
    g_oz_variant.store(v);
  }
  return v;
}

// C tile += scratch tiles of the K ranges ks = 1 .. splits-1, in that fixed order (deterministic).
__global__ void __launch_bounds__(256) oz2_tail_combine_kernel(GemmOut out, int M, int N, int tiles_m, int tiles_n, Oz2Tail tail) {
  const int idx = blockIdx.x;  // split tile
  const int tile = tail.n_full + idx;
  constexpr int GROUP = 4;
  const int per_group = GROUP * tiles_n;
  const int gid = tile / per_group, first_m = gid * GROUP;
  const int gsz = min(tiles_m - first_m, GROUP), rem = tile - gid * per_group;
  const int m0 = (first_m + rem % gsz) * 2 * kOzBM, n0 = (rem / gsz) * kOz2TileN;
  for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < 2 * kOzBM * kOz2TileN; e += gridDim.y * blockDim.x) {
    const int r = e / kOz2TileN, c = e % kOz2TileN;
    const int m = m0 + r, n = n0 + c;
    if (m >= M || n >= N) continue;
    double* dst = out.C + (int64_t)(m / out.m_inner) * out.c_outer + (int64_t)(m % out.m_inner) * out.c_inner + n;
    double acc = *dst;
    for (int ks = 1; ks < tail.splits; ++ks) acc += tail.scratch[(int64_t)((ks - 1) * tail.rem + idx) * (2 * kOzBM * kOz2TileN) + e];
    *dst = acc;
  }
}

// scratch for the tail tiles of oz2_gemm (grow-only, per process)
static int oz2_tail_scratch(size_t need, double** out) {
  static void* ptr = nullptr;
  static size_t bytes = 0;
  if (bytes < need) {
    if (ptr) {
      TNPY_CUDA_OK(cudaDeviceSynchronize());
      TNPY_CUDA_OK(cudaFree(ptr));
      ptr = nullptr;
      bytes = 0;
    }
    TNPY_CUDA_OK(cudaMalloc(&ptr, need));
    bytes = need;
  }
  *out = static_cast<double*>(ptr);
  return TNPY_OK;
}

template <int S>
static int oz2_gemm(const int8_t* As, const double* scaleA, const int8_t* Bs, const double* scaleB, GemmOut out, int M,
                    int N, int64_t Kp, int accumulate, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  TNPY_TRY(oz_make_map(&tmA, As, Kp, M, S, kOzBM, 4));
  TNPY_TRY(oz_make_map(&tmB, Bs, Kp, N, S, kOzBN, 4));
  constexpr int smem = kOz2Slots * kOz2Slot + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    TNPY_CUDA_OK(cudaFuncSetAttribute(oz2_mma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles_m = ceil_div(M, 2 * kOzBM), tiles_n = ceil_div(N, kOz2TileN), tiles = tiles_m * tiles_n;
  const int KT = (int)(Kp / kOzBK);
  // One CTA pair per SM pair at a time (shared memory): the last wave holds tiles % pairs tiles.  Cut those
  // along K so that the last wave is as wide as the machine; parts ks > 0 go to scratch tiles and are added
  // back in a fixed order.
  const int pairs = sm_count() / 2;
  Oz2Tail tail{tiles, 0, 1, nullptr};
  const int rem = tiles % pairs;
  if (rem > 0 && getenv("TNPY_OZAKI_NO_TAIL_SPLIT") == nullptr) {
    int splits = pairs / rem;
    while (splits > 1 && KT / splits < 8) --splits;
    if (splits > 4) splits = 4;
    const double full = (double)(tiles / pairs);
    // worth it only when the last wave is a sizeable part of the run (each part pays its own two epilogues)
    if (splits > 1 && (full + 1.0 / splits + 0.04) / (full + 1.0) < 0.93) {
      tail = Oz2Tail{tiles - rem, rem, splits, nullptr};
      TNPY_TRY(oz2_tail_scratch((size_t)(splits - 1) * rem * 2 * kOzBM * kOz2TileN * sizeof(double), &tail.scratch));
    }
  }
  const int items = tail.n_full + tail.rem * tail.splits;
  oz2_mma_kernel<S><<<2 * items, kOz2Threads, smem, stream>>>(tmA, tmB, scaleA, scaleB, out, M, N, KT, accumulate, tail);
  TNPY_LAUNCH_OK();
  if (tail.splits > 1) {
    oz2_tail_combine_kernel<<<dim3(tail.rem, 16), 256, 0, stream>>>(out, M, N, tiles_m, tiles_n, tail);
    TNPY_LAUNCH_OK();
  }
  return TNPY_OK;
}

template <int S>
static int oz_gemm(const int8_t* As, const double* scaleA, const int8_t* Bs, const double* scaleB, GemmOut out, int M,
                   int N, int64_t Kp, int accumulate, cudaStream_t stream) {
  if (oz_variant() == 2) return oz2_gemm<S>(As, scaleA, Bs, scaleB, out, M, N, Kp, accumulate, stream);
  CUtensorMap tmA, tmB;
  TNPY_TRY(oz_make_map(&tmA, As, Kp, M, S, kOzBM));
  TNPY_TRY(oz_make_map(&tmB, Bs, Kp, N, S, kOzBN));
  constexpr int smem = kOzStages * S * (kOzBM + kOzBN) * kOzBK + 1024 + 128;
  static bool configured = false;
  if (!configured) {
    TNPY_CUDA_OK(cudaFuncSetAttribute(oz_mma_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int grid = ceil_div(N, kOzBN) * ceil_div(M, kOzBM);
  oz_mma_kernel<S><<<grid, 256, smem, stream>>>(tmA, tmB, scaleA, scaleB, out, M, N, (int)(Kp / kOzBK), accumulate);
  TNPY_LAUNCH_OK();
  return TNPY_OK;
}

static int64_t oz_kp_(int K) { return ((int64_t)K + kOzBK - 1) / kOzBK * kOzBK; }

// Library-internal grow-only scratch for the slices (the chains call gemm_tn without a workspace for
// this purpose).  One buffer per process; calls are expected on one stream at a time.
struct OzScratch {
  void* ptr = nullptr;
  size_t bytes = 0;
};
static int oz_scratch(size_t need, void** out) {
  static OzScratch sc;
  if (sc.bytes < need) {
    if (sc.ptr) {
      TNPY_CUDA_OK(cudaDeviceSynchronize());
      TNPY_CUDA_OK(cudaFree(sc.ptr));
      sc.ptr = nullptr;
      sc.bytes = 0;
    }
    const size_t want = need + need / 8;
    TNPY_CUDA_OK(cudaMalloc(&sc.ptr, want));
    sc.bytes = want;
  }
  *out = sc.ptr;
  return TNPY_OK;
}

static std::atomic<int> g_oz_slices{8};
static std::atomic<int> g_oz_scope_slices{0};  // override inside an eigensolve whose tolerance allows fewer slices
int ozaki_slices() {
  const int o = g_oz_scope_slices.load();
  return o ? o : g_oz_slices.load();
}
// 0 clears the override.  7 slices: error ~2e-14 |A|^T|B| per GEMM, far below a residual tolerance >= 1e-10 ||A||.
void ozaki_scope_slices(int slices) { g_oz_scope_slices.store(slices); }

bool ozaki_applicable(int M, int N, int K) {
  // below ~chi = 1024 the slicing passes and extra launches cost more than the faster MMA saves (measured at chi = 512)
  return K <= 65536 && K >= 64 && M >= 128 && N >= 64 && (double)M * N * K >= 6.0e9;
}

// Slices of constant B operands (the environments L and R during one local eigensolve) are kept across calls:
// between ozaki_const_scope(true) and ozaki_const_scope(false) the caller vouches that the B operand behind a
// given (pointer, ld, K, N) does not change, so it is sliced once per scope instead of once per matvec.
struct OzConstEntry {
  const double* ptr = nullptr;
  int64_t ld = 0;
  int K = 0, N = 0, S = 0;
  bool valid = false;
  void* buf = nullptr;
  size_t bytes = 0;
};
static OzConstEntry g_oz_const[4];
static std::atomic<int> g_oz_const_depth{0};
static int g_oz_const_next = 0;

void ozaki_const_scope(bool on) {
  if (on) {
    if (g_oz_const_depth.fetch_add(1) == 0)
      for (auto& e : g_oz_const) e.valid = false;
  } else {
    if (g_oz_const_depth.fetch_sub(1) == 1)
      for (auto& e : g_oz_const) e.valid = false;
  }
}

// Returns the cached (or freshly filled) slices of B; *fresh tells the caller to run the slicing kernels.
static int oz_const_lookup(const double* B, int64_t ldb, int K, int N, int S, int64_t Kp, int8_t** slices, double** scale,
                           unsigned long long** colmax, bool* fresh) {
  for (auto& e : g_oz_const)
    if (e.valid && e.ptr == B && e.ld == ldb && e.K == K && e.N == N && e.S == S) {
      Workspace ws(e.buf, e.bytes);
      *slices = ws.take<int8_t>((size_t)S * N * Kp);
      *scale = ws.take<double>(N);
      *colmax = ws.take<unsigned long long>(N);
      *fresh = false;
      return TNPY_OK;
    }
  OzConstEntry& e = g_oz_const[g_oz_const_next];
  g_oz_const_next = (g_oz_const_next + 1) % 4;
  const size_t need = Workspace::need((size_t)S * N * Kp, 1) + 2 * Workspace::need(N) + 512;
  if (e.bytes < need) {
    if (e.buf) {
      TNPY_CUDA_OK(cudaDeviceSynchronize());
      TNPY_CUDA_OK(cudaFree(e.buf));
      e.buf = nullptr;
      e.bytes = 0;
    }
    TNPY_CUDA_OK(cudaMalloc(&e.buf, need));
    e.bytes = need;
  }
  e.ptr = B; e.ld = ldb; e.K = K; e.N = N; e.S = S; e.valid = true;
  Workspace ws(e.buf, e.bytes);
  *slices = ws.take<int8_t>((size_t)S * N * Kp);
  *scale = ws.take<double>(N);
  *colmax = ws.take<unsigned long long>(N);
  *fresh = true;
  return TNPY_OK;
}

// C (+)= A^T B through the int8 tensor cores; operands are sliced into the internal scratch.
int ozaki_gemm(const double* A, int64_t lda, const double* B, int64_t ldb, GemmOut out, int M, int N, int K,
               int accumulate, cudaStream_t stream) {
  const int S = ozaki_slices();
  const int64_t Kp = oz_kp_(K);
  const size_t need = Workspace::need((size_t)S * M * Kp, 1) + Workspace::need((size_t)S * N * Kp, 1) +
                      2 * Workspace::need(M) + 2 * Workspace::need(N) + 1024;
  void* base = nullptr;
  TNPY_TRY(oz_scratch(need, &base));
  Workspace ws(base, need);
  int8_t* As = ws.take<int8_t>((size_t)S * M * Kp);
  int8_t* Bs = ws.take<int8_t>((size_t)S * N * Kp);
  double* sa = ws.take<double>(M);
  double* sb = ws.take<double>(N);
  unsigned long long* ma = ws.take<unsigned long long>(M);
  unsigned long long* mb = ws.take<unsigned long long>(N);
  if (!As || !Bs || !sa || !sb || !ma || !mb) {
    set_error("ozaki_gemm: internal scratch layout failed");
    return TNPY_EWORKSPACE;
  }
  bool slice_b = true;
  if (g_oz_const_depth.load() > 0) TNPY_TRY(oz_const_lookup(B, ldb, K, N, S, Kp, &Bs, &sb, &mb, &slice_b));
  switch (S) {
    case 6:
      TNPY_TRY(oz_slice<6>(A, lda, K, M, sa, ma, As, Kp, stream));
      if (slice_b) TNPY_TRY(oz_slice<6>(B, ldb, K, N, sb, mb, Bs, Kp, stream));
      return oz_gemm<6>(As, sa, Bs, sb, out, M, N, Kp, accumulate, stream);
    case 7:
      TNPY_TRY(oz_slice<7>(A, lda, K, M, sa, ma, As, Kp, stream));
      if (slice_b) TNPY_TRY(oz_slice<7>(B, ldb, K, N, sb, mb, Bs, Kp, stream));
      return oz_gemm<7>(As, sa, Bs, sb, out, M, N, Kp, accumulate, stream);
    default:
      TNPY_TRY(oz_slice<8>(A, lda, K, M, sa, ma, As, Kp, stream));
      if (slice_b) TNPY_TRY(oz_slice<8>(B, ldb, K, N, sb, mb, Bs, Kp, stream));
      return oz_gemm<8>(As, sa, Bs, sb, out, M, N, Kp, accumulate, stream);
  }
}

}  // namespace tnpy

using namespace tnpy;

extern "C" int tnpy_ozaki_const_scope(int on) {
  ozaki_const_scope(on != 0);
  return TNPY_OK;
}

extern "C" int tnpy_set_ozaki_variant(int variant) {
  if (variant != 1 && variant != 2) {
    set_error("tnpy_set_ozaki_variant: variant must be 1 (single CTA, 128x64) or 2 (CTA pair, 256x128, two passes)");
    return TNPY_EINVAL;
  }
  g_oz_variant.store(variant);
  return TNPY_OK;
}

extern "C" int tnpy_set_ozaki_slices(int slices) {
  if (slices < 6 || slices > kOzMaxSlices) {
    set_error("tnpy_set_ozaki_slices: slices must be 6, 7 or 8");
    return TNPY_EINVAL;
  }
  g_oz_slices.store(slices);
  return TNPY_OK;
}

static int64_t oz_kp(int K) { return ((int64_t)K + kOzBK - 1) / kOzBK * kOzBK; }

extern "C" size_t tnpy_ozaki_workspace_bytes(int M, int N, int K, int slices) {
  const int64_t Kp = oz_kp(K);
  return Workspace::need((size_t)slices * M * Kp, 1) + Workspace::need((size_t)slices * N * Kp, 1) +
         2 * Workspace::need(M) + 2 * Workspace::need(N) + 1024;
}

// C[m,n] (+)= sum_k A[k,m] B[k,n] in FP64 accuracy on the int8 tensor cores (slices in 6..8).
// phase: 0 = slice both operands and multiply, 1 = slice only (fills the workspace), 2 = multiply only
// (workspace already holds the slices of these operands) -- lets a caller time / reuse the parts.
extern "C" int tnpy_ozaki_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                  int M, int N, int K, int slices, int accumulate, int phase, void* workspace,
                                  size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TNPY_CHECK_ARG(A && B && C && M > 0 && N > 0 && K > 0, "bad argument");
  TNPY_CHECK_ARG(slices >= 6 && slices <= kOzMaxSlices, "slices must be 6, 7 or 8");
  TNPY_CHECK_ARG(K <= 65536, "K too large for exact int32 accumulation");
  const int64_t Kp = oz_kp(K);
  Workspace ws(workspace, workspace_bytes);
  int8_t* As = ws.take<int8_t>((size_t)slices * M * Kp);
  int8_t* Bs = ws.take<int8_t>((size_t)slices * N * Kp);
  double* sa = ws.take<double>(M);
  double* sb = ws.take<double>(N);
  unsigned long long* ma = ws.take<unsigned long long>(M);
  unsigned long long* mb = ws.take<unsigned long long>(N);
  if (!As || !Bs || !sa || !sb || !ma || !mb) {
    set_error("tnpy_ozaki_gemm_tn: workspace too small");
    return TNPY_EWORKSPACE;
  }
  GemmOut out = plain_out(C, ldc, M);
  if (phase != 2) {
    switch (slices) {
      case 6: TNPY_TRY(oz_slice<6>(A, lda, K, M, sa, ma, As, Kp, stream)); TNPY_TRY(oz_slice<6>(B, ldb, K, N, sb, mb, Bs, Kp, stream)); break;
      case 7: TNPY_TRY(oz_slice<7>(A, lda, K, M, sa, ma, As, Kp, stream)); TNPY_TRY(oz_slice<7>(B, ldb, K, N, sb, mb, Bs, Kp, stream)); break;
      default: TNPY_TRY(oz_slice<8>(A, lda, K, M, sa, ma, As, Kp, stream)); TNPY_TRY(oz_slice<8>(B, ldb, K, N, sb, mb, Bs, Kp, stream)); break;
    }
  }
  if (phase != 1) {
    switch (slices) {
      case 6: return oz_gemm<6>(As, sa, Bs, sb, out, M, N, Kp, accumulate, stream);
      case 7: return oz_gemm<7>(As, sa, Bs, sb, out, M, N, Kp, accumulate, stream);
      default: return oz_gemm<8>(As, sa, Bs, sb, out, M, N, Kp, accumulate, stream);
    }
  }
  return TNPY_OK;
}
