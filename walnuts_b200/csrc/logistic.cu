// Batched log-density gradient of Bayesian logistic regression for ALL chains of
// a tick, on the 5th-generation tensor cores (tcgen05 + TMEM accumulators + TMA).
//
//   logp_c = sum_n [y_n z_nc - softplus(z_nc)] - 1/2 |theta_c|^2,   Z = X Theta^T
//   grad_c = X^T (y - sigmoid(z_c)) - theta_c
//
// (the density itself is not in the reference repo — the CPU
// statement the tests check it against is the `Logistic` target of the test oracle.)  Per tick:
//
//   pack      Theta fp64 [C][ld]  ->  bf16 hi / lo planes [Cpad][Dpad]
//   GEMM 1    Z^T[c][n] = (hi + lo)[c][:] . X[n][:]    (K = 2*Dpad, fp32 in TMEM; a chain
//             is a TMEM lane, so one thread owns one chain's row of the tile)
//             epilogue: sigmoid(z) -> R^T[c][n] bf16 ; sum_n softplus(z) -> SP[c]
//   GEMM 2    G[c][d] = R^T[c][:] . X^T[d][:]          (K = Npad, fp32 in TMEM)
//   finalize  grad = X^T y - G - theta ; logp = (X^T y).theta - SP - 1/2 |theta|^2 (fp64)
//
// Both GEMMs are one kernel: D[m][n] = sum_k A[m][k] B[n][k], A and B bf16 K-major,
// 128 x BN x 64 tiles, TMA (SWIZZLE_128B) -> shared -> tcgen05.mma.kind::f16 issued by
// one thread, accumulators double-buffered in TMEM, epilogue warps read them back with
// tcgen05.ld while the next tile's MMAs run (persistent CTAs, one per SM).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "engine.cuh"
#include "logistic.cuh"

namespace wb200 {

// ---------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
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
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map,
                                            uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global tile store through the TMA (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src,
                                             int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {  // staging buffer may be rewritten
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {  // generic writes -> async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait, and the wait as a separate step that the registers
// pass through (so no use of them can be scheduled above it): lets an epilogue warp
// fetch its next 32 columns while it works on the current ones
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                 "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                 "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]),
                 "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;             // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO: 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): F32 accum, BF16 x BF16, K-major
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

constexpr int BM = 128, BK = 64;
// warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare; then the epilogue warps
// (4 lane quarters x column halves when there are 8)
template <int EPI> constexpr int epi_warps() { return EPI == 1 ? 8 : 4; }
template <int EPI> constexpr int gemm_threads() { return 128 + 32 * epi_warps<EPI>(); }

struct GemmParams {
  int num_m_tiles, num_n_tiles;
  int num_k_blocks;        // K blocks (of each plane of A)
  // split-K (epilogue 2 only): a work item is (tile, split); split s accumulates K blocks
  // [s * k_blocks_per_split, ...) and ADDS its partial tile to `out` (zeroed beforehand).
  // Fills the SMs when there are fewer output tiles than SMs (G = R^T X has 64 tiles for a
  // half batch of 4096 chains, each with a 1564-block K loop).
  int k_splits, k_blocks_per_split;
  // epilogue 1 (logistic): rows = chains, columns = data rows
  int n_valid;             // data rows < n_valid are real
  __nv_bfloat16* RT;       // [Cpad][ldrt]  sigmoid(z), 0 for padding rows
  long long ldrt;
  double* SP;              // [Cpad] sum_n softplus
  // epilogue 2 (plain store)
  float* out;              // [M][ldo]
  long long ldo;
};

// A pipeline stage holds one K block of B and of EVERY plane of A: GEMM 1's hi and lo
// planes meet the same X tile, so it is fetched once per K block instead of once per plane
// (a third less L2 -> shared-memory traffic per tile: 512 instead of 768 KB).
template <int BN, int PLANES = 1>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? (PLANES == 2 ? 3 : 4) : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = PLANES * kABytes + kBBytes;
  static constexpr int kTileBytes = kStages * kStageBytes;
  static constexpr int kBarrierBytes = 1024;  // keeps the staging area 1024-aligned
  // GEMM 1 epilogue staging: 8 warps x [32 rows][128 B] in the SWIZZLE_128B layout
  static constexpr int kStagingPerWarp = 32 * 128;
  static constexpr int kStagingBytes = 8 * kStagingPerWarp;
  static constexpr int kTotal = kTileBytes + 1024 /*align*/ + kBarrierBytes + kStagingBytes;
};
static_assert(GemmSmem<256>::kTotal <= 227 * 1024, "shared memory budget");
static_assert(GemmSmem<256, 2>::kTotal <= 227 * 1024, "shared memory budget");
constexpr int gemm_planes(int epi) { return epi == 1 ? 2 : 1; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// GEMM 1 works in base-2 units: Theta is scaled by log2(e) when it is packed, so the
// accumulator holds w = z * log2(e).  With t = 2^-w, d = 1 + t:
//   sigmoid(z) = 1 / d ;  softplus(z) = ln2 * (w + log2(d)) =: ln2 * u
// three MUFU ops (ex2, rcp, lg2), one clamp and three adds per element.  The clamp
// keeps t finite for very negative z (w >= -115: t <= 2^115, u -> 0); for large
// positive z, t underflows to 0 and u = w, both correct to fp32 rounding.
__device__ __forceinline__ void sigmoid_softplus2(float w, float& sig, float& u) {
  w = fmaxf(w, -115.0f);
  const float d = 1.0f + ex2_approx(-w);
  sig = rcp_approx(d);
  u = w + lg2_approx(d);
}

// GEMM 1 epilogue for 32 columns of one chain: sigmoid -> bf16, base-2 softplus sum.
// Padded data rows (n >= N) are zero rows of X: w = 0, sigmoid = 1/2 (meets the zero
// columns of X^T in GEMM 2) and u = 1 exactly, which the finalize kernel subtracts --
// so there is no masking here.
// The 64 bytes go to the warp's staging tile ([32 rows][128 B], SWIZZLE_128B: 16-byte
// slot s of row r lives at slot s ^ (r & 7)), `half` selects columns 0-31 / 32-63; the
// tile leaves through a TMA store, so R^T is written in whole 128-byte lines.  (Direct
// per-thread stores, one row per lane, held the tensor pipe at 65-78 %: measured with
// the stores removed it runs at 94 %.)
__device__ __forceinline__ void epi1_chunk(const uint32_t (&v)[32], uint8_t* stage_row,
                                           int lane, int half, float& sp_acc) {
  uint32_t packed[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float s0, s1, u0, u1;
    sigmoid_softplus2(__uint_as_float(v[i]), s0, u0);
    sigmoid_softplus2(__uint_as_float(v[i + 1]), s1, u1);
    sp_acc += u0 + u1;
    __nv_bfloat162 b2 = __floats2bfloat162_rn(s0, s1);
    packed[i / 2] = *reinterpret_cast<uint32_t*>(&b2);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int slot = (4 * half + q) ^ (lane & 7);
    *reinterpret_cast<uint4*>(stage_row + 16 * slot) =
        make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
  }
}

// D[m][n] = sum_k A[m][k] B[n][k]; A may come in two planes that are summed (hi + lo).
// Persistent: each CTA walks tiles t = blockIdx.x, + gridDim.x, ... (m fastest), with a
// double-buffered TMEM accumulator so the epilogue of tile i overlaps the MMAs of i+1.
// EPI = 1: logistic residual epilogue; EPI = 2: store fp32 tile
template <int BN, int EPI>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap mapA0,
                   const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapOut, const GemmParams gp) {
  constexpr int PLANES = gemm_planes(EPI);
  using S = GemmSmem<BN, PLANES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kTileBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tmem_full = empty_bar + S::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* staging = smem + S::kTileBytes + S::kBarrierBytes;  // EPI 1 only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = gp.num_m_tiles * gp.num_n_tiles;
  const int num_work = num_tiles * gp.k_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(full_bar + s, 1);
      mbar_init(empty_bar + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full + b, 1);
      mbar_init(tmem_empty + b, epi_warps<EPI>());  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "n"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer
    if (lane == 0) {
      int it = 0;  // running k-block count across tiles (stage ring position)
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int t = w % num_tiles, ks = w / num_tiles;
        const int m0 = (t % gp.num_m_tiles) * BM;
        const int n0 = (t / gp.num_m_tiles) * BN;
        const int kb0 = ks * gp.k_blocks_per_split;
        const int kb1 = min(gp.num_k_blocks, kb0 + gp.k_blocks_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % S::kStages;
          const uint32_t ph = (it / S::kStages) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          uint8_t* a_dst = tiles + s * S::kStageBytes;
          uint8_t* b_dst = a_dst + PLANES * S::kABytes;
          mbar_expect_tx(full_bar + s, S::kStageBytes);
          const int kk = kb * BK;
          tma_load_2d(a_dst, &mapA0, full_bar + s, kk, m0);
          if constexpr (PLANES == 2) {
            tma_load_2d(a_dst + S::kABytes, &mapA1, full_bar + s, kk, m0);
          }
          tma_load_2d(b_dst, &mapB, full_bar + s, kk, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int it = 0, as = 0;
      uint32_t aph = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int ks = w / num_tiles;
        const int kb0 = ks * gp.k_blocks_per_split;
        const int kb1 = min(gp.num_k_blocks, kb0 + gp.k_blocks_per_split);
        mbar_wait(tmem_empty + as, aph ^ 1);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % S::kStages;
          const uint32_t ph = (it / S::kStages) & 1;
          mbar_wait(full_bar + s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(tiles + s * S::kStageBytes);
          const uint32_t b_addr = a_addr + PLANES * S::kABytes;
          const uint64_t bdesc = make_kmajor_sw128_desc(b_addr);
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            const uint64_t adesc = make_kmajor_sw128_desc(a_addr + pl * S::kABytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128 B swizzle span: +2 in (addr >> 4)
              umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc,
                        kb != kb0 || pl != 0 || k != 0);
            }
          }
          umma_commit(empty_bar + s);  // frees the stage when these MMAs retire
        }
        umma_commit(tmem_full + as);  // accumulator complete
        as ^= 1;
        if (as == 0) aph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: TMEM -> registers -> global
    const int wq = (warp - 4) & 3;   // TMEM lane quarter this warp may read
    const int part = (warp - 4) >> 2;  // which slice of the columns it handles
    constexpr int kChunks = BN / 32 / (epi_warps<EPI>() / 4);
    int as = 0;
    uint32_t aph = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int t = w % num_tiles;
      const int m0 = (t % gp.num_m_tiles) * BM;
      const int n0 = (t / gp.num_m_tiles) * BN;
      mbar_wait(tmem_full + as, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + wq * 32 + lane;  // A row of this thread (its TMEM lane)
      const uint32_t tmem_acc = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + as * BN;
      if constexpr (EPI == 1) {
        // row = chain, columns = data rows n0 + ...; R^T holds sigmoid(z) (the y term of
        // the gradient is the constant X^T y, added by the finalize kernel)
        float sp_acc = 0.0f;
        static_assert(kChunks % 2 == 0, "the epilogue is software-pipelined in pairs");
        const int j0 = part * kChunks;
        uint8_t* stage = staging + (warp - 4) * S::kStagingPerWarp;
        uint8_t* stage_row = stage + lane * 128;
        uint32_t va[32], vb[32];
        tmem_ld32_issue(tmem_acc + j0 * 32, va);
#pragma unroll
        for (int jj = 0; jj < kChunks; jj += 2) {
          tmem_ld_wait(va);
          tmem_ld32_issue(tmem_acc + (j0 + jj + 1) * 32, vb);
          // the previous TMA store must have read the staging tile before it is rewritten
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          epi1_chunk(va, stage_row, lane, 0, sp_acc);
          tmem_ld_wait(vb);
          if (jj + 2 < kChunks) tmem_ld32_issue(tmem_acc + (j0 + jj + 2) * 32, va);
          epi1_chunk(vb, stage_row, lane, 1, sp_acc);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&mapOut, stage, n0 + (j0 + jj) * 32, m0 + wq * 32);
          }
        }
        atomicAdd(gp.SP + row, static_cast<double>(sp_acc));
      } else {
#pragma unroll 1
        for (int jj = 0; jj < kChunks; ++jj) {
          const int j = part * kChunks + jj;
          uint32_t v[32];
          tmem_ld32(tmem_acc + j * 32, v);
          float* dst_f = gp.out + static_cast<long long>(row) * gp.ldo + n0 + j * 32;
          if (gp.k_splits > 1) {
#pragma unroll
            for (int q = 0; q < 32; ++q) atomicAdd(dst_f + q, __uint_as_float(v[q]));
          } else {
            float4* dst = reinterpret_cast<float4*>(dst_f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                   __uint_as_float(v[4 * q + 2]),
                                   __uint_as_float(v[4 * q + 3]));
            }
          }
        }
      }
      // hand the accumulator back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
      as ^= 1;
      if (as == 0) aph ^= 1;
    }
    if constexpr (EPI == 1) {
      if (lane == 0) tma_store_wait_all();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(2 * BN));
  }
}

// ---------------------------------------------------------------------------
// Theta fp64 -> bf16 hi / lo planes
__global__ void pack_theta_kernel(const double* TH, int ld, int C, int D, int Dpad,
                                  __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(C) * Dpad) return;
  const int c = static_cast<int>(i / Dpad), d = static_cast<int>(i % Dpad);
  // scaled by log2(e): GEMM 1 accumulates z in base-2 units (see sigmoid_softplus2)
  const double t = d < D ? 1.4426950408889634 * TH[static_cast<long long>(c) * ld + d] : 0.0;
  const __nv_bfloat16 h = __double2bfloat16(t);
  const double rem = t - static_cast<double>(__bfloat162float(h));
  hi[i] = h;
  lo[i] = __double2bfloat16(rem);
}

// grad = b - G32 - theta (G32 = sum_n sigmoid x) ; logp = b.theta - SP - 1/2 |theta|^2
__global__ void logistic_finalize_kernel(const double* TH, int ld, int C, int D,
                                         const float* G32, long long ldg, const double* b,
                                         const double* SP, double n_pad, double* G,
                                         double* LP) {
  const int c = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const double* th = TH + static_cast<long long>(c) * ld;
  double bt = 0.0, ss = 0.0;
  for (int d = lane; d < ld; d += 32) {
    if (d < D) {
      const double t = th[d];
      G[static_cast<long long>(c) * ld + d] =
          (b[d] - static_cast<double>(G32[static_cast<long long>(c) * ldg + d])) - t;
      bt += b[d] * t;
      ss += t * t;
    } else {
      G[static_cast<long long>(c) * ld + d] = 0.0;
    }
  }
  for (int m = 16; m >= 1; m >>= 1) {
    bt += __shfl_xor_sync(0xffffffffu, bt, m);
    ss += __shfl_xor_sync(0xffffffffu, ss, m);
  }
  // SP = sum over the Npad data rows of u (base-2 softplus); the n_pad zero rows add 1 each
  if (lane == 0) LP[c] = bt - 0.6931471805599453 * (SP[c] - n_pad) - 0.5 * ss;
}

// ---------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    WB200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) {
      throw std::runtime_error("cuTensorMapEncodeTiled is not available in this driver");
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 matrix [rows][cols] row-major (cols contiguous), box = 64 cols x box_rows
static CUtensorMap make_map(const void* base, long long rows, long long cols, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  }
  return m;
}

static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

// static data of the model, shared by every batch evaluated against it
struct LogisticData {
  int N, D;
  long long Npad, Dpad;
  int bn2;
  DeviceBuffer<__nv_bfloat16> X, XT;
  DeviceBuffer<double> b;
  CUtensorMap mapX, mapXT;
};

struct LogisticGrad::Impl {
  std::shared_ptr<LogisticData> data;
  int C, ld;
  long long Cpad;
  DeviceBuffer<__nv_bfloat16> hi, lo, RT;
  DeviceBuffer<float> G32;
  DeviceBuffer<double> SP;
  CUtensorMap mapHi, mapLo, mapRT, mapRTst;
  int device = 0, sms = 148;
};

static void logistic_batch_buffers(LogisticGrad::Impl& m, int C, int ld, cudaStream_t stream);

LogisticGrad::LogisticGrad(const double* Xh, const double* yh, size_t N, int D, int C, int ld,
                           cudaStream_t stream)
    : impl_(new Impl) {
  Impl& m = *impl_;
  m.data = std::make_shared<LogisticData>();
  LogisticData& d = *m.data;
  d.N = static_cast<int>(N); d.D = D;
  d.Npad = round_up(static_cast<long long>(N), 256);
  d.Dpad = round_up(D, 64);
  d.bn2 = (d.Dpad % 256 == 0) ? 256 : (d.Dpad % 128 == 0 ? 128 : 64);
  // host staging: X and X^T in bf16 (the benchmark's X is bf16-representable, so this
  // is exact; otherwise it rounds to nearest), b = X^T y in fp64
  std::vector<__nv_bfloat16> xb(static_cast<size_t>(d.Npad) * d.Dpad, __float2bfloat16(0.0f));
  std::vector<__nv_bfloat16> xt(static_cast<size_t>(d.Dpad) * d.Npad, __float2bfloat16(0.0f));
  std::vector<double> bh(d.Dpad, 0.0);
  for (size_t n = 0; n < N; ++n) {
    for (int j = 0; j < D; ++j) {
      const __nv_bfloat16 v = __double2bfloat16(Xh[n * D + j]);
      xb[n * d.Dpad + j] = v;
      xt[static_cast<size_t>(j) * d.Npad + n] = v;
      bh[j] += static_cast<double>(__bfloat162float(v)) * yh[n];
    }
  }
  d.X.alloc(xb.size()); d.XT.alloc(xt.size()); d.b.alloc(bh.size());
  WB200_CUDA(cudaMemcpyAsync(d.X.ptr, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaMemcpyAsync(d.XT.ptr, xt.data(), xt.size() * 2, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaMemcpyAsync(d.b.ptr, bh.data(), bh.size() * 8, cudaMemcpyHostToDevice, stream));
  WB200_CUDA(cudaStreamSynchronize(stream));
  d.mapX = make_map(d.X.ptr, d.Npad, d.Dpad, 256);         // GEMM 1 B (data rows)
  d.mapXT = make_map(d.XT.ptr, d.Dpad, d.Npad, d.bn2);     // GEMM 2 B
  WB200_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<256, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  GemmSmem<256, 2>::kTotal));
  WB200_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<64, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  GemmSmem<64>::kTotal));
  WB200_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<128, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  GemmSmem<128>::kTotal));
  WB200_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<256, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  GemmSmem<256>::kTotal));
  logistic_batch_buffers(m, C, ld, stream);
}

LogisticGrad::LogisticGrad(const LogisticGrad& data_of, int C, cudaStream_t stream)
    : impl_(new Impl) {
  impl_->data = data_of.impl_->data;
  logistic_batch_buffers(*impl_, C, data_of.impl_->ld, stream);
}

static void logistic_batch_buffers(LogisticGrad::Impl& m, int C, int ld, cudaStream_t stream) {
  const LogisticData& d = *m.data;
  m.C = C; m.ld = ld;
  m.Cpad = round_up(C, 128);
  WB200_CUDA(cudaGetDevice(&m.device));
  WB200_CUDA(cudaDeviceGetAttribute(&m.sms, cudaDevAttrMultiProcessorCount, m.device));
  m.hi.alloc(static_cast<size_t>(m.Cpad) * d.Dpad);
  m.lo.alloc(static_cast<size_t>(m.Cpad) * d.Dpad);
  m.RT.alloc(static_cast<size_t>(m.Cpad) * d.Npad);
  m.G32.alloc(static_cast<size_t>(m.Cpad) * d.Dpad);
  m.SP.alloc(m.Cpad);
  WB200_CUDA(cudaMemsetAsync(m.hi.ptr, 0, m.hi.count * 2, stream));
  WB200_CUDA(cudaMemsetAsync(m.lo.ptr, 0, m.lo.count * 2, stream));
  WB200_CUDA(cudaStreamSynchronize(stream));
  m.mapHi = make_map(m.hi.ptr, m.Cpad, d.Dpad, BM);        // GEMM 1 A planes (chains)
  m.mapLo = make_map(m.lo.ptr, m.Cpad, d.Dpad, BM);
  m.mapRT = make_map(m.RT.ptr, m.Cpad, d.Npad, BM);        // GEMM 2 A
  m.mapRTst = make_map(m.RT.ptr, m.Cpad, d.Npad, 32);      // GEMM 1 epilogue stores
}

LogisticGrad::~LogisticGrad() { delete impl_; }

int LogisticGrad::kernels_per_eval() const { return 5; }

double LogisticGrad::flops_per_eval() const {
  const Impl& m = *impl_;
  const LogisticData& d = *m.data;
  // what the tensor cores execute: hi + lo passes of GEMM 1, one pass of GEMM 2
  return 2.0 * d.Npad * m.Cpad * (2.0 * d.Dpad) + 2.0 * m.Cpad * d.Dpad * d.Npad;
}

void LogisticGrad::evaluate(const double* TH, double* G, double* LP, cudaStream_t stream) {
  Impl& m = *impl_;
  const LogisticData& d = *m.data;
  const long long n = static_cast<long long>(m.C) * d.Dpad;
  pack_theta_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      TH, m.ld, m.C, d.D, static_cast<int>(d.Dpad), m.hi.ptr, m.lo.ptr);
  WB200_CUDA(cudaMemsetAsync(m.SP.ptr, 0, m.Cpad * sizeof(double), stream));
  GemmParams g1{};
  g1.num_m_tiles = static_cast<int>(m.Cpad / BM);
  g1.num_n_tiles = static_cast<int>(d.Npad / 256);
  g1.num_k_blocks = static_cast<int>(d.Dpad / BK);
  g1.k_splits = 1; g1.k_blocks_per_split = g1.num_k_blocks;
  g1.n_valid = d.N; g1.RT = m.RT.ptr; g1.ldrt = d.Npad; g1.SP = m.SP.ptr;
  const int grid1 = std::min(m.sms, g1.num_m_tiles * g1.num_n_tiles);
  gemm_kmajor_kernel<256, 1><<<grid1, gemm_threads<1>(), GemmSmem<256, 2>::kTotal, stream>>>(
      m.mapHi, m.mapLo, d.mapX, m.mapRTst, g1);
  WB200_CUDA(cudaGetLastError());
  GemmParams g2{};
  g2.num_m_tiles = static_cast<int>(m.Cpad / BM);
  g2.num_n_tiles = static_cast<int>(d.Dpad / d.bn2);
  g2.num_k_blocks = static_cast<int>(d.Npad / BK);
  // fewer output tiles than half the SMs: split the K loop so that every SM has work
  const int tiles2 = g2.num_m_tiles * g2.num_n_tiles;
  g2.k_splits = std::max(1, std::min({8, m.sms / std::max(tiles2, 1), g2.num_k_blocks / 64}));
  g2.k_blocks_per_split = (g2.num_k_blocks + g2.k_splits - 1) / g2.k_splits;
  // every split must own at least one K block (its accumulator is read back and added)
  g2.k_splits = (g2.num_k_blocks + g2.k_blocks_per_split - 1) / g2.k_blocks_per_split;
  g2.out = m.G32.ptr; g2.ldo = d.Dpad;
  if (g2.k_splits > 1) {
    WB200_CUDA(cudaMemsetAsync(m.G32.ptr, 0, m.G32.count * sizeof(float), stream));
  }
  const int grid2 = std::min(m.sms, tiles2 * g2.k_splits);
  if (d.bn2 == 256) {
    gemm_kmajor_kernel<256, 2><<<grid2, gemm_threads<2>(), GemmSmem<256>::kTotal, stream>>>(
        m.mapRT, m.mapRT, d.mapXT, m.mapRT, g2);
  } else if (d.bn2 == 128) {
    gemm_kmajor_kernel<128, 2><<<grid2, gemm_threads<2>(), GemmSmem<128>::kTotal, stream>>>(
        m.mapRT, m.mapRT, d.mapXT, m.mapRT, g2);
  } else {
    gemm_kmajor_kernel<64, 2><<<grid2, gemm_threads<2>(), GemmSmem<64>::kTotal, stream>>>(
        m.mapRT, m.mapRT, d.mapXT, m.mapRT, g2);
  }
  WB200_CUDA(cudaGetLastError());
  logistic_finalize_kernel<<<(m.C + 7) / 8, 256, 0, stream>>>(
      TH, m.ld, m.C, d.D, m.G32.ptr, d.Dpad, d.b.ptr, m.SP.ptr,
      static_cast<double>(d.Npad - d.N), G, LP);
  WB200_CUDA(cudaGetLastError());
}

}  // namespace wb200

// ---------------------------------------------------------------------------
// Stand-alone operator for parity tests and tensor-pipe measurements
extern "C" int wb200_logistic_logp_grad(const double* X, const double* y, size_t N, int D,
                                        const double* theta, size_t C, double* logp,
                                        double* grad, int repeats, float* ms_per_eval,
                                        WalnutpyError** err) {
  using namespace wb200;
  return catch_exceptions(err, [&] {
    require_gpu();
    if (D < 1 || N < 1 || C < 1) throw std::invalid_argument("empty problem");
    const int ld = (D + 1) & ~1;
    cudaStream_t st = nullptr;
    LogisticGrad lg(X, y, N, D, static_cast<int>(C), ld, st);
    DeviceBuffer<double> TH, G, LP;
    TH.alloc(C * ld); G.alloc(C * ld); LP.alloc(C);
    WB200_CUDA(cudaMemset(TH.ptr, 0, C * ld * 8));
    upload_rows(TH.ptr, ld, theta, D, C, st);
    lg.evaluate(TH.ptr, G.ptr, LP.ptr, st);
    WB200_CUDA(cudaDeviceSynchronize());
    if (repeats > 0 && ms_per_eval) {
      cudaEvent_t e0, e1;
      WB200_CUDA(cudaEventCreate(&e0));
      WB200_CUDA(cudaEventCreate(&e1));
      WB200_CUDA(cudaEventRecord(e0, st));
      for (int i = 0; i < repeats; ++i) lg.evaluate(TH.ptr, G.ptr, LP.ptr, st);
      WB200_CUDA(cudaEventRecord(e1, st));
      WB200_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      WB200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      *ms_per_eval = ms / repeats;
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
    }
    download_rows(grad, D, G.ptr, ld, C, st);
    WB200_CUDA(cudaMemcpy(logp, LP.ptr, C * 8, cudaMemcpyDeviceToHost));
    WB200_CUDA(cudaDeviceSynchronize());
  });
}
