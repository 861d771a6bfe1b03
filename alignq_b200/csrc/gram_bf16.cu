// bf16 Gram / SYRK on the 5th-gen tensor cores: G = X X^T (/ F) for X [B <= 256, F] bf16 row-major.
//
// This is the dense contraction at the heart of corr() (cdf_alignment_admm/resnet-56-cifar-10/model/
// quantization.py:134-137: torch.matmul(x_std, x_std^T) / F) for operands that are ALREADY
// standardised and stored in bf16 -- the tensor-bound micro-shape of SURVEY.md 8(d)
// (B = 256, F = 2^20: 137 GFLOP over 537 MB, arithmetic intensity 256 flop/B = the bf16 ridge).
// The fp32 training path (gram_tc.cu) has to transform its operands first and is ALU/HBM-bound.
//
// Three pipelines feed the same tcgen05 MMAs (K-major SWIZZLE_128B descriptors, fp32 accumulators in TMEM, split-K
// over the feature dimension, one fp32 partial per CTA, deterministic reduce):
//   gram_bf16_ldgsts_kernel  DEFAULT.  cp.async (LDGSTS) producers, 256 contiguous bytes per row and request, the
//                            128-byte swizzle applied by hand.  B = 256, F = 2^20: 108 us = 1.28 PFLOP/s;
//                            F = 2^22: 1.42 PFLOP/s, 5.6 TB/s of HBM reads.
//   gram_bf16_kernel         TMA producer (cp.async.bulk.tensor boxes of 256 rows x 128 B, 6-stage mbarrier ring).
//                            Same MMAs; tops out at 4.3 TB/s on this shape (124 us) -- kept for A/B runs
//                            (ALIGNQ_GRAM16_PATH=tma).
//   gram_bf16_pair_kernel    two SMs per tile (tcgen05.mma.cta_group::2, M = 256), remote mbarrier arrives and a
//                            multicast tcgen05.commit.  Bit-correct, but 197 us: not the default
//                            (ALIGNQ_GRAM16_PATH=pair).
// What the experiments showed: the load pipeline alone streams at 6.7 TB/s and the M = 128 MMAs alone run
// at their floor; what decides the overall rate is how early a landed stage is published to the MMA warp (publishing
// a stage only when the ring is full serialises producer and consumer), and the size of the memory requests.
//
// Structure of the TMA variant (one CTA per SM, 192 threads):
//   warp 0   TMA producer: one cp.async.bulk.tensor box [256 rows x 64 cols] (32 KB, SWIZZLE_128B) per
//            stage into a 6-stage ring, mbarrier expect_tx / complete_tx
//   warp 1   MMA issuer: per 16-column k-step two tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = 256):
//            rows [0,128) x all 256 rows (N = 256) and rows [128,256) x rows [128,256) (N = 128; the missing
//            block is the transpose of one already computed) -- A and B descriptors point into the SAME stage
//            (K-major, 128B-swizzled); accumulators in TMEM (256 + 128 fp32 columns);
//            tcgen05.commit releases the stage to the producer
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> smem transpose -> coalesced fp32 partial [256 x 256]
// followed by gram_bf16_reduce_kernel (sum of the per-CTA partials, deterministic).
#include <cuda.h>
#include <cstdlib>
#include <cstdio>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace g16 {

using namespace tc;

constexpr int ROWS = 256;                        // padded batch (TMA zero-fills rows >= B)
constexpr int BK = 64;                           // bf16 columns per stage = 128 bytes = one swizzle atom
constexpr int STAGES = 6;
constexpr int STAGE_BYTES = ROWS * BK * 2;       // 32 KB
constexpr int NTHREADS = 192;
constexpr int EPI_FLOATS = 4 * 32 * 33;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_FLOATS * 4 + 1024 /*align slack*/ + 256;

// K-major SWIZZLE_128B descriptor: 8-row x 128-byte atoms, SBO = 1024 B between 8-row groups, LBO unused (1)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(NTHREADS, 1)
gram_bf16_kernel(const __grid_constant__ CUtensorMap tmap, int B, int64_t ktiles, float* __restrict__ partials) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + EPI_FLOATS);
  uint64_t* empty = full + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's k-tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t my_tiles = (ktiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    // ---------------- TMA producer ----------------
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % STAGES);
      if (i >= STAGES) mbar_wait(&empty[s], (uint32_t)((i / STAGES - 1) & 1));
      mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
      const int64_t kt = blockIdx.x + i * gridDim.x;
      tma_load_2d(smem + s * STAGE_BYTES, &tmap, (int)(kt * BK), 0, &full[s]);
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t IDESC = make_idesc(1u /*bf16*/, 128u, 256u);
    constexpr uint32_t IDESC_SYM = make_idesc(1u, 128u, 128u);   // lower row half: only its diagonal block (G is symmetric)
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int s = (int)(i % STAGES);
      mbar_wait(&full[s], (uint32_t)((i / STAGES) & 1));
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
      for (int ks = 0; ks < BK / 16; ++ks) {
        const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
        const uint64_t db = desc_sw128(sb + ks * 32);                       // all 256 rows: N = 256
        const uint64_t da0 = db;                                            // rows   0..127
        const uint64_t da1 = desc_sw128(sb + 128 * 128 + ks * 32);          // rows 128..255
        umma<false>(tmem_base, da0, db, IDESC, acc);                        // G[0:128, 0:256]
        umma<false>(tmem_base + 256, da1, da1, IDESC_SYM, acc);             // G[128:256, 128:256]; the rest is G[0:128, 128:256]^T
      }
      umma_commit(&empty[s]);
    }
    umma_commit(done);
  }
  // ---------------- epilogue (warps 2..5) ----------------
  if (warp >= 2) {
    mbar_wait(done, 0);
    tc_fence_after();
    const int qd = warp & 3;                       // TMEM lane quarter this warp may access (warp id % 4)
    float* tr = epi + (warp - 2) * (32 * 33);
    float* out = partials + (size_t)blockIdx.x * B * B;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int row0 = half * 128 + qd * 32;
      if (row0 >= B) continue;
#pragma unroll 1
      for (int col0 = half * 128; col0 < B; col0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + half * 256 + (col0 - half * 128), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        if (col0 + lane < B) {
          for (int r = 0; r < 32; ++r) {
            if (row0 + r >= B) break;
            out[(size_t)(row0 + r) * B + col0 + lane] = tr[r * 33 + lane];
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// Variant fed by cp.async (LDGSTS) instead of TMA.  Measured on B200: TMA boxes of 256 rows x 128 bytes top
// out at 4.3 TB/s chip-wide on this shape (one L2 request per 128-byte row), while plain 16-byte-per-lane
// loads of the same matrix (e.g. torch.sum(x, 0)) run at 6.5 TB/s.  Here the four epilogue warps, idle during
// the main loop, are the producers: one warp instruction fetches 2 rows x 256 contiguous bytes (two k-tiles),
// and the destination address applies the 128-byte swizzle by hand (16-byte chunk index XOR row % 8), so the
// MMA side (same SWIZZLE_128B K-major descriptors) is unchanged.
//   warp 0      MMA issuer (+ TMEM alloc)
//   warps 1-4   producers: super-stage = 2 k-tiles (64 KB), 3 super-stages; cp.async groups, then
//               fence.proxy.async + mbarrier.arrive per super-stage; afterwards the epilogue
constexpr int LG_THREADS = 160;
constexpr int LG_SS_TILES = 2;                                  // k-tiles per super-stage
constexpr int LG_SS_BYTES = LG_SS_TILES * STAGE_BYTES;          // 64 KB
constexpr int LG_NSS = 3;
constexpr int LG_LA = 1;                                        // a super-stage is published LG_LA iterations after its loads were
                                                                // issued; the ring's remaining NSS - LA slots are slack for the MMA side
constexpr int LG_PRODUCERS = 128;
constexpr int LG_SMEM_BYTES = LG_NSS * LG_SS_BYTES + EPI_FLOATS * 4 + 1024 + 256;

__device__ __forceinline__ void lg_cp_async16(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lg_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lg_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void lg_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(LG_THREADS, 1)
gram_bf16_ldgsts_kernel(const __nv_bfloat16* __restrict__ x, int B, int64_t F, int64_t nsuper, float* __restrict__ partials) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + LG_NSS * LG_SS_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + EPI_FLOATS);
  uint64_t* empty = full + LG_NSS;
  uint64_t* done = empty + LG_NSS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < LG_NSS; ++s) { mbar_init(&full[s], LG_PRODUCERS); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t my_super = (nsuper - blockIdx.x + gridDim.x - 1) / gridDim.x;     // super-tiles b, b + grid, ...

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t IDESC_N256 = make_idesc(1u /*bf16*/, 128u, 256u);
      constexpr uint32_t IDESC_SYM = make_idesc(1u, 128u, 128u);
      const bool two_halves = B > 128;                               // B <= 128: one M = 128, N = 128 product is all there is
      const uint32_t IDESC = two_halves ? IDESC_N256 : IDESC_SYM;
      for (int64_t i = 0; i < my_super; ++i) {
        const int s = (int)(i % LG_NSS);
        mbar_wait(&full[s], (uint32_t)((i / LG_NSS) & 1));
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < LG_SS_TILES; ++j) {
          const uint32_t sb = smem_u32(smem + s * LG_SS_BYTES + j * STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint32_t acc = (i > 0 || j > 0 || ks > 0) ? 1u : 0u;
            const uint64_t db = desc_sw128(sb + ks * 32);
            const uint64_t da1 = desc_sw128(sb + 128 * 128 + ks * 32);
            umma<false>(tmem_base, db, db, IDESC, acc);                       // G[0:128, 0:256]
            if (two_halves) umma<false>(tmem_base + 256, da1, da1, IDESC_SYM, acc);   // G[128:256, 128:256]
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(done);
    }
  } else {
    // ---------------- producers (warps 1..4) ----------------
    const int p = threadIdx.x - 32;
    const int c = p & 15, rsub = p >> 4;                            // 16-byte chunk of the 256-byte row segment; row % 8
    const uint32_t dst_off = (uint32_t)((c >> 3) * STAGE_BYTES + rsub * 128 + (((c & 7) ^ rsub) << 4));
    const uint32_t smem_base = smem_u32(smem);
    auto signal = [&](int64_t i) {                                  // super-stage i has landed (this thread's part)
      fence_proxy_async();
      lg_arrive(&full[i % LG_NSS]);
    };
    for (int64_t i = 0; i < my_super; ++i) {
      const int s = (int)(i % LG_NSS);
      if (i >= LG_NSS) mbar_wait(&empty[s], (uint32_t)((i / LG_NSS - 1) & 1));
      const int64_t col = (blockIdx.x + i * gridDim.x) * (LG_SS_TILES * BK) + c * 8;      // bf16 column of this chunk
      int64_t left = (F - col) * 2;
      left = left < 0 ? 0 : (left > 16 ? 16 : left);
      const uint32_t dst0 = smem_base + s * LG_SS_BYTES + dst_off;
      // rows >= B are never read back (a Gram entry only depends on its own two rows), so their 8-row groups are
      // not even zero-filled; the ragged last group is
      const int nq = (B + 7) >> 3;
#pragma unroll 8
      for (int q = 0; q < nq; ++q) {
        const int row = q * 8 + rsub;
        const uint32_t nbytes = (row < B) ? (uint32_t)left : 0u;
        const __nv_bfloat16* src = x + (nbytes ? (int64_t)row * F + col : 0);
        lg_cp_async16(dst0 + q * 1024, src, nbytes);
      }
      lg_commit();
      if (i >= LG_LA) { lg_wait<LG_LA>(); signal(i - LG_LA); }
    }
    // drain the last LG_LA groups in order
    lg_wait<0>();
    for (int64_t i = (my_super > LG_LA ? my_super - LG_LA : 0); i < my_super; ++i) signal(i);

    // ---------------- epilogue ----------------
    mbar_wait(done, 0);
    tc_fence_after();
    const int qd = warp & 3;                       // TMEM lane quarter this warp may access (warp id % 4)
    float* tr = epi + (warp - 1) * (32 * 33);
    float* out = partials + (size_t)blockIdx.x * B * B;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int row0 = half * 128 + qd * 32;
      if (row0 >= B) continue;
#pragma unroll 1
      for (int col0 = half * 128; col0 < B; col0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + half * 256 + (col0 - half * 128), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        if (col0 + lane < B) {
          for (int r = 0; r < 32; ++r) {
            if (row0 + r >= B) break;
            out[(size_t)(row0 + r) * B + col0 + lane] = tr[r * 33 + lane];
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------
// CTA-pair variant (128 < B <= 256): two SMs of one TPC run ONE tcgen05.mma.cta_group::2 per k-step.
// Why: with one CTA per SM the kernel was bound by shared-memory operand reads, not by HBM or the tensor pipe --
// the load pipeline alone runs at 6.7 TB/s and the M = 128 MMAs alone at their floor, but together they took the
// SUM of both times (20 KB of operand reads per 16-column k-step at ~107 B/clk next to the incoming copies).
// In a pair, CTA r keeps only rows [128 r, 128 r + 128) of every k-tile (half the HBM and smem-write traffic per
// SM); the M = 256, N = 256 MMA takes A = the local 128 rows and B = both halves (each SM supplies its 128 rows,
// the pair link broadcasts them), so per SM and k-step 8 KB of operand reads feed twice the math, and the full
// G comes out (no mirror pass).  Accumulator: 128 lanes x 256 fp32 columns of TMEM per CTA.
//   warp 0      rank-0 CTA: MMA issuer for the pair; both CTAs: TMEM alloc/dealloc (cta_group::2)
//   warps 1-4   producers (cp.async, hand-swizzled, 2 k-tiles = 256 B per row and request), then the epilogue
// full[s] lives in the leader CTA and counts the 256 producer threads of BOTH CTAs (remote mbarrier.arrive through
// mapa); empty[s] / done are signalled in both CTAs by one multicast tcgen05.commit.
constexpr int P2_ROWS = 128;                                    // rows per CTA
constexpr int P2_TILE_BYTES = P2_ROWS * BK * 2;                 // 16 KB
constexpr int P2_SS_TILES = 2;
constexpr int P2_SS_BYTES = P2_SS_TILES * P2_TILE_BYTES;        // 32 KB
constexpr int P2_NSS = 6;
constexpr int P2_LA = 3;                                        // publish distance (4 super-stages = 128 KB in flight per SM)
constexpr int P2_SMEM_BYTES = P2_NSS * P2_SS_BYTES + EPI_FLOATS * 4 + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void arrive_cluster(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {          // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LG_THREADS, 1)
gram_bf16_pair_kernel(const __nv_bfloat16* __restrict__ x, int B, int64_t F, int64_t nsuper, float* __restrict__ partials) {
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs: the dynamic smem base is the same, so aligning by address keeps them equal
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + P2_NSS * P2_SS_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(epi + EPI_FLOATS);
  uint64_t* empty = full + P2_NSS;
  uint64_t* done = empty + P2_NSS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P2_NSS; ++s) { mbar_init(&full[s], 2 * (LG_PRODUCERS / 32)); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                              // both CTAs: barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t my_super = (nsuper - pair + npairs - 1) / npairs;     // super-tiles pair, pair + npairs, ...

  if (warp == 0) {
    if (rank == 0 && lane == 0) {
      // ---------------- MMA issuer (leader CTA only) ----------------
      constexpr uint32_t IDESC = make_idesc(1u /*bf16*/, 256u, 256u);
      for (int64_t i = 0; i < my_super; ++i) {
        const int s = (int)(i % P2_NSS);
        wait_cluster(&full[s], (uint32_t)((i / P2_NSS) & 1));
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < P2_SS_TILES; ++j) {
          const uint32_t sb = smem_u32(smem + s * P2_SS_BYTES + j * P2_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t d = desc_sw128(sb + ks * 32);           // A: 128 local rows per CTA; B: the same rows = N half
            umma2(tmem_base, d, d, IDESC, (i > 0 || j > 0 || ks > 0) ? 1u : 0u);
          }
        }
        umma2_commit_both(&empty[s]);
      }
      umma2_commit_both(done);
    }
  } else {
    // ---------------- producers (warps 1..4 of both CTAs) ----------------
    const int p = threadIdx.x - 32;
    const int c = p & 15, rsub = p >> 4;
    const uint32_t dst_off = (uint32_t)((c >> 3) * P2_TILE_BYTES + rsub * 128 + (((c & 7) ^ rsub) << 4));
    const uint32_t smem_base = smem_u32(smem);
    const int row_base = (int)rank * P2_ROWS;
    auto signal = [&](int64_t i) {                                  // this warp's part of super-stage i has landed
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive_cluster(map_to_cta(smem_u32(&full[i % P2_NSS]), 0));     // one (remote) arrive per warp
    };
    for (int64_t i = 0; i < my_super; ++i) {
      const int s = (int)(i % P2_NSS);
      if (i >= P2_NSS) wait_cluster(&empty[s], (uint32_t)((i / P2_NSS - 1) & 1));
      const int64_t col = (pair + i * npairs) * (P2_SS_TILES * BK) + c * 8;
      int64_t left = (F - col) * 2;
      left = left < 0 ? 0 : (left > 16 ? 16 : left);
      const uint32_t dst0 = smem_base + s * P2_SS_BYTES + dst_off;
#pragma unroll 8
      for (int q = 0; q < P2_ROWS / 8; ++q) {
        const int row = row_base + q * 8 + rsub;
        const uint32_t nbytes = (row < B) ? (uint32_t)left : 0u;
        const __nv_bfloat16* src = x + (nbytes ? (int64_t)row * F + col : 0);
        lg_cp_async16(dst0 + q * 1024, src, nbytes);
      }
      lg_commit();
      if (i >= P2_LA) { lg_wait<P2_LA>(); signal(i - P2_LA); }
    }
    lg_wait<0>();
    for (int64_t i = (my_super > P2_LA ? my_super - P2_LA : 0); i < my_super; ++i) signal(i);

    // ---------------- epilogue: this CTA's 128 rows x B columns ----------------
    wait_cluster(done, 0);
    tc_fence_after();
    const int qd = warp & 3;
    float* tr = epi + (warp - 1) * (32 * 33);
    float* out = partials + (size_t)pair * B * B;
    const int row0 = row_base + qd * 32;
    if (row0 < B) {
#pragma unroll 1
      for (int col0 = 0; col0 < B; col0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + col0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        if (col0 + lane < B) {
          for (int r = 0; r < 32; ++r) {
            if (row0 + r >= B) break;
            out[(size_t)(row0 + r) * B + col0 + lane] = tr[r * 33 + lane];
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                              // the peer may still be reading this CTA's TMEM-side state
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// Sum of the per-CTA partials.  32 consecutive elements per block; the block's 4 warps each sum a
// quarter of the parts (coalesced 128-byte rows, 4 loads in flight per thread) and warp 0 combines the
// four sums in fixed order (deterministic).  Elements of the never-computed lower-left block
// (i >= 128, j < 128) are skipped here; the thread that finishes G[j][i] of the upper-right block writes them too.
__global__ void __launch_bounds__(128)
gram_bf16_reduce_kernel(const float* __restrict__ partials, int nparts, int B, float scale, int full, float* __restrict__ G) {
  __shared__ float part[4][32];
  const size_t bb = (size_t)B * B;
  const size_t e = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  const int i = (int)(e / B), j = (int)(e - (size_t)i * B);
  const bool live = e < bb && (full || !(i >= 128 && j < 128));
  float acc = 0.f;
  if (live) {
#pragma unroll 4
    for (int p = w; p < nparts; p += 4) acc += partials[(size_t)p * bb + e];
  }
  part[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && live) {
    const float v = (((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x]) * scale;
    G[e] = v;
    if (!full && i < 128 && j >= 128) G[(size_t)j * B + i] = v;           // the lower-left block is the transpose of this one
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace g16
}  // namespace alignq

using namespace alignq;

extern "C" size_t alignq_gram_bf16_ws_bytes(int B) {
  if (B < 1) return 0;
  return (size_t)ALIGNQ_NUM_SMS * (size_t)B * B * sizeof(float);
}

extern "C" int alignq_gram_bf16(const void* x_bf16, int B, int64_t F, int divide_by_F, float* G, void* ws,
                                size_t ws_bytes, alignq_stream_t stream) {
  using namespace g16;
  if (B < 1 || B > ROWS || F < 1 || !x_bf16 || !G || !ws) return ALIGNQ_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x_bf16) & 15u) || (F % 8) != 0) return ALIGNQ_EALIGN;   // TMA: 16-byte aligned rows
  const int64_t cap = (int64_t)(ws_bytes / ((size_t)B * B * sizeof(float)));
  if (cap < 1) return ALIGNQ_ENOSPACE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // default: the cp.async-fed single-CTA pipeline; ALIGNQ_GRAM16_PATH=tma|pair selects the other two for A/B runs
  static const int path = []() { const char* e = getenv("ALIGNQ_GRAM16_PATH"); return e ? (int)e[0] : 0; }();
  const int use_tma = path == 't';
  int64_t grid;
  int full = 0;
  const int use_pair = path == 'p' && B > P2_ROWS;
  if (!use_tma && use_pair) {
    // CTA pairs (see gram_bf16_pair_kernel): one partial per pair, full matrix
    const int64_t nsuper = (F + P2_SS_TILES * BK - 1) / (P2_SS_TILES * BK);
    cudaError_t e = cudaFuncSetAttribute(gram_bf16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    // The kernel is persistent: every pair must be resident at once.  Not every TPC has both SMs available, so ask
    // the runtime how many 2-CTA clusters fit instead of assuming 148 / 2.
    static int max_pairs = 0;
    if (max_pairs == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ALIGNQ_NUM_SMS, 1, 1);
      cfg.blockDim = dim3(LG_THREADS, 1, 1);
      cfg.dynamicSmemBytes = P2_SMEM_BYTES;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gram_bf16_pair_kernel, &cfg) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = ALIGNQ_NUM_SMS / 2; }
      max_pairs = n;
      if (getenv("ALIGNQ_DEBUG")) fprintf(stderr, "[alignq] gram_bf16: %d co-resident CTA pairs\n", n);
    }
    int64_t npairs = nsuper < max_pairs ? nsuper : max_pairs;
    if (const char* e2 = getenv("ALIGNQ_GRAM16_PAIRS")) { const int v = atoi(e2); if (v >= 1 && v < npairs) npairs = v; }
    if (npairs > cap) npairs = cap;
    gram_bf16_pair_kernel<<<(unsigned)(2 * npairs), LG_THREADS, P2_SMEM_BYTES, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_bf16), B, F, nsuper, reinterpret_cast<float*>(ws));
    grid = npairs;
    full = 1;
  } else if (!use_tma) {
    // default: cp.async-fed pipeline (see gram_bf16_ldgsts_kernel)
    const int64_t nsuper = (F + LG_SS_TILES * BK - 1) / (LG_SS_TILES * BK);
    grid = nsuper < ALIGNQ_NUM_SMS ? nsuper : ALIGNQ_NUM_SMS;
    if (grid > cap) grid = cap;
    cudaError_t e = cudaFuncSetAttribute(gram_bf16_ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    gram_bf16_ldgsts_kernel<<<(unsigned)grid, LG_THREADS, LG_SMEM_BYTES, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(x_bf16), B, F, nsuper, reinterpret_cast<float*>(ws));
  } else {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return ALIGNQ_EINVAL;
    alignas(64) CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)F, (cuuint64_t)B};
    const cuuint64_t gstride[1] = {(cuuint64_t)F * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)ROWS};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_bf16), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ALIGNQ_EINVAL;
    const int64_t ktiles = (F + BK - 1) / BK;
    grid = ktiles < ALIGNQ_NUM_SMS ? ktiles : ALIGNQ_NUM_SMS;
    if (grid > cap) grid = cap;
    cudaError_t e = cudaFuncSetAttribute(gram_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    gram_bf16_kernel<<<(unsigned)grid, NTHREADS, SMEM_BYTES, s>>>(tmap, B, ktiles, reinterpret_cast<float*>(ws));
  }
  ALIGNQ_LAUNCH_CHECK();
  const size_t bb = (size_t)B * B;
  gram_bf16_reduce_kernel<<<(unsigned)((bb + 31) / 32), 128, 0, s>>>(reinterpret_cast<float*>(ws), (int)grid, B,
                                                                     divide_by_F ? 1.0f / (float)F : 1.0f, full, G);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
