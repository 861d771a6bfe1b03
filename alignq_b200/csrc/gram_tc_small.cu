// Small-batch (B <= 32) variant of the fused tcgen05 correlation forward (gram_tc.cu) -- config 5 of
// BASELINE.json (ResNet-50 DANN, 28 images per GPU, F up to 802 816).  Same algorithm and numerics modes;
// what changes is the shape of the work:
//   * tiles are 64 feature columns wide (not 32) and only 32 batch rows tall, so the 512 threads own
//     2 rows x 2 columns each and no thread idles on rows >= B;
//   * the MMA is M = 64 (the smallest cta_group::1 shape), N = 32: operand tiles are 8 row groups instead of
//     16 (rows 32..63 stay zero), accumulators are 32 TMEM columns per Gram;
//   * TMEM layout for M = 64: row m lives in lane (m % 16) + 32 * (m / 16)  (cute tmem_frg_1sm, M_MMA = 64).
#include <cuda_bf16.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "tc_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace tcs {

using namespace tc;

constexpr int KB = 64;            // feature columns per tile
constexpr int NT = 512, NW = 16;
constexpr int RB = 32;            // batch rows per tile (B <= 32)
constexpr int LBO = 144;
constexpr int NRAW = 6;
constexpr int RAW_TILE = RB * KB * 4;      // 8 KB

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MODE, bool FUSED>
struct Cfg {
  static constexpr bool TF32 = (MODE == ALIGNQ_GRAM_TF32X3);
  static constexpr int ESZ = TF32 ? 4 : 2;
  static constexpr int CH = 16 / ESZ;
  static constexpr int NCH = KB / CH;                      // 16 / 8 core matrices along K
  static constexpr int SBO = NCH * LBO;
  static constexpr int TILE_BYTES = 8 * SBO;               // 64 rows (8 groups); rows >= 32 are zero padding of the M = 64 MMA
  static constexpr int NSRC = FUSED ? 2 : 1;
  static constexpr int NOPER = NSRC * (TF32 ? 2 : 1);
  static constexpr int NACC = NSRC;
  static constexpr int UMMA_K = 32 / ESZ;
  static constexpr int KSTEPS = KB / UMMA_K;               // 8 / 4
  static constexpr int TMEM_COLS = NACC * 32 < 32 ? 32 : NACC * 32;
  static constexpr int STAGE_BYTES = NOPER * TILE_BYTES;
  static constexpr int OFF_RAW = 2 * STAGE_BYTES;
  static constexpr int OFF_RED = OFF_RAW + NRAW * RAW_TILE;
  static constexpr int OFF_CS = OFF_RED + NW * KB * 16;
  static constexpr int OFF_BAR = OFF_CS + KB * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 64;
  static constexpr uint32_t IDESC = make_idesc(TF32 ? 2u : 1u, 64u, 32u);
};

template <int MODE, bool FUSED>
__global__ void __launch_bounds__(NT, 1)
gram_tc_small_kernel(const float* __restrict__ x, int B, int64_t F, float eps, ActQ q, float* __restrict__ y,
                     float* __restrict__ partials, int64_t ntiles) {
  using C = Cfg<MODE, FUSED>;
  extern __shared__ __align__(128) uint8_t smem[];
  float4* red = reinterpret_cast<float4*>(smem + C::OFF_RED);
  float4* colstat = reinterpret_cast<float4*>(smem + C::OFF_CS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < C::OFF_RAW / 16; i += NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float invB = 1.0f / (float)B, invBm1 = 1.0f / (float)(B - 1);

  auto fetch_async = [&](int64_t tile, int slot) {          // 32 rows x 16 chunks of 16 B: one per thread
    if (tile < ntiles) {
      const int r = threadIdx.x >> 4, c16 = threadIdx.x & 15;
      const int64_t f = tile * KB + c16 * 4;
      int64_t left = (F - f) * 4;
      left = left < 0 ? 0 : (left > 16 ? 16 : left);
      const uint32_t nbytes = (r < B) ? (uint32_t)left : 0u;
      const float* src = x + (nbytes ? (int64_t)r * F + f : 0);
      cp_async16_zfill(smem_u32(smem + C::OFF_RAW + slot * RAW_TILE) + r * (KB * 4) + c16 * 16, src, nbytes);
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int p = 0; p < NRAW - 1; ++p) fetch_async(blockIdx.x + (int64_t)p * gridDim.x, p);
  cp_async_wait<NRAW - 2>();
  __syncthreads();

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    uint8_t* st = smem + s * C::STAGE_BYTES;
    fetch_async(tile + (int64_t)(NRAW - 1) * gridDim.x, (it + NRAW - 1) % NRAW);
    const float* raw = reinterpret_cast<const float*>(smem + C::OFF_RAW + (it % NRAW) * RAW_TILE);
    // ---- 1. this thread's 2 rows x 2 columns ---------------------------------------------------------
    float xv[2][2], tv[2][2], px[2], pt[2], s1[2], s2[2], u1[2], u2[2];
    bool colv[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = lane + 32 * j;
      const int64_t f = tile * KB + c;
      colv[j] = f < F;
      px[j] = raw[c];
      pt[j] = FUSED ? act_map_t(px[j], q.ar) : 0.f;
      s1[j] = s2[j] = u1[j] = u2[j] = 0.f;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = warp + NW * i;
        const bool v = colv[j] && r < B;
        xv[i][j] = raw[r * KB + c];
        const float d = xv[i][j] - px[j];
        s1[j] += v ? d : 0.f;
        s2[j] = v ? fmaf(d, d, s2[j]) : s2[j];
        if (FUSED) {
          tv[i][j] = act_map_t(xv[i][j], q.ar);
          if (v && y) y[(int64_t)r * F + f] = act_quant_from_t(tv[i][j], q);
          const float e = tv[i][j] - pt[j];
          u1[j] += v ? e : 0.f;
          u2[j] = v ? fmaf(e, e, u2[j]) : u2[j];
        }
      }
      red[warp * KB + c] = make_float4(s1[j], s2[j], u1[j], u2[j]);
    }
    __syncthreads();
    // ---- 2. column statistics: warps 0 and 1 finish 32 columns each --------------------------------
    if (warp < 2) {
      const int c = lane + 32 * warp;
      float S1 = 0.f, S2 = 0.f, U1 = 0.f, U2 = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const float4 p = red[w * KB + c]; S1 += p.x; S2 += p.y; U1 += p.z; U2 += p.w; }
      const float pxc = raw[c], ptc = FUSED ? act_map_t(pxc, q.ar) : 0.f;
      float vx = (S2 - S1 * S1 * invB) * invBm1;
      vx = (vx < 0.f) ? 0.f : vx;
      float4 o = make_float4(pxc + S1 * invB, 1.0f / (sqrtf(vx) + eps), 0.f, 0.f);
      if (FUSED) {
        float vt = (U2 - U1 * U1 * invB) * invBm1;
        vt = (vt < 0.f) ? 0.f : vt;
        o.z = ptc + U1 * invB;
        o.w = 1.0f / (sqrtf(vt) + eps);
      }
      colstat[c] = o;
    }
    __syncthreads();
    // ---- 3. standardise, convert, store operands -------------------------------------------------------
    if (it >= 2) mbar_wait(&bars[s], ((it >> 1) - 1) & 1);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = lane + 32 * j;
      const float4 cs = colstat[c];
      uint8_t* dst0 = st + (c / C::CH) * LBO + (c % C::CH) * C::ESZ;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = warp + NW * i;
        if (r >= B) break;
        uint8_t* dst = dst0 + (r >> 3) * C::SBO + (r & 7) * 16;
        const float a = colv[j] ? (xv[i][j] - cs.x) * cs.y : 0.f;
        const float b = (FUSED && colv[j]) ? (tv[i][j] - cs.z) * cs.w : 0.f;
        if (C::TF32) {
          const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
          *reinterpret_cast<float*>(dst) = ah;
          *reinterpret_cast<float*>(dst + C::TILE_BYTES) = a - ah;
          if (FUSED) {
            const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
            *reinterpret_cast<float*>(dst + 2 * C::TILE_BYTES) = bh;
            *reinterpret_cast<float*>(dst + 3 * C::TILE_BYTES) = b - bh;
          }
        } else {
          *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(a);
          if (FUSED) *reinterpret_cast<__nv_bfloat16*>(dst + C::TILE_BYTES) = __float2bfloat16_rn(b);
        }
      }
    }
    fence_proxy_async();
    cp_async_wait<NRAW - 2>();
    __syncthreads();
    // ---- 4. MMAs (M = 64, N = 32) ------------------------------------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sb = smem_u32(st);
#pragma unroll
      for (int ks = 0; ks < C::KSTEPS; ++ks) {
        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
        const uint32_t koff = ks * 2 * LBO;
#pragma unroll
        for (int src = 0; src < C::NSRC; ++src) {
          if (C::TF32) {
            const uint64_t dh = make_desc(sb + (2 * src) * C::TILE_BYTES + koff, LBO, C::SBO);
            const uint64_t dl = make_desc(sb + (2 * src + 1) * C::TILE_BYTES + koff, LBO, C::SBO);
            umma<true>(tmem_base + src * 32, dh, dh, C::IDESC, acc);
            umma<true>(tmem_base + src * 32, dh, dl, C::IDESC, 1u);
            umma<true>(tmem_base + src * 32, dl, dh, C::IDESC, 1u);
          } else {
            const uint64_t d = make_desc(sb + src * C::TILE_BYTES + koff, LBO, C::SBO);
            umma<false>(tmem_base + src * 32, d, d, C::IDESC, acc);
          }
        }
      }
      umma_commit(&bars[s]);
    }
  }
  // ---- epilogue: row m of the M = 64 accumulator sits in lane (m % 16) + 32 (m / 16) ----------------------
  if (threadIdx.x == 0) umma_commit(&bars[2]);
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  if (warp < 2) {
    float* out = partials + (size_t)blockIdx.x * C::NACC * B * B;
#pragma unroll 1
    for (int a = 0; a < C::NACC; ++a) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + a * 32, v);
      const int row = 16 * warp + lane;
      if (lane < 16 && row < B) {
        float* o = out + ((size_t)a * B + row) * B;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < B) o[j] = __uint_as_float(v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int MODE, bool FUSED>
static int launch(const float* x, int B, int64_t F, float eps, ActQ q, float* y, float* partials, int64_t cap,
                  int* nparts, cudaStream_t s) {
  using C = Cfg<MODE, FUSED>;
  const int64_t ntiles = (F + KB - 1) / KB;
  int64_t grid = (ntiles + 3) / 4;
  if (grid > ALIGNQ_NUM_SMS) grid = ALIGNQ_NUM_SMS;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaError_t e = cudaFuncSetAttribute(gram_tc_small_kernel<MODE, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  gram_tc_small_kernel<MODE, FUSED><<<(unsigned)grid, NT, C::SMEM_BYTES, s>>>(x, B, F, eps, q, y, partials, ntiles);
  ALIGNQ_LAUNCH_CHECK();
  *nparts = (int)grid;
  return ALIGNQ_OK;
}

}  // namespace tcs

// Partials [cta][nacc][B][B] like gram_tc.cu; the caller reduces them with the same kernel.
int gram_tc_small_partials(const float* x, int B, int64_t F, float eps, ActQ q, int fused, float* y, float* partials,
                           int64_t cap, int gram_mode, int* nparts, cudaStream_t s) {
  if (B > 32 || B < 2) return ALIGNQ_ERANGE;
  if (gram_mode == ALIGNQ_GRAM_TF32X3)
    return fused ? tcs::launch<ALIGNQ_GRAM_TF32X3, true>(x, B, F, eps, q, y, partials, cap, nparts, s)
                 : tcs::launch<ALIGNQ_GRAM_TF32X3, false>(x, B, F, eps, q, y, partials, cap, nparts, s);
  if (gram_mode == ALIGNQ_GRAM_BF16)
    return fused ? tcs::launch<ALIGNQ_GRAM_BF16, true>(x, B, F, eps, q, y, partials, cap, nparts, s)
                 : tcs::launch<ALIGNQ_GRAM_BF16, false>(x, B, F, eps, q, y, partials, cap, nparts, s);
  return ALIGNQ_EINVAL;
}

}  // namespace alignq
