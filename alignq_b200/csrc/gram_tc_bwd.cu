// tcgen05 backward of the fused activation quantizer + ADMM correlation term:
//   gx = corr_bwd(x, -g dLdD) + (corr_bwd(t, +g dLdD) + gy) * 2 ar phi(x),      g = d(total)/d(trans_loss)
//   corr_bwd(X, dG): gS = (dG + dG^T) Xs / F;  per column  gx_i = (gS_i - mean_i gS) / (sd + eps)
//                                                                 - c_i sum_i(gS_i c_i) / ((sd + eps)^2 (B - 1) sd)
// i.e. the autograd backward of activation_quantize_fn.forward (cdf_alignment_admm/resnet-56-cifar-10/
// model/quantization.py:109-123) through corr() (:134-137), ADMM.forward (utils/admm.py:24-33) and the
// straight-through rounding (:33-36) in ONE pass over x and gy.
//
// The two [B,B] x [B,32] products per 32-column tile run on the tensor cores:
//   D[i, n] = sum_j Wsym[i, j] * Xs[j, n]          M = 128 (i), N = 32 (columns), K = 128 (batch j)
// A = Wsym (fixed for the layer: split once, kept in shared memory in the canonical K-major core-matrix
// layout), B = the standardised tile transposed (N x K, K-major).  Operand precision (NS = number of bf16
// terms per fp32 value):
//   NS = 3 (gram_mode tf32x3, the fp32-parity mode): v = H + M + L, three bf16 terms = 24 mantissa bits, and
//          SIX kind::f16 MMAs per k-step: H H into the main accumulator, H M + M H + H L + L H + M M into a
//          second one (the tensor core truncates when it adds into the fp32 accumulator, ~4.6e-8 per
//          accumulate: the five small products must not ride on the large one); dropped terms <= 2^-24.
//          Measured against fp64 autograd: at the level of the reference's own fp32 autograd (1e-6).
//          (Round 1 used two terms, 16 bits: 3e-5 -- above north_star's 1e-5.)
//   NS = 1 (gram_mode bf16): one bf16 term, one MMA per k-step (1e-2 mode).
// fp32 accumulators in TMEM (32 columns per source and accumulator).  Thread (warp w, lane n) owns rows 8w..8w+7 of
// column n, so one 16-byte store fills a whole core-matrix row of the B operand and all global
// accesses are coalesced 128-byte rows.  The accumulators come back through tcgen05.ld + a smem
// transpose; the per-column reductions reuse the forward's two-level warp reduction.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "tc_common.cuh"
#include "tc_small_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace tcb {

using namespace tc;

constexpr int KB = 32;                        // columns per tile
constexpr int NT = 512, NW = 16, RPT = 8;     // thread (w, n): rows 8w .. 8w+7 of column n
constexpr int LBO = 144;                      // B operand: K-adjacent core matrices (padded: conflict-free stores)
constexpr int SBO = 16 * LBO;                 // 8-row groups: K = 128 bf16 = 16 core matrices
constexpr int B_TILE = 4 * SBO;               // Xs^T operand, 32 rows (columns of x)     9 216 B
constexpr int LBO_A = 128;                    // A operand is written once per CTA: dense
constexpr int SBO_A = 16 * LBO_A;
constexpr int A_TILE = 16 * SBO_A;            // Wsym operand, 128 rows                  32 768 B
constexpr int NRAW = 4;
constexpr int RAW_TILE = 128 * KB * 4;        // 16 KB
constexpr int GS_LD = 33;
constexpr int GS_BYTES = 2 * 128 * GS_LD * 4; // [2][128][33] floats
template <int NS>
struct Lay {                                  // NS operand terms per value: A [NS], B [x: NS, t: NS]
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + NS * A_TILE;
  // the gS staging tile is only alive between the MMAs of a tile and the end of that tile: when the B operands
  // are large enough (NS = 3: 55 KB) it lives on top of them
  static constexpr bool GS_ON_B = (2 * NS * B_TILE >= GS_BYTES);
  static constexpr int OFF_RAW = OFF_B + 2 * NS * B_TILE;
  static constexpr int OFF_GS = GS_ON_B ? OFF_B : OFF_RAW + NRAW * RAW_TILE;
  static constexpr int OFF_RED = OFF_RAW + NRAW * RAW_TILE + (GS_ON_B ? 0 : GS_BYTES);   // [NW][32] float4
  static constexpr int OFF_CS = OFF_RED + NW * KB * 16;      // column stats: 3 x [32] float4
  static constexpr int OFF_BAR = OFF_CS + 3 * KB * 16;
  static constexpr int SMEM_BYTES = OFF_BAR + 64;
  static constexpr int ACC_PER_SRC = (NS > 1) ? 2 : 1;       // main + cross accumulator
  static constexpr int TMEM_COLS = (NS > 1) ? 128 : 64;
};
static_assert(Lay<3>::SMEM_BYTES <= 232448, "gram_tc_bwd: shared memory");

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NS>
__global__ void __launch_bounds__(NT, 1)
gram_tc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ Wsym, int Bp,
                   const float* __restrict__ gloss, int B, int64_t F, float ar, float eps, int64_t ntiles,
                   float* __restrict__ gx) {
  pdl_wait();                         // programmatic dependent of wsym_kernel: nothing is touched before this
  using LY = Lay<NS>;
  constexpr int OFF_A = LY::OFF_A, OFF_B = LY::OFF_B, OFF_RAW = LY::OFF_RAW;
  extern __shared__ __align__(128) uint8_t smem[];
  float* GS = reinterpret_cast<float*>(smem + LY::OFF_GS);
  float4* red = reinterpret_cast<float4*>(smem + LY::OFF_RED);
  float4* cs = reinterpret_cast<float4*>(smem + LY::OFF_CS);      // [0..31] x: (mean, rinv, sd, -)  [32..63] t  [64..95] second pass
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + LY::OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- setup: clear operands, split Wsym into the A tiles, barrier, TMEM ---------------------------
  for (int i = threadIdx.x; i < (OFF_RAW - OFF_A) / 16; i += NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int c = threadIdx.x; c < 128 * 16; c += NT) {            // chunk (row i, 8 consecutive j)
    const int i = c >> 4, ch = c & 15;
    if (i < B && 8 * ch < B) {
      const float* src = Wsym + (size_t)i * Bp + 8 * ch;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (8 * ch + k < B) ? __ldg(src + k) : 0.f;
      tcsmall::store_chunk_n<NS>(smem + OFF_A + (i >> 3) * SBO_A + (i & 7) * 16 + ch * LBO_A, A_TILE, v);
    }
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, LY::TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const float gl = gloss ? __ldg(gloss) : 1.0f;
  const float sx = -gl / (float)F, st = gl / (float)F;           // corr_bwd(X, -dD), corr_bwd(T, +dD)
  const float invB = 1.0f / (float)B, invBm1 = 1.0f / (float)(B - 1);
  const float gscale = 2.0f * ar * kInvSqrt2Pi;
  constexpr uint32_t IDESC = make_idesc(1u /*bf16*/, 128u, 32u);

  auto fetch_async = [&](int64_t tile, int slot) {
    if (tile < ntiles) {
      const uint32_t dst0 = smem_u32(smem + OFF_RAW + slot * RAW_TILE);
#pragma unroll
      for (int j = 0; j < 1024 / NT; ++j) {
        const int c = threadIdx.x + j * NT;
        const int r = c >> 3, c16 = c & 7;
        const int64_t f = tile * KB + c16 * 4;
        int64_t left = (F - f) * 4;
        left = left < 0 ? 0 : (left > 16 ? 16 : left);
        const uint32_t nbytes = (r < B) ? (uint32_t)left : 0u;
        const float* src = x + (nbytes ? (int64_t)r * F + f : 0);
        cp_async16_zfill(dst0 + r * (KB * 4) + c16 * 16, src, nbytes);
      }
    }
    cp_async_commit();
  };
#pragma unroll 1
  for (int p = 0; p < NRAW - 1; ++p) fetch_async(blockIdx.x + (int64_t)p * gridDim.x, p);
  cp_async_wait<NRAW - 2>();
  __syncthreads();

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    fetch_async(tile + (int64_t)(NRAW - 1) * gridDim.x, (it + NRAW - 1) % NRAW);
    const float* raw = reinterpret_cast<const float*>(smem + OFF_RAW + (it % NRAW) * RAW_TILE);
    const int64_t f = tile * KB + lane;
    const bool colv = f < F;
    // ---- 1. column slice: rows 8w .. 8w+7 of column `lane` ------------------------------------------
    float xv[RPT], tv[RPT], gv[RPT];
    const float px = raw[lane], pt = act_map_t(px, ar);
    float s1 = 0.f, s2 = 0.f, u1 = 0.f, u2 = 0.f;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int r = 8 * warp + k;
      const bool v = colv && r < B;
      xv[k] = raw[r * KB + lane];
      gv[k] = (v && gy) ? __ldg(gy + (int64_t)r * F + f) : 0.f;
      tv[k] = act_map_t(xv[k], ar);
      const float d = xv[k] - px, e = tv[k] - pt;
      s1 += v ? d : 0.f;  s2 = v ? fmaf(d, d, s2) : s2;
      u1 += v ? e : 0.f;  u2 = v ? fmaf(e, e, u2) : u2;
    }
    red[warp * KB + lane] = make_float4(s1, s2, u1, u2);
    __syncthreads();
    if (warp == 0) {
      float S1 = 0.f, S2 = 0.f, U1 = 0.f, U2 = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const float4 p = red[w * KB + lane]; S1 += p.x; S2 += p.y; U1 += p.z; U2 += p.w; }
      float vx = (S2 - S1 * S1 * invB) * invBm1;  vx = vx < 0.f ? 0.f : vx;
      float vt = (U2 - U1 * U1 * invB) * invBm1;  vt = vt < 0.f ? 0.f : vt;
      const float sdx = sqrtf(vx), sdt = sqrtf(vt);
      cs[lane] = make_float4(px + S1 * invB, 1.0f / (sdx + eps), sdx, 0.f);
      cs[32 + lane] = make_float4(pt + U1 * invB, 1.0f / (sdt + eps), sdt, 0.f);
    }
    __syncthreads();
    const float4 cx4 = cs[lane], ct4 = cs[32 + lane];
    // ---- 2. B operands: one 16-byte core-matrix row (8 batch rows) per thread and tile -------------
    {
      float cxs[RPT], cts[RPT];
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const bool v = colv && (8 * warp + k) < B;
        cxs[k] = v ? (xv[k] - cx4.x) * cx4.y : 0.f;
        cts[k] = v ? (tv[k] - ct4.x) * ct4.y : 0.f;
      }
      uint8_t* dst = smem + OFF_B + (lane >> 3) * SBO + (lane & 7) * 16 + warp * LBO;
      tcsmall::store_chunk_n<NS>(dst, B_TILE, cxs);                       // x terms: tiles 0 .. NS-1
      tcsmall::store_chunk_n<NS>(dst + NS * B_TILE, B_TILE, cts);         // t terms: tiles NS .. 2 NS-1
    }
    fence_proxy_async();
    __syncthreads();
    // ---- 3. MMAs -------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + OFF_A), sb = smem_u32(smem + OFF_B);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t koa = ks * 2 * LBO_A, kob = ks * 2 * LBO;
        const uint32_t acc = ks > 0 ? 1u : 0u;
        const uint64_t ah = make_desc(sa + koa, LBO_A, SBO_A);
#pragma unroll
        for (int src = 0; src < 2; ++src) {
          const uint32_t d_main = tmem_base + 32 * LY::ACC_PER_SRC * src, d_cross = d_main + 32;
          const uint64_t bh = make_desc(sb + (NS * src) * B_TILE + kob, LBO, SBO);
          umma<false>(d_main, ah, bh, IDESC, acc);                          // H H
          if (NS == 3) {
            const uint64_t am = make_desc(sa + A_TILE + koa, LBO_A, SBO_A), al = make_desc(sa + 2 * A_TILE + koa, LBO_A, SBO_A);
            const uint64_t bm = make_desc(sb + (NS * src + 1) * B_TILE + kob, LBO, SBO);
            const uint64_t bl = make_desc(sb + (NS * src + 2) * B_TILE + kob, LBO, SBO);
            umma<false>(d_cross, ah, bm, IDESC, acc);                       // the five products <= 2^-8 of H H
            umma<false>(d_cross, am, bh, IDESC, 1u);
            umma<false>(d_cross, ah, bl, IDESC, 1u);
            umma<false>(d_cross, al, bh, IDESC, 1u);
            umma<false>(d_cross, am, bm, IDESC, 1u);
          }
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, (uint32_t)(it & 1));
    tc_fence_after();
    // ---- 4. accumulators -> smem (transpose to the column-per-lane mapping) -----------------------
    if (warp < 4) {
#pragma unroll 1
      for (int src = 0; src < 2; ++src) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * LY::ACC_PER_SRC * src, v);
        float* g = GS + (src * 128 + warp * 32 + lane) * GS_LD;
        if (NS > 1) {
          uint32_t w[32];
          tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * LY::ACC_PER_SRC * src + 32, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
    }
    __syncthreads();
    // ---- 5. per-column reductions of gS and gS * c -----------------------------------------------
    float gsx[RPT], gst[RPT];
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int r = 8 * warp + k;
      const bool v = colv && r < B;
      gsx[k] = v ? sx * GS[r * GS_LD + lane] : 0.f;                         // gS = -+ g Wsym Xs / F
      gst[k] = v ? st * GS[(128 + r) * GS_LD + lane] : 0.f;
      a1 += gsx[k];  a2 = fmaf(gsx[k], xv[k] - cx4.x, a2);
      b1 += gst[k];  b2 = fmaf(gst[k], tv[k] - ct4.x, b2);
    }
    red[warp * KB + lane] = make_float4(a1, a2, b1, b2);
    __syncthreads();
    if (warp == 0) {
      float A1 = 0.f, A2 = 0.f, B1 = 0.f, B2 = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const float4 p = red[w * KB + lane]; A1 += p.x; A2 += p.y; B1 += p.z; B2 += p.w; }
      // (mean gS, coefficient of c) per source; the std term is masked at sd == 0 as torch's std_backward does
      const float kx = (cx4.z > 0.f) ? -A2 * cx4.y * cx4.y * invBm1 / cx4.z : 0.f;
      const float kt = (ct4.z > 0.f) ? -B2 * ct4.y * ct4.y * invBm1 / ct4.z : 0.f;
      cs[64 + lane] = make_float4(A1 * invB, kx, B1 * invB, kt);
    }
    __syncthreads();
    const float4 c2 = cs[64 + lane];
    // ---- 6. combine and store ------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int r = 8 * warp + k;
      if (!(colv && r < B)) continue;
      const float bx = (gsx[k] - c2.x) * cx4.y + c2.y * (xv[k] - cx4.x);
      const float bt = (gst[k] - c2.z) * ct4.y + c2.w * (tv[k] - ct4.x);
      const float vv = __fmul_rn(xv[k], kInvSqrt2);
      const float dphi = gscale * gauss_kernel_from_v(vv);
      gx[(int64_t)r * F + f] = bx + (bt + gv[k]) * dphi;
    }
    cp_async_wait<NRAW - 2>();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, LY::TMEM_COLS);
}

}  // namespace tcb

int gram_tc_backward_small(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B,
                           int64_t F, float ar, float eps, float* gx, int split, cudaStream_t s);

int gram_tc_backward(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B, int64_t F,
                     float ar, float eps, float* gx, int split, cudaStream_t s) {
  using namespace tcb;
  if (B > 128 || B < 2) return ALIGNQ_ERANGE;
  if (B <= 32) return gram_tc_backward_small(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, split, s);   // thread-per-column variant
  if (!aligned16(x) || (F % 4) != 0) return ALIGNQ_EALIGN;       // cp.async row segments
  const int64_t ntiles = (F + KB - 1) / KB;
  const int64_t tiles_per_cta = (ntiles + ALIGNQ_NUM_SMS - 1) / ALIGNQ_NUM_SMS;      // one wave, equal work per CTA
  int64_t grid = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
  if (grid > ALIGNQ_NUM_SMS) grid = ALIGNQ_NUM_SMS;
  if (grid < 1) grid = 1;
  cudaError_t e;
  if (split) {
    e = cudaFuncSetAttribute(gram_tc_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<3>::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(gram_tc_bwd_kernel<3>, dim3((unsigned)grid), dim3(NT), Lay<3>::SMEM_BYTES, s, x, gy, Wsym, Bp, gloss, B, F, ar, eps, ntiles, gx);
    if (e != cudaSuccess) return (int)e;
  } else {
    e = cudaFuncSetAttribute(gram_tc_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<1>::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(gram_tc_bwd_kernel<1>, dim3((unsigned)grid), dim3(NT), Lay<1>::SMEM_BYTES, s, x, gy, Wsym, Bp, gloss, B, F, ar, eps, ntiles, gx);
    if (e != cudaSuccess) return (int)e;
  }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

}  // namespace alignq
