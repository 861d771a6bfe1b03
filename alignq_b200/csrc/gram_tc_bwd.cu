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
// A = Wsym (fixed for the layer: split once into bf16 H + L, kept in shared memory in the canonical
// K-major core-matrix layout), B = the standardised tile transposed (N x K, K-major, bf16 H + L),
// three kind::f16 MMAs per k-step (H H + H L + L H: 16 mantissa bits per operand, ~3e-5 relative),
// fp32 accumulators in TMEM (32 columns per source).  Thread (warp w, lane n) owns rows 8w..8w+7 of
// column n, so one 16-byte store fills a whole core-matrix row of the B operand and all global
// accesses are coalesced 128-byte rows.  The accumulators come back through tcgen05.ld + a smem
// transpose; the per-column reductions reuse the forward's two-level warp reduction.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "tc_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace tcb {

using namespace tc;

constexpr int KB = 32;                        // columns per tile
constexpr int NT = 512, NW = 16, RPT = 8;     // thread (w, n): rows 8w .. 8w+7 of column n
constexpr int LBO = 144;                      // K-adjacent core matrices (padded)
constexpr int SBO = 16 * LBO;                 // 8-row groups: K = 128 bf16 = 16 core matrices
constexpr int A_TILE = 16 * SBO;              // Wsym operand, 128 rows                  36 864 B
constexpr int B_TILE = 4 * SBO;               // Xs^T operand, 32 rows (columns of x)     9 216 B
constexpr int NRAW = 3;
constexpr int RAW_TILE = 128 * KB * 4;        // 16 KB
constexpr int GS_LD = 33;
constexpr int OFF_A = 0;                                   // [H, L]
constexpr int OFF_B = OFF_A + 2 * A_TILE;                  // [xH, xL, tH, tL]
constexpr int OFF_RAW = OFF_B + 4 * B_TILE;
constexpr int OFF_GS = OFF_RAW + NRAW * RAW_TILE;          // [2][128][33] floats
constexpr int OFF_RED = OFF_GS + 2 * 128 * GS_LD * 4;      // [NW][32] float4
constexpr int OFF_CS = OFF_RED + NW * KB * 16;             // column stats: 3 x [32] float4
constexpr int OFF_BAR = OFF_CS + 3 * KB * 16;
constexpr int SMEM_BYTES = OFF_BAR + 64;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& h, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(v);
  l = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

template <bool SPLIT>
__global__ void __launch_bounds__(NT, 1)
gram_tc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ Wsym, int Bp,
                   const float* __restrict__ gloss, int B, int64_t F, float ar, float eps, int64_t ntiles,
                   float* __restrict__ gx) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* GS = reinterpret_cast<float*>(smem + OFF_GS);
  float4* red = reinterpret_cast<float4*>(smem + OFF_RED);
  float4* cs = reinterpret_cast<float4*>(smem + OFF_CS);      // [0..31] x: (mean, rinv, sd, -)  [32..63] t  [64..95] second pass
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
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
      __nv_bfloat16 h[8], l[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) split_bf16(v[k], h[k], l[k]);
      uint8_t* dst = smem + OFF_A + (i >> 3) * SBO + (i & 7) * 16 + ch * LBO;
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
      if (SPLIT)
        *reinterpret_cast<uint4*>(dst + A_TILE) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
    }
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
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
      __nv_bfloat16 h[RPT], l[RPT];
      uint8_t* dst = smem + OFF_B + (lane >> 3) * SBO + (lane & 7) * 16 + warp * LBO;
#pragma unroll
      for (int k = 0; k < RPT; ++k) split_bf16(cxs[k], h[k], l[k]);
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
      if (SPLIT) *reinterpret_cast<uint4*>(dst + B_TILE) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
#pragma unroll
      for (int k = 0; k < RPT; ++k) split_bf16(cts[k], h[k], l[k]);
      *reinterpret_cast<uint4*>(dst + 2 * B_TILE) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
      if (SPLIT) *reinterpret_cast<uint4*>(dst + 3 * B_TILE) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
    }
    fence_proxy_async();
    __syncthreads();
    // ---- 3. MMAs -------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + OFF_A), sb = smem_u32(smem + OFF_B);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t koff = ks * 2 * LBO;
        const uint64_t ah = make_desc(sa + koff, LBO, SBO), al = make_desc(sa + A_TILE + koff, LBO, SBO);
#pragma unroll
        for (int src = 0; src < 2; ++src) {
          const uint64_t bh = make_desc(sb + (2 * src) * B_TILE + koff, LBO, SBO);
          const uint64_t bl = make_desc(sb + (2 * src + 1) * B_TILE + koff, LBO, SBO);
          umma<false>(tmem_base + 32 * src, ah, bh, IDESC, ks > 0 ? 1u : 0u);
          if (SPLIT) {
            umma<false>(tmem_base + 32 * src, ah, bl, IDESC, 1u);
            umma<false>(tmem_base + 32 * src, al, bh, IDESC, 1u);
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
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * src, v);
        float* g = GS + (src * 128 + warp * 32 + lane) * GS_LD;
#pragma unroll
        for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(v[j]);
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
  if (warp == 0) tmem_dealloc(tmem_base, 64);
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
    e = cudaFuncSetAttribute(gram_tc_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    gram_tc_bwd_kernel<true><<<(unsigned)grid, NT, SMEM_BYTES, s>>>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, ntiles, gx);
  } else {
    e = cudaFuncSetAttribute(gram_tc_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    gram_tc_bwd_kernel<false><<<(unsigned)grid, NT, SMEM_BYTES, s>>>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, ntiles, gx);
  }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

}  // namespace alignq
