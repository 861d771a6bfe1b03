// Small-batch (B <= 32) tcgen05 backward of the fused activation quantizer + ADMM correlation term -- config 5
// of BASELINE.json (ResNet-50 DANN, 28 images per GPU, F up to 802 816).  Same math as gram_tc_bwd.cu
// (autograd backward of cdf_alignment_admm/resnet-56-cifar-10/model/quantization.py:109-123 through corr()
// :134-137 and the straight-through rounding :33-36); what changes is who owns what:
//
//   * ONE THREAD OWNS ONE FEATURE COLUMN (all B <= 32 batch rows in registers).  Column statistics, the
//     reductions of gS and gS*c, and the final combine are thread-local: no cross-warp reduction, one
//     __syncthreads per tile.  Global loads/stores are coalesced across the warp (consecutive columns).
//   * the per-tile product is flipped so the feature columns are the MMA's M dimension:
//         D[n, i] = sum_j XsT[n, j] * Wsym[i, j]        M = 128 columns, N = 32 (i), K = 32 (batch j)
//     A = the standardised tile transposed (thread n writes row n: 4 x 16-byte core-matrix rows),
//     B = Wsym (symmetric; split once per CTA).  The accumulator row n lands in TMEM lane n, so
//     tcgen05.ld hands thread n exactly its own column of W Xs: no shared-memory transpose.
//     Operand precision as in gram_tc_bwd.cu: NS = 3 bf16 terms per value (24 mantissa bits, six MMAs per
//     k-step, the five small products in their own accumulator) for gram_mode tf32x3, NS = 1 for bf16.
//   * 128 threads per CTA, several CTAs per SM (NS = 3: 87 KB smem, 128 TMEM columns -> two per SM; NS = 1:
//     three) overlap each other's load -> convert -> MMA -> combine phases.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "tc_common.cuh"
#include "tc_small_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace tcsb {

using namespace tc;
using namespace tcsmall;

constexpr int NT = 128;                 // threads = columns per tile = MMA M
constexpr int RB = 32;                  // batch rows: MMA N and K
constexpr int LBO = 128;                // K-adjacent core matrices, dense
constexpr int SBO = 4 * LBO;            // 8-row groups: K = 32 bf16 = 4 core matrices
constexpr int A_TILE = 16 * SBO;        // XsT operand: 128 rows x 64 B            8 KB
constexpr int W_TILE = 4 * SBO;         // Wsym operand: 32 rows x 64 B             2 KB
template <int NS>
struct Lay {
  static constexpr int OFF_W = 0;                          // [NS] terms
  static constexpr int OFF_A = OFF_W + NS * W_TILE;        // [x: NS terms, t: NS terms]
  static constexpr int OFF_XS = OFF_A + 2 * NS * A_TILE;   // x of the NEXT tile, [32][128] floats (own column per thread)
  static constexpr int OFF_GY = OFF_XS + RB * NT * 4;      // gy of THIS tile, same shape
  static constexpr int OFF_BAR = OFF_GY + RB * NT * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 64;
  static constexpr int CTAS_PER_SM = (NS > 1) ? 2 : 3;
  static constexpr int ACC_PER_SRC = (NS > 1) ? 2 : 1;
  static constexpr int TMEM_COLS = (NS > 1) ? 128 : 64;
};

// RBT = batch rows carried per thread (B rounded up to the next instantiated size; K entries RBT..31 stay zero)
template <int NS, int RBT>
__global__ void __launch_bounds__(NT, Lay<NS>::CTAS_PER_SM)
gram_tc_bwd_small_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ Wsym, int Bp,
                         const float* __restrict__ gloss, int B, int64_t F, float ar, float eps, int64_t ntiles,
                         float* __restrict__ gx) {
  pdl_wait();                         // programmatic dependent of wsym_kernel: nothing is touched before this
  using LY = Lay<NS>;
  constexpr int OFF_W = LY::OFF_W, OFF_A = LY::OFF_A, OFF_XS = LY::OFF_XS, OFF_GY = LY::OFF_GY, OFF_BAR = LY::OFF_BAR;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n = threadIdx.x;

  // ---- setup: clear operands, split Wsym (B operand, row i = output batch index, K = j), barrier, TMEM ----
  for (int i = threadIdx.x; i < OFF_BAR / 16; i += NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int i = n >> 2, ch = n & 3;                          // 32 rows x 4 chunks of 8 consecutive j
    if (i < B && 8 * ch < B) {
      const float* src = Wsym + (size_t)i * Bp + 8 * ch;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (8 * ch + k < B) ? __ldg(src + k) : 0.f;
      store_chunk_n<NS>(smem + OFF_W + (i >> 3) * SBO + (i & 7) * 16 + ch * LBO, W_TILE, v);
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
  uint8_t* arow = smem + OFF_A + (n >> 3) * SBO + (n & 7) * 16;  // this column's row of the A operands

  // Asynchronous, register-free prefetch: thread n copies ITS OWN column (4 bytes per row) of the next tile's x and
  // of this tile's gy into shared memory and later reads back only what it copied, so no barrier is involved.
  const uint32_t xs_u32 = smem_u32(smem + OFF_XS) + n * 4, gs_u32 = smem_u32(smem + OFF_GY) + n * 4;
  const float* xs = reinterpret_cast<const float*>(smem + OFF_XS) + n;
  const float* gs = reinterpret_cast<const float*>(smem + OFF_GY) + n;
  // rows r >= B re-read row 0 (the pivot of the statistics): always a valid address, no per-row predicate
  auto fetch_col = [&](uint32_t dst, const float* src, int64_t tile) {
    const int64_t f = tile * NT + n;
    if (src != nullptr && tile < ntiles && f < F) {
      const float* p = src + f;
#pragma unroll
      for (int r = 0; r < RBT; ++r) cp_async4(dst + r * (NT * 4), p + ((r < B) ? (int64_t)r * F : 0), 4u);
    }
    cp_async_commit();
  };
  fetch_col(xs_u32, x, blockIdx.x);

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int64_t f = tile * NT + n;
    const bool colv = f < F;
    // ---- 1. the column: values, map, statistics (all thread-local) -------------------------------------
    // Rows r >= B carry the pivot (row 0, see fetch_col): they add nothing to the statistics; their operand entries (K index
    // j >= B) meet the zero rows/columns of Wsym, and accumulator entries i >= B come out exactly 0, so nothing
    // below is predicated per element except the global stores.
    cp_async_wait_all();
    float xv[RBT], tv[RBT];
#pragma unroll
    for (int r = 0; r < RBT; ++r) xv[r] = xs[r * NT];
    const float px = xv[0];
    fetch_col(xs_u32, x, tile + gridDim.x);                      // next tile's x (this buffer is in registers now)
    fetch_col(gs_u32, gy, tile);                                 // this tile's gy, read in step 5
    float s1 = 0.f, s2 = 0.f, u1 = 0.f, u2 = 0.f, pt = 0.f;
#pragma unroll
    for (int r = 0; r < RBT; ++r) {
      tv[r] = act_map_t(xv[r], ar);
      if (r == 0) pt = tv[0];
      const float d = xv[r] - px, e = tv[r] - pt;
      s1 += d;  s2 = fmaf(d, d, s2);
      u1 += e;  u2 = fmaf(e, e, u2);
    }
    float vx = (s2 - s1 * s1 * invB) * invBm1;  vx = vx < 0.f ? 0.f : vx;
    float vt = (u2 - u1 * u1 * invB) * invBm1;  vt = vt < 0.f ? 0.f : vt;
    const float sdx = sqrtf(vx), sdt = sqrtf(vt);
    const float mx = px + s1 * invB, mt = pt + u1 * invB;
    const float rx = colv ? 1.0f / (sdx + eps) : 0.f, rt = colv ? 1.0f / (sdt + eps) : 0.f;   // columns >= F: zero rows
    // ---- 2. A operands: row n = this column, K = batch (4 core-matrix rows of 8) --------------------------
#pragma unroll
    for (int c = 0; c < (RBT + 7) / 8; ++c) {
      float cx[8], ct[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = (8 * c + k < RBT) ? 8 * c + k : 0;           // compile-time: K entries >= RBT are zeros
        cx[k] = (8 * c + k < RBT) ? (xv[r] - mx) * rx : 0.f;
        ct[k] = (8 * c + k < RBT) ? (tv[r] - mt) * rt : 0.f;
      }
      store_chunk_n<NS>(arow + c * LBO, A_TILE, cx);
      store_chunk_n<NS>(arow + c * LBO + NS * A_TILE, A_TILE, ct);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- 3. MMAs ---------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sw = smem_u32(smem + OFF_W), sa = smem_u32(smem + OFF_A);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t koff = ks * 2 * LBO;
        const uint32_t acc = ks > 0 ? 1u : 0u;
        const uint64_t wh = make_desc(sw + koff, LBO, SBO);
#pragma unroll
        for (int src = 0; src < 2; ++src) {
          const uint32_t d_main = tmem_base + 32 * LY::ACC_PER_SRC * src, d_cross = d_main + 32;
          const uint64_t ah = make_desc(sa + (NS * src) * A_TILE + koff, LBO, SBO);
          umma<false>(d_main, ah, wh, IDESC, acc);                          // H H
          if (NS == 3) {
            const uint64_t wm = make_desc(sw + W_TILE + koff, LBO, SBO), wl = make_desc(sw + 2 * W_TILE + koff, LBO, SBO);
            const uint64_t am = make_desc(sa + (NS * src + 1) * A_TILE + koff, LBO, SBO);
            const uint64_t al = make_desc(sa + (NS * src + 2) * A_TILE + koff, LBO, SBO);
            umma<false>(d_cross, ah, wm, IDESC, acc);                       // the five products <= 2^-8 of H H
            umma<false>(d_cross, am, wh, IDESC, 1u);
            umma<false>(d_cross, ah, wl, IDESC, 1u);
            umma<false>(d_cross, al, wh, IDESC, 1u);
            umma<false>(d_cross, am, wm, IDESC, 1u);
          }
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, (uint32_t)(it & 1));
    tc_fence_after();
    // ---- 4. x source: gS = -g W Xs / F, its two column sums, the corr_bwd(x) part of gx ---------------------
    float o[RBT];
    {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
      if (NS > 1) {
        uint32_t w[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32, w);
#pragma unroll
        for (int r = 0; r < RBT; ++r) v[r] = __float_as_uint(__uint_as_float(v[r]) + __uint_as_float(w[r]));
      }
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int r = 0; r < RBT; ++r) {
        o[r] = sx * __uint_as_float(v[r]);
        a1 += o[r];
        a2 = fmaf(o[r], xv[r] - mx, a2);
      }
      // the std term is masked at sd == 0 as torch's std_backward does
      const float kx = (sdx > 0.f) ? -a2 * rx * rx * invBm1 / sdx : 0.f;
      const float mg = a1 * invB;
#pragma unroll
      for (int r = 0; r < RBT; ++r) o[r] = (o[r] - mg) * rx + kx * (xv[r] - mx);
    }
    // ---- 5. t source, straight-through factor, store ---------------------------------------------------------
    {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * LY::ACC_PER_SRC, v);
      if (NS > 1) {
        uint32_t w[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * LY::ACC_PER_SRC + 32, w);
#pragma unroll
        for (int r = 0; r < RBT; ++r) v[r] = __float_as_uint(__uint_as_float(v[r]) + __uint_as_float(w[r]));
      }
      tc_fence_before();
      float b1 = 0.f, b2 = 0.f;
#pragma unroll
      for (int r = 0; r < RBT; ++r) {
        const float g = st * __uint_as_float(v[r]);
        v[r] = __float_as_uint(g);
        b1 += g;
        b2 = fmaf(g, tv[r] - mt, b2);
      }
      const float kt = (sdt > 0.f) ? -b2 * rt * rt * invBm1 / sdt : 0.f;
      const float mg = b1 * invB;
      cp_async_wait_all();
      float* gp = gx + f;
#pragma unroll
      for (int r = 0; r < RBT; ++r) {
        const float gv = gs[r * NT];
        const float bt = (__uint_as_float(v[r]) - mg) * rt + kt * (tv[r] - mt);
        const float vv = __fmul_rn(xv[r], kInvSqrt2);
        const float dphi = gscale * gauss_kernel_from_v(vv);
        st_if(gp + (int64_t)r * F, o[r] + (bt + gv) * dphi, colv && r < B);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, LY::TMEM_COLS);
}

}  // namespace tcsb

template <int NS, int RBT>
static int launch_bwd_small(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B, int64_t F,
                            float ar, float eps, float* gx, int64_t ntiles, int64_t grid, cudaStream_t s) {
  using namespace tcsb;
  cudaError_t e = cudaFuncSetAttribute(gram_tc_bwd_small_kernel<NS, RBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Lay<NS>::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  e = launch_pdl(gram_tc_bwd_small_kernel<NS, RBT>, dim3((unsigned)grid), dim3(NT), Lay<NS>::SMEM_BYTES, s, x, gy, Wsym, Bp, gloss, B, F, ar, eps,
                 ntiles, gx);
  if (e != cudaSuccess) return (int)e;
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
template <int NS>
static int launch_bwd_small_b(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B, int64_t F,
                              float ar, float eps, float* gx, int64_t ntiles, int64_t grid, cudaStream_t s) {
  if (B <= 8) return launch_bwd_small<NS, 8>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s);
  if (B <= 16) return launch_bwd_small<NS, 16>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s);
  if (B <= 24) return launch_bwd_small<NS, 24>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s);
  if (B <= 28) return launch_bwd_small<NS, 28>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s);
  return launch_bwd_small<NS, 32>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s);
}

int gram_tc_backward_small(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B,
                           int64_t F, float ar, float eps, float* gx, int split, cudaStream_t s) {
  using namespace tcsb;
  if (B > RB || B < 2) return ALIGNQ_ERANGE;
  const int64_t ntiles = (F + NT - 1) / NT;
  const int per_sm = split ? Lay<3>::CTAS_PER_SM : Lay<1>::CTAS_PER_SM;
  int64_t grid = ntiles;
  if (grid > (int64_t)ALIGNQ_NUM_SMS * per_sm) grid = (int64_t)ALIGNQ_NUM_SMS * per_sm;
  if (grid < 1) grid = 1;
  return split ? launch_bwd_small_b<3>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s)
               : launch_bwd_small_b<1>(x, gy, Wsym, Bp, gloss, B, F, ar, eps, gx, ntiles, grid, s);
}

}  // namespace alignq
