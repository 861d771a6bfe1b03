// Small-batch (B <= 32) variant of the fused tcgen05 correlation forward (gram_tc.cu) -- config 5 of
// BASELINE.json (ResNet-50 DANN, 28 images per GPU, F up to 802 816).  Same algorithm
// (cdf_alignment_admm/resnet-56-cifar-10/model/quantization.py:109-123, corr() :134-137); what changes is who
// owns what:
//
//   * ONE THREAD OWNS ONE FEATURE COLUMN (all B <= 32 batch rows in registers): the map t = (2 Phi(x) - 1) ar,
//     the quantized output y, both column statistics and the standardisation are thread-local -- no cross-warp
//     reduction; global loads/stores are coalesced across the warp (consecutive columns).
//   * operands are MN-major: thread n (feature k = n) writes its 32 standardised batch values as four 16-byte
//     core-matrix rows per operand, x rows 0..31 and t rows 32..63 of ONE 64-row operand [Xs; Ts].  One
//     M = 64, N = 64, K = 16 MMA per k-step then yields both Grams as the diagonal blocks of
//     [Xs; Ts][Xs; Ts]^T (rows 0..31 x cols 0..31 and rows 32..63 x cols 32..63).
//   * numerics: 'tf32x3' mode = bf16 H + L split operands, three products (H H + H L + L H: 16 mantissa bits,
//     the backward's scheme); 'bf16' mode = one product.  The tensor core truncates when it adds into the
//     fp32 accumulator (measured ~5.6e-8 relative per accumulate), so a long accumulation chain drifts: each
//     128-column tile gets a FRESH accumulator (24 accumulates) that is then added into registers with
//     round-to-nearest.  Two accumulator sets / operand stages ping-pong, so the MMAs of tile i overlap the
//     flush of tile i-1 and the loads of tile i+1.
//   * 128 threads per CTA, three CTAs per SM (64 KB smem, 128 TMEM columns each).
#include <cuda_bf16.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "tc_common.cuh"
#include "tc_small_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace tcs {

using namespace tc;
using namespace tcsmall;

constexpr int NT = 128;                 // threads = feature columns per tile = K per tile
constexpr int RB = 32;                  // batch rows per source
constexpr int LBO = 128;                // k-groups (8 features): one 8 x 16 B core matrix
constexpr int SBO = 16 * LBO;           // m-groups (8 batch rows): 16 k-groups per tile
constexpr int OP_TILE = 8 * SBO;        // [x rows 0..31; t rows 0..31] x 128 features, bf16: 16 KB
constexpr int STAGE = 2 * OP_TILE;      // H, L
constexpr int OFF_BAR = 2 * STAGE;
constexpr int SMEM_BYTES = OFF_BAR + 64;
constexpr int CTAS_PER_SM = 3;
constexpr int KSTEPS = NT / 16;

// RBT = batch rows carried per thread (B rounded up to the next instantiated size: rows RBT..31 of the operands stay zero)
template <bool SPLIT, bool FUSED, int RBT>
__global__ void __launch_bounds__(NT, CTAS_PER_SM)
gram_tc_small_kernel(const float* __restrict__ x, int B, int64_t F, float eps, ActQ q, float* __restrict__ y,
                     float* __restrict__ partials, int64_t ntiles, double* __restrict__ zero_acc) {
  pdl_trigger();                      // the finish kernel launches early and waits for this grid (common.cuh)
  constexpr int NCOLS = FUSED ? 64 : 32;                          // accumulator columns per set = MMA N
  // both operands MN-major (bits 15, 16)
  constexpr uint32_t IDESC = make_idesc(1u /*bf16*/, 64u, (uint32_t)NCOLS) | (1u << 15) | (1u << 16);
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n = threadIdx.x;
  if (zero_acc && blockIdx.x == 0 && threadIdx.x < 4) zero_acc[threadIdx.x] = 0.0;     // arms gram_finish_tc_kernel

  for (int i = threadIdx.x; i < OFF_BAR / 16; i += NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 2 * NCOLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float invB = 1.0f / (float)B, invBm1 = 1.0f / (float)(B - 1);
  const bool q32 = q.a_bit == 32, q1 = q.a_bit == 1;
  // accumulator rows of the M = 64 MMA: row m sits in TMEM lane (m % 16) + 32 (m / 16), so lanes 0..15 of
  // warp w hold rows 16w .. 16w+15 (warps 0, 1: x rows; warps 2, 3: t rows -> columns 32..63)
  const uint32_t tmem_mine = tmem_base + ((uint32_t)(warp * 32) << 16) + (warp >= 2 ? 32u : 0u);
  const bool flusher = FUSED || warp < 2;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  auto flush = [&](int set) {
    uint32_t v[32];
    tmem_ld32(tmem_mine + set * NCOLS, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
  };

  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int64_t f = tile * NT + n;
    const bool colv = f < F;
    {                                                            // next tile's lines -> L2 (lane r: row r)
      const int64_t fn = (tile + gridDim.x) * NT + warp * 32;
      if (lane < B && fn < F) prefetch_l2(x + (int64_t)lane * F + fn);
    }
    // ---- 1. the column: values, map, output, statistics (all thread-local) ----------------------------
    // Rows r >= B carry the pivot value (row 0): they add nothing to the statistics, and whatever they put into
    // operand rows >= B only reaches Gram entries with an index >= B, which are never read -- so nothing below
    // is predicated per element except the global loads and stores.
    float xv[RBT], tv[RBT];
    const bool ystore = colv && y != nullptr;
    const float* xp = x + f;
    if (colv) {                                  // rows r >= B re-read row 0: always a valid address
#pragma unroll
      for (int r = 0; r < RBT; ++r) xv[r] = ld_once(xp + ((r < B) ? (int64_t)r * F : 0));
    } else {
#pragma unroll
      for (int r = 0; r < RBT; ++r) xv[r] = 0.f;
    }
    const float px = xv[0];
    float s1 = 0.f, s2 = 0.f, u1 = 0.f, u2 = 0.f, pt = 0.f;
#pragma unroll
    for (int r = 0; r < RBT; ++r) {
      const float d = xv[r] - px;
      s1 += d;  s2 = fmaf(d, d, s2);
      if (FUSED) {
        tv[r] = act_map_t(xv[r], q.ar);
        if (r == 0) pt = tv[0];
        {                                       // uniform_q (QB:110), branch-free so the 32 rows interleave
          const float t = tv[r];
          const float yq = __fmul_rn(rintf(__fmul_rn(t, q.n)), q.inv_n);
          const float ys = (t > 0.0f) ? 1.0f : ((t < 0.0f) ? -1.0f : t);
          const float yv = q32 ? t : (q1 ? ys : yq);
          st_if(y + (int64_t)r * F + f, yv, ystore && r < B);
        }
        const float e = tv[r] - pt;
        u1 += e;  u2 = fmaf(e, e, u2);
      }
    }
    float vx = (s2 - s1 * s1 * invB) * invBm1;  vx = vx < 0.f ? 0.f : vx;
    const float mx = px + s1 * invB, rx = colv ? 1.0f / (sqrtf(vx) + eps) : 0.f;     // columns >= F: zero operand
    float mt = 0.f, rt = 0.f;
    if (FUSED) {
      float vt = (u2 - u1 * u1 * invB) * invBm1;  vt = vt < 0.f ? 0.f : vt;
      mt = pt + u1 * invB;
      rt = colv ? 1.0f / (sqrtf(vt) + eps) : 0.f;
    }
    // ---- 2. operands (stage s is free: its MMAs of tile it-2 were waited for in iteration it-1) ---------
    uint8_t* dst = smem + s * STAGE + (n >> 3) * LBO + (n & 7) * 16;
#pragma unroll
    for (int c = 0; c < (RBT + 7) / 8; ++c) {
      float cx[8], ct[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = (8 * c + k < RBT) ? 8 * c + k : 0;           // compile-time: rows >= RBT are zeros
        cx[k] = (8 * c + k < RBT) ? (xv[r] - mx) * rx : 0.f;
        ct[k] = (FUSED && 8 * c + k < RBT) ? (tv[r] - mt) * rt : 0.f;
      }
      store_chunk<SPLIT>(dst + c * SBO, OP_TILE, cx);
      if (FUSED) store_chunk<SPLIT>(dst + (4 + c) * SBO, OP_TILE, ct);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- 3. MMAs of this tile into accumulator set s (fresh accumulator) ---------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sb = smem_u32(smem + s * STAGE);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint64_t dh = make_desc(sb + ks * 2 * LBO, LBO, SBO);
        umma<false>(tmem_base + s * NCOLS, dh, dh, IDESC, ks > 0 ? 1u : 0u);
        if (SPLIT) {
          const uint64_t dl = make_desc(sb + OP_TILE + ks * 2 * LBO, LBO, SBO);
          umma<false>(tmem_base + s * NCOLS, dh, dl, IDESC, 1u);
          umma<false>(tmem_base + s * NCOLS, dl, dh, IDESC, 1u);
        }
      }
      umma_commit(&bars[s]);
    }
    // ---- 4. previous tile's accumulator -> registers (round-to-nearest adds) -----------------------------
    if (it > 0) {
      mbar_wait(&bars[s ^ 1], (uint32_t)(((it - 1) >> 1) & 1));
      tc_fence_after();
      if (flusher) flush(s ^ 1);
    }
  }
  if (it > 0) {
    const int last = it - 1;
    mbar_wait(&bars[last & 1], (uint32_t)((last >> 1) & 1));
    tc_fence_after();
    if (flusher) flush(last & 1);
  }
  // ---- epilogue: lanes 0..15 of warp w own accumulator rows 16w .. 16w+15 ---------------------------------
  if (flusher && lane < 16) {
    const int a = warp >> 1, row = 16 * (warp & 1) + lane;
    if (row < B) {
      float* o = partials + (((size_t)blockIdx.x * (FUSED ? 2 : 1) + a) * B + row) * B;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < B) o[j] = acc[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 2 * NCOLS);
}

template <bool SPLIT, bool FUSED, int RBT>
static int launch_rbt(const float* x, int B, int64_t F, float eps, ActQ q, float* y, float* partials, int64_t cap,
                  int* nparts, double* zero_acc, cudaStream_t s) {
  const int64_t ntiles = (F + NT - 1) / NT;
  int64_t grid = ntiles;
  if (grid > (int64_t)ALIGNQ_NUM_SMS * CTAS_PER_SM) grid = (int64_t)ALIGNQ_NUM_SMS * CTAS_PER_SM;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaError_t e = cudaFuncSetAttribute(gram_tc_small_kernel<SPLIT, FUSED, RBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  gram_tc_small_kernel<SPLIT, FUSED, RBT><<<(unsigned)grid, NT, SMEM_BYTES, s>>>(x, B, F, eps, q, y, partials, ntiles, zero_acc);
  ALIGNQ_LAUNCH_CHECK();
  *nparts = (int)grid;
  return ALIGNQ_OK;
}

template <bool SPLIT, bool FUSED>
static int launch(const float* x, int B, int64_t F, float eps, ActQ q, float* y, float* partials, int64_t cap,
                  int* nparts, double* zero_acc, cudaStream_t s) {
  if (B <= 8) return launch_rbt<SPLIT, FUSED, 8>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
  if (B <= 16) return launch_rbt<SPLIT, FUSED, 16>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
  if (B <= 24) return launch_rbt<SPLIT, FUSED, 24>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
  if (B <= 28) return launch_rbt<SPLIT, FUSED, 28>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
  return launch_rbt<SPLIT, FUSED, 32>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
}

}  // namespace tcs

// Partials [cta][nacc][B][B] like gram_tc.cu; the caller reduces them with the same kernel.
int gram_tc_small_partials(const float* x, int B, int64_t F, float eps, ActQ q, int fused, float* y, float* partials,
                           int64_t cap, int gram_mode, int* nparts, double* zero_acc, cudaStream_t s) {
  if (B > tcs::RB || B < 2) return ALIGNQ_ERANGE;
  if (gram_mode == ALIGNQ_GRAM_TF32X3)
    return fused ? tcs::launch<true, true>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s)
                 : tcs::launch<true, false>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
  if (gram_mode == ALIGNQ_GRAM_BF16)
    return fused ? tcs::launch<false, true>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s)
                 : tcs::launch<false, false>(x, B, F, eps, q, y, partials, cap, nparts, zero_acc, s);
  return ALIGNQ_EINVAL;
}

}  // namespace alignq
