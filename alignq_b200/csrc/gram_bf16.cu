// bf16 Gram / SYRK on the 5th-gen tensor cores: G = X X^T (/ F) for X [B <= 256, F] bf16 row-major.
//
// This is the dense contraction at the heart of corr() (cdf_alignment_admm/resnet-56-cifar-10/model/
// quantization.py:134-137: torch.matmul(x_std, x_std^T) / F) for operands that are ALREADY
// standardised and stored in bf16 -- the tensor-bound micro-shape of SURVEY.md 8(d)
// (B = 256, F = 2^20: 137 GFLOP over 537 MB, arithmetic intensity 256 flop/B = the bf16 ridge).
// The fp32 training path (gram_tc.cu) has to transform its operands first and is ALU/HBM-bound.
//
// Structure (one CTA per SM, split-K over the feature dimension, 192 threads):
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

// Sum of the per-CTA partials.  32 consecutive elements per block; the block's 4 warps each sum a
// quarter of the parts (coalesced 128-byte rows, 4 loads in flight per thread) and warp 0 combines the
// four sums in fixed order (deterministic).  Elements of the never-computed lower-left block
// (i >= 128, j < 128) are skipped here and mirrored from the upper-right block afterwards.
__global__ void __launch_bounds__(128)
gram_bf16_reduce_kernel(const float* __restrict__ partials, int nparts, int B, float scale, float* __restrict__ G) {
  __shared__ float part[4][32];
  const size_t bb = (size_t)B * B;
  const size_t e = (size_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  const int i = (int)(e / B), j = (int)(e - (size_t)i * B);
  const bool live = e < bb && !(i >= 128 && j < 128);
  float acc = 0.f;
  if (live) {
#pragma unroll 4
    for (int p = w; p < nparts; p += 4) acc += partials[(size_t)p * bb + e];
  }
  part[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && live) G[e] = (((part[0][threadIdx.x] + part[1][threadIdx.x]) + part[2][threadIdx.x]) + part[3][threadIdx.x]) * scale;
}

// G[i][j] = G[j][i] for i >= 128, j < 128 (32 x 32 tiles through shared memory, coalesced both ways)
__global__ void __launch_bounds__(256)
gram_bf16_mirror_kernel(float* __restrict__ G, int B) {
  __shared__ float tile[32][33];
  const int ti = 4 + blockIdx.y, tj = blockIdx.x;                        // destination tile (rows >= 128, cols < 128)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = tj * 32 + ty + 8 * r, col = ti * 32 + tx;            // source element
    tile[ty + 8 * r][tx] = (row < B && col < B) ? G[(size_t)row * B + col] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = ti * 32 + ty + 8 * r, col = tj * 32 + tx;
    if (row < B && col < B) G[(size_t)row * B + col] = tile[tx][ty + 8 * r];
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
  int64_t grid = ktiles < ALIGNQ_NUM_SMS ? ktiles : ALIGNQ_NUM_SMS;
  const int64_t cap = (int64_t)(ws_bytes / ((size_t)B * B * sizeof(float)));
  if (cap < 1) return ALIGNQ_ENOSPACE;
  if (grid > cap) grid = cap;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaFuncSetAttribute(gram_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  gram_bf16_kernel<<<(unsigned)grid, NTHREADS, SMEM_BYTES, s>>>(tmap, B, ktiles, reinterpret_cast<float*>(ws));
  ALIGNQ_LAUNCH_CHECK();
  const size_t bb = (size_t)B * B;
  gram_bf16_reduce_kernel<<<(unsigned)((bb + 31) / 32), 128, 0, s>>>(reinterpret_cast<float*>(ws), (int)grid, B,
                                                                     divide_by_F ? 1.0f / (float)F : 1.0f, G);
  if (B > 128) {
    ALIGNQ_LAUNCH_CHECK();
    gram_bf16_mirror_kernel<<<dim3(4, (B - 128 + 31) / 32), 256, 0, s>>>(G, B);
  }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
