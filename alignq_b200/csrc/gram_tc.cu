// tcgen05 (5th-gen tensor core) correlation / Gram kernels for sm_100a.
//
// Replaces corr() (cdf_alignment_admm/resnet-56-cifar-10/model/quantization.py:134-137;
// cdf_alignment_admm/dann_office/model/quantization.py:158-161) and, fused with the activation
// map, the forward of activation_quantize_fn with method == 'ours' (quantization.py:109-123):
// ONE read of x yields y (quantized activation), corr(x, x) and corr(t, t).
//
// Transform-then-MMA mainloop (operands need per-column batch statistics before the MMA, so they
// cannot come straight from TMA): per 32-column tile, all 512 threads
//   1. load x [B x 32] (coalesced 128 B rows), apply the CDF map (t) and write y,
//   2. reduce the column statistics of x and t across the 16 warps (pivot-shifted single pass),
//   3. standardise, convert and store the operands into shared memory in the canonical UMMA
//      K-major no-swizzle layout (8-row x 16-byte core matrices),
//   4. one elected thread issues tcgen05.mma (M = N = 128, A and B descriptors on the same tiles)
//      accumulating in TMEM; tcgen05.commit -> mbarrier releases the smem stage.
// Two smem stages overlap the MMAs of tile i with the ALU work of tile i+1.  Split-K over ONE wave of CTAs with
// equal tile counts; every CTA dumps its TMEM accumulators as fp32 partials, which gram_reduce_tc_kernel sums --
// or, for the fused forward, gram_finish_tc_kernel, which also evaluates ADMM.forward(D) and dL/dD in the same
// launch.  Batches of 2..32 rows take the thread-per-column kernels of gram_tc_small.cu instead.
//
// Numerics modes
//   ALIGNQ_GRAM_TF32X3: kind::tf32 with the operand split xs = H + L (H = top 19 bits):
//       G = H H^T + H L^T + L H^T  (the L L^T term, <= 2^-22 relative, is dropped): three MMAs per k-step.
//       The tensor core TRUNCATES when it adds into the fp32 accumulator (~4.6e-8 relative per accumulate), so
//       the small cross products have their own accumulator and both are flushed into the fp32 partials every
//       FLUSH_TILES tiles: fp32-level accuracy independent of F (tested to 1e-5 relative).
//   ALIGNQ_GRAM_BF16:   kind::f16 with bf16 operands, one MMA per k-step (tested to 1e-2 relative).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "tc_common.cuh"
#include "../../include/alignq_b200.h"

namespace cg = cooperative_groups;

namespace alignq {
int gram_tc_small_partials(const float* x, int B, int64_t F, float eps, ActQ q, int fused, float* y, float* partials,
                           int64_t cap, int gram_mode, int* nparts, double* zero_acc, cudaStream_t s);

namespace tc {

constexpr int KB = 32;            // feature columns per tile
constexpr int NT = 512;           // threads per CTA: 16 warps (4 per scheduler hide the erff / FMA chains)
constexpr int NW = NT / 32;       // warps
constexpr int RPT = 128 / NW;     // rows per thread: row r = warp + NW * i
constexpr int LBO = 144;          // bytes between K-adjacent core matrices (128 + 16 pad: conflict-free stores)

constexpr int RAW_TILE_BYTES = 128 * KB * (int)sizeof(float);   // one raw x tile [128 rows][32 cols] fp32

template <int MODE, bool FUSED, bool STAGED>
struct Cfg {
  static constexpr bool TF32 = (MODE == ALIGNQ_GRAM_TF32X3);
  static constexpr int ESZ = TF32 ? 4 : 2;                 // operand element bytes
  static constexpr int CH = 16 / ESZ;                      // elements per 16-byte core-matrix row
  static constexpr int NCH = KB / CH;                      // core matrices along K per tile
  static constexpr int SBO = NCH * LBO;                    // bytes between 8-row groups
  static constexpr int TILE_BYTES = 16 * SBO;              // 128 rows
  static constexpr int NSRC = FUSED ? 2 : 1;               // x [, t]
  static constexpr int NOPER = NSRC * (TF32 ? 2 : 1);      // operand tiles per stage (H, L per source)
  static constexpr int NACC = NSRC;                        // one fp32 accumulator [128 x 128] per source
  static constexpr int UMMA_K = 32 / ESZ;
  static constexpr int KSTEPS = KB / UMMA_K;
  // tf32x3: the small cross products (H L^T + L H^T) get their own accumulator, so only ONE large accumulate per
  // k-step hits the main accumulator (see FLUSH_TILES)
  static constexpr int ACC_COLS = TF32 ? 256 : 128;        // TMEM columns per source
  static constexpr int TMEM_COLS = NACC * ACC_COLS;        // 128 .. 512: powers of two
  // The tensor core TRUNCATES when it adds into the fp32 accumulator (measured: ~4.6e-8 relative drift per
  // accumulate, 3.5e-5 at B = 128, F = 262144 with one long chain).  tf32x3 therefore caps the chain: every
  // FLUSH_TILES tiles (64 large accumulates) the accumulators are added into the CTA's fp32 partials with
  // round-to-nearest and restarted.  bf16 mode (tolerance 1e-2) never flushes.
  static constexpr int FLUSH_TILES = TF32 ? 16 : 0;
  static constexpr int STAGE_BYTES = NOPER * TILE_BYTES;
  static constexpr int RED_BYTES = (NW + 1) * KB * 4 * (int)sizeof(float);   // per-warp partials + finished column stats
  // epilogue scratch: per-warp transpose tiles (global-partials path), or -- tf32 modes, cluster path -- this CTA's whole
  // accumulators [NACC][128][129] fp32, which the other CTAs of the cluster read through distributed shared memory
  static constexpr int CL_LD = 129;
  static constexpr int CL_BYTES = NACC * 128 * CL_LD * (int)sizeof(float);
  static constexpr bool CLUSTER_OK = TF32;                 // bf16 mode's operand stages are too small to host the tile
  static constexpr int EPI_BYTES = CLUSTER_OK ? CL_BYTES : NW * 32 * 33 * (int)sizeof(float);
  static constexpr int OPER_BYTES = (2 * STAGE_BYTES > EPI_BYTES) ? 2 * STAGE_BYTES : EPI_BYTES;
  // STAGED: ring of raw x tiles filled by cp.async (16 B, zero-filling) several tiles ahead, so that
  // >= 48 KB per SM are in flight (one register-prefetched tile is only 16 KB: latency-bound at ~1.3 TB/s)
  static constexpr int NRAW = STAGED ? ((TF32 && FUSED) ? 4 : 6) : 0;
  static constexpr int RAW_BYTES = NRAW * RAW_TILE_BYTES;
  static constexpr int SMEM_BYTES = OPER_BYTES + RAW_BYTES + RED_BYTES + 64;
  static constexpr uint32_t IDESC = make_idesc(TF32 ? 2u : 1u, 128u, 128u);
};

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MODE, bool FUSED, bool STAGED>
__global__ void __launch_bounds__(NT, 1)
gram_tc_kernel(const float* __restrict__ x, int B, int64_t F, float eps, ActQ q, float* __restrict__ y,
               float* __restrict__ partials, int64_t ntiles, double* __restrict__ zero_acc, int ncl) {
  pdl_trigger();                      // the finish kernel launches early and waits for this grid (common.cuh)
  using C = Cfg<MODE, FUSED, STAGED>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* stage_base = smem;
  uint8_t* raw_base = smem + C::OPER_BYTES;
  float4* red = reinterpret_cast<float4*>(smem + C::OPER_BYTES + C::RAW_BYTES);
  float4* colstat = red + NW * KB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OPER_BYTES + C::RAW_BYTES + C::RED_BYTES);   // [0,1] stage free, [2] accumulators done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (zero_acc && blockIdx.x == 0 && threadIdx.x < 4) zero_acc[threadIdx.x] = 0.0;   // arms gram_finish_tc_kernel (next in stream)

  // ---- one-time setup ------------------------------------------------------------------------
  for (int i = threadIdx.x; i < 2 * C::STAGE_BYTES / 16; i += NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
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
  // Prefetch.  STAGED: cp.async ring of raw tiles (needs 16-byte aligned rows: x aligned, F % 4 == 0);
  // otherwise the next tile's column of x (16 rows per thread) is prefetched into registers.
  float xn[RPT], pn = 0.f;
  auto fetch_regs = [&](int64_t tile) {
    const int64_t f = tile * KB + lane;
    const bool colv = tile < ntiles && f < F;
    pn = colv ? __ldg(x + f) : 0.f;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = warp + NW * i;
      xn[i] = (colv && r < B) ? __ldg(x + (int64_t)r * F + f) : 0.f;
    }
  };
  auto fetch_async = [&](int64_t tile, int slot) {          // 1024 x 16 B chunks per tile
    if (tile < ntiles) {
      const uint32_t dst0 = smem_u32(raw_base + slot * RAW_TILE_BYTES);
#pragma unroll
      for (int j = 0; j < 1024 / NT; ++j) {
        const int c = threadIdx.x + j * NT;
        const int r = c >> 3, c16 = c & 7;
        const int64_t f = tile * KB + c16 * 4;
        int64_t left = (F - f) * 4;                          // bytes of this row still inside the matrix
        left = left < 0 ? 0 : (left > 16 ? 16 : left);
        const uint32_t nbytes = (r < B) ? (uint32_t)left : 0u;
        const float* src = x + (nbytes ? (int64_t)r * F + f : 0);
        cp_async16_zfill(dst0 + r * (KB * 4) + c16 * 16, src, nbytes);
      }
    }
    cp_async_commit();                                       // (possibly empty) group keeps the counting uniform
  };
  if (STAGED) {
#pragma unroll 1
    for (int p = 0; p < C::NRAW - 1; ++p) fetch_async(blockIdx.x + (int64_t)p * gridDim.x, p);
    cp_async_wait<C::NRAW - 2>();                            // tile 0 has landed (this thread's chunks)
    __syncthreads();                                         // ... and everyone else's
  } else {
    fetch_regs(blockIdx.x);
  }
  // Accumulators -> this CTA's fp32 partials (add = false: store, true: read-modify-write of the CTA's own slot).
  // TMEM (lane = row, 32 consecutive columns per thread) -> per-warp smem transpose -> 128-byte row stores.
  int nflush = 0;
  auto dump = [&](bool add) {
    const int qd = warp & 3, cb = warp >> 2;                       // lane quarter, 32-column block
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);       // the operand stages are idle when this runs
    float* out = partials + (size_t)blockIdx.x * C::NACC * B * B;
#pragma unroll 1
    for (int a = 0; a < C::NACC; ++a) {
#pragma unroll 1
      for (int col0 = cb * 32; col0 < 128; col0 += (NW / 4) * 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + a * C::ACC_COLS + col0, v);
        if (C::TF32) {
          uint32_t w[32];
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + a * C::ACC_COLS + 128 + col0, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
        }
        __syncwarp();
        if (col0 + lane < B) {
          float* o = out + ((size_t)a * B + qd * 32) * B + col0 + lane;
          const int nrow = (B - qd * 32) < 32 ? (B - qd * 32) : 32;
          float prev[32];                                          // all loads in flight before the first store
#pragma unroll
          for (int r = 0; r < 32; ++r) prev[r] = (add && r < nrow) ? __ldcg(o + (size_t)r * B) : 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r)
            if (r < nrow) o[(size_t)r * B] = prev[r] + tr[r * 33 + lane];
        }
        __syncwarp();
      }
    }
  };
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    uint8_t* st = stage_base + s * C::STAGE_BYTES;
    if (C::FLUSH_TILES > 0 && it > 0 && (it % (C::FLUSH_TILES > 0 ? C::FLUSH_TILES : 1)) == 0) {
      // every MMA issued so far has to retire before the accumulators are read; the operand stages are idle then,
      // so the transpose scratch may use them, and the two barriers of step 2 separate it from the next operand stores
      if (threadIdx.x == 0) umma_commit(&bars[2]);
      mbar_wait(&bars[2], (uint32_t)(nflush & 1));
      tc_fence_after();
      dump(nflush > 0);
      ++nflush;
      tc_fence_before();
    }

    // ---- 1. take the prefetched column, start the next fetch, map + quantise ---------------------
    const int64_t f = tile * KB + lane;
    const bool colv = f < F;
    float xv[RPT], tv[RPT], px;                                  // px = pivot: row 0 of this column
    if (STAGED) {
      // refill the slot consumed in the previous iteration (everyone passed that iteration's last barrier)
      fetch_async(tile + (int64_t)(C::NRAW - 1) * gridDim.x, (it + C::NRAW - 1) % C::NRAW);
      const float* raw = reinterpret_cast<const float*>(raw_base + (it % C::NRAW) * RAW_TILE_BYTES);
#pragma unroll
      for (int i = 0; i < RPT; ++i) xv[i] = raw[(warp + NW * i) * KB + lane];
      px = raw[lane];
    } else {
      px = pn;
#pragma unroll
      for (int i = 0; i < RPT; ++i) xv[i] = xn[i];
      fetch_regs(tile + gridDim.x);
    }
    const float pt = FUSED ? act_map_t(px, q.ar) : 0.f;
    float s1 = 0.f, s2 = 0.f, u1 = 0.f, u2 = 0.f;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = warp + NW * i;
      const bool v = colv && r < B;
      const float d = xv[i] - px;
      s1 += v ? d : 0.f;
      s2 = v ? fmaf(d, d, s2) : s2;
      if (FUSED) {
        tv[i] = act_map_t(xv[i], q.ar);
        if (v && y) y[(int64_t)r * F + f] = act_quant_from_t(tv[i], q);
        const float e = tv[i] - pt;
        u1 += v ? e : 0.f;
        u2 = v ? fmaf(e, e, u2) : u2;
      }
    }
    // ---- 2. column statistics across the warps: partials -> smem, warp 0 finishes the 32 columns ----
    red[warp * KB + lane] = make_float4(s1, s2, u1, u2);
    __syncthreads();
    if (warp == 0) {
      float S1 = 0.f, S2 = 0.f, U1 = 0.f, U2 = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) { const float4 p = red[w * KB + lane]; S1 += p.x; S2 += p.y; U1 += p.z; U2 += p.w; }
      float vx = (S2 - S1 * S1 * invB) * invBm1;
      vx = (vx < 0.f) ? 0.f : vx;
      float4 o;
      o.x = px + S1 * invB;                       // mean of x
      o.y = 1.0f / (sqrtf(vx) + eps);             // 1 / (std + eps)
      o.z = 0.f; o.w = 0.f;
      if (FUSED) {
        float vt = (U2 - U1 * U1 * invB) * invBm1;
        vt = (vt < 0.f) ? 0.f : vt;
        o.z = pt + U1 * invB;
        o.w = 1.0f / (sqrtf(vt) + eps);
      }
      colstat[lane] = o;
    }
    __syncthreads();
    const float4 cs = colstat[lane];
    const float mx = cs.x, rx = cs.y, mt = cs.z, rt = cs.w;
    // ---- 3. standardise, convert, store operands (row r = warp + 8 i -> row group i, row-in-group warp) ----
    if (it >= 2) mbar_wait(&bars[s], ((it >> 1) - 1) & 1);       // MMAs that read this stage have retired
    const int chunk = lane / C::CH, within = lane % C::CH;
    uint8_t* dst0 = st + chunk * LBO + within * C::ESZ;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = warp + NW * i;                               // row group r / 8, row in group r % 8
      if (r >= B) break;
      uint8_t* dst = dst0 + (r >> 3) * C::SBO + (r & 7) * 16;
      const float a = colv ? (xv[i] - mx) * rx : 0.f;
      const float b = (FUSED && colv) ? (tv[i] - mt) * rt : 0.f;
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
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    if (STAGED) cp_async_wait<(C::NRAW >= 2 ? C::NRAW - 2 : 0)>();     // next raw tile has landed (this thread's chunks)
    __syncthreads();
    // ---- 4. MMAs: one thread issues for the whole CTA --------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sb = smem_u32(st);
#pragma unroll
      for (int ks = 0; ks < C::KSTEPS; ++ks) {
        const bool fresh = (C::FLUSH_TILES > 0) ? (it % (C::FLUSH_TILES > 0 ? C::FLUSH_TILES : 1)) == 0 : it == 0;
        const uint32_t acc = (!fresh || ks > 0) ? 1u : 0u;
        const uint32_t koff = ks * 2 * LBO;
#pragma unroll
        for (int src = 0; src < C::NSRC; ++src) {
          if (C::TF32) {
            const uint64_t dh = make_desc(sb + (2 * src) * C::TILE_BYTES + koff, LBO, C::SBO);
            const uint64_t dl = make_desc(sb + (2 * src + 1) * C::TILE_BYTES + koff, LBO, C::SBO);
            umma<true>(tmem_base + src * C::ACC_COLS, dh, dh, C::IDESC, acc);            // H H^T -> main
            umma<true>(tmem_base + src * C::ACC_COLS + 128, dh, dl, C::IDESC, acc);      // H L^T -> cross
            umma<true>(tmem_base + src * C::ACC_COLS + 128, dl, dh, C::IDESC, 1u);       // L H^T -> cross
          } else {
            const uint64_t d = make_desc(sb + src * C::TILE_BYTES + koff, LBO, C::SBO);
            umma<false>(tmem_base + src * C::ACC_COLS, d, d, C::IDESC, acc);
          }
        }
      }
      umma_commit(&bars[s]);        // implies tcgen05.fence::before_thread_sync; frees the stage when the MMAs retire
    }
  }
  // ---- epilogue: TMEM -> registers -> fp32 partials -----------------------------------------------
  if (threadIdx.x == 0) umma_commit(&bars[2]);
  mbar_wait(&bars[2], (uint32_t)(nflush & 1));
  tc_fence_after();
  if (C::CLUSTER_OK && ncl > 1) {
    // Split-K reduction INSIDE the cluster (distributed shared memory) instead of through HBM: every CTA parks its
    // accumulators in its own shared memory, then CTA r of the cluster adds rows [128 r / ncl, 128 (r+1) / ncl) of all
    // ncl CTAs and writes that slice of ONE partial set per cluster.  With clusters of 8 the partials written (and
    // re-read by the finish kernel) drop from 128 sets = 16.8 MB to 16 sets = 2.1 MB per layer.  (Only used when the
    // CTA never flushed mid-way: nflush == 0.)
    cg::cluster_group cluster = cg::this_cluster();
    float* mine = reinterpret_cast<float*>(smem);
    {
      const int qd = warp & 3, cb = warp >> 2;
#pragma unroll 1
      for (int a = 0; a < C::NACC; ++a) {
#pragma unroll 1
        for (int col0 = cb * 32; col0 < 128; col0 += (NW / 4) * 32) {
          uint32_t v[32], w[32];
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + a * C::ACC_COLS + col0, v);
          tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + a * C::ACC_COLS + 128 + col0, w);
          float* o = mine + ((size_t)a * 128 + qd * 32 + lane) * C::CL_LD + col0;
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
        }
      }
    }
    tc_fence_before();
    cluster.sync();                                        // every CTA's tile is complete (and every CTA is resident)
    const int rank = (int)cluster.block_rank();
    const int rows = 128 / ncl;
    float* out = partials + (size_t)(blockIdx.x / ncl) * C::NACC * B * B;
    for (int e = threadIdx.x; e < C::NACC * rows * 128; e += NT) {
      const int col = e & 127, rr = (e >> 7) % rows, a = e / (rows * 128);
      const int row = rank * rows + rr;
      if (row >= B || col >= B) continue;
      const size_t off = ((size_t)a * 128 + row) * C::CL_LD + col;
      float part[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) part[r] = (r < ncl) ? *cluster.map_shared_rank(mine + off, r) : 0.f;   // all loads in flight
      float sum = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) sum += part[r];                                        // fixed order: deterministic
      out[((size_t)a * B + row) * B + col] = sum;
    }
    cluster.sync();                                        // nobody leaves while its shared memory is still being read
  } else {
    dump(nflush > 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// G = (1/F) sum_cta acc;  fused: D = G_t - G_x (two separately rounded Grams, as quantization.py:118-122).
// Block (32 elements, 32 part lanes): lane pl sums parts pl, pl + 32, ... of its element (coalesced 128-byte rows),
// then the 32 lane sums are added in a fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(1024)
gram_reduce_tc_kernel(const float* __restrict__ partials, int nparts, int B, float invF, int nacc, int fused,
                      float* __restrict__ G, float* __restrict__ D) {
  __shared__ float sm[2][32][33];
  const size_t bb = (size_t)B * B;
  const size_t e = (size_t)blockIdx.x * 32 + threadIdx.x;
  const int pl = threadIdx.y;
  for (int a = 0; a < nacc; ++a) {
    float acc = 0.f;
    if (e < bb)
      for (int p = pl; p < nparts; p += 32) acc += partials[((size_t)p * nacc + a) * bb + e];
    sm[a][pl][threadIdx.x] = acc;
  }
  __syncthreads();
  if (pl != 0 || e >= bb) return;
  float gx = 0.f, gt = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) gx += sm[0][k][threadIdx.x];
  gx = __fmul_rn(gx, invF);
  if (G) G[e] = gx;
  if (fused) {
#pragma unroll
    for (int k = 0; k < 32; ++k) gt += sm[1][k][threadIdx.x];
    D[e] = (fused == 2) ? __fmul_rn(gt, invF) : __fsub_rn(__fmul_rn(gt, invF), gx);      // 2: the t sum itself
  }
}

static void launch_reduce_tc(const float* partials, int nparts, int B, int64_t F, int nacc, int fused, float* G, float* D,
                             cudaStream_t s) {
  const int bb = B * B;
  gram_reduce_tc_kernel<<<(bb + 31) / 32, dim3(32, 32), 0, s>>>(partials, nparts, B, 1.0f / (float)F, nacc, fused, G, D);
}

// Fused split-K reduction + ADMM.forward(D) (cdf_alignment_admm/resnet-56-cifar-10/utils/admm.py:24-33) + dL/dD for
// the fused forward: same block shape as gram_reduce_tc_kernel, then every block adds its three partial sums
// (sum|Z|, sum R^2, sum U|R|, fp64) to acc[0..2] and takes a ticket; the LAST block finishes loss and dL/dD.
// acc[0..3] (3 sums + the ticket) sit at the start of the workspace and are zeroed by the Gram kernel that runs just
// before.  One launch instead of reduce + an 8-CTA cluster launch; arithmetic identical to admm_loss_kernel.
__global__ void __launch_bounds__(1024)
gram_finish_tc_kernel(const float* __restrict__ partials, int nparts, int B, float invF, AdmmFinish f,
                      float* __restrict__ D, double* __restrict__ acc) {
  pdl_wait();                         // programmatic dependent of the Gram kernel: nothing is touched before this

  // Round 2: ONE element per thread (1024 elements per block, 16 blocks at B = 128), every thread walks the (few, after the
  // in-cluster reduction) partial sets of its element with coalesced loads in a fixed order, then a block reduction and
  // three fp64 atomics per BLOCK.  The first version used 32 elements x 32 part lanes per block: 512 blocks, 1536 same-
  // address fp64 atomics (they serialise in L2) -- 19 us for 4 MB of partials.
  __shared__ double red[3][32];
  __shared__ unsigned last_flag;
  const int bb = B * B;
  const int e = blockIdx.x * 1024 + threadIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  if (e < bb) {
    float gx = 0.f, gt = 0.f;
#pragma unroll 4
    for (int p = 0; p < nparts; ++p) {
      gx += partials[((size_t)p * 2 + 0) * bb + e];
      gt += partials[((size_t)p * 2 + 1) * bb + e];
    }
    gx = __fmul_rn(gx, invF);
    const float d = __fsub_rn(__fmul_rn(gt, invF), gx);
    D[e] = d;
    const int i = e / B, j = e - i * B;
    const float z = f.Z[(size_t)i * f.dim + j], u = f.U[(size_t)i * f.dim + j];
    const float r = __fsub_rn(d, z);
    s0 = (double)fabsf(z);
    s1 = (double)__fmul_rn(r, r);
    s2 = (double)__fmul_rn(u, fabsf(r));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    s0 = red[0][lane]; s1 = red[1][lane]; s2 = red[2][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { atomicAdd(acc + 0, s0); atomicAdd(acc + 1, s1); atomicAdd(acc + 2, s2); }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(acc + 3), 1u);
    last_flag = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!last_flag) return;
  __threadfence();
  const double S0 = __ldcg(acc + 0), S1 = __ldcg(acc + 1), S2 = __ldcg(acc + 2);
  const float inv_bb = 1.0f / (float)bb;
  const float rms = sqrtf((float)(S1 / bb));                  // mean(...) ** 0.5
  const int tid = threadIdx.x;
  if (tid == 0 && f.loss) {
    const float reg = f.mu * (float)(S0 / bb);
    const float con = (f.rho / 2.0f) * rms;
    *f.loss = (reg + con) + (float)(S2 / bb);
  }
  if (f.dLdD) {
    const float k1 = (f.rho / 2.0f) / rms * inv_bb;           // rms == 0 -> inf, and 0 * inf = NaN as autograd gives
    for (int q = tid; q < bb; q += 1024) {
      const int i = q / B, j = q - i * B;
      const float z = f.Z[(size_t)i * f.dim + j], u = f.U[(size_t)i * f.dim + j];
      const float r = __fsub_rn(__ldcg(D + q), z);
      const float sg = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
      f.dLdD[q] = k1 * r + u * sg * inv_bb;
    }
  }
}

// The same for MANY partial sets (the small-batch kernels write up to 3 per SM): 32 elements x 32 part lanes per block.
__global__ void __launch_bounds__(1024)
gram_finish_tc_wide_kernel(const float* __restrict__ partials, int nparts, int B, float invF, AdmmFinish f,
                      float* __restrict__ D, double* __restrict__ acc) {
  pdl_wait();                         // programmatic dependent of the Gram kernel: nothing is touched before this

  __shared__ float sm[2][32][33];
  __shared__ unsigned last_flag;
  const int bb = B * B;
  const int e = blockIdx.x * 32 + threadIdx.x;
  const int pl = threadIdx.y;
  for (int a = 0; a < 2; ++a) {
    float v = 0.f;
    if (e < bb)
      for (int p = pl; p < nparts; p += 32) v += partials[((size_t)p * 2 + a) * bb + e];
    sm[a][pl][threadIdx.x] = v;
  }
  __syncthreads();
  if (pl == 0) {                                               // warp 0: one element per lane
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (e < bb) {
      float gx = 0.f, gt = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) { gx += sm[0][k][threadIdx.x]; gt += sm[1][k][threadIdx.x]; }
      gx = __fmul_rn(gx, invF);
      const float d = __fsub_rn(__fmul_rn(gt, invF), gx);
      D[e] = d;
      const int i = e / B, j = e - i * B;
      const float z = f.Z[(size_t)i * f.dim + j], u = f.U[(size_t)i * f.dim + j];
      const float r = __fsub_rn(d, z);
      s0 = (double)fabsf(z);
      s1 = (double)__fmul_rn(r, r);
      s2 = (double)__fmul_rn(u, fabsf(r));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (threadIdx.x == 0) { atomicAdd(acc + 0, s0); atomicAdd(acc + 1, s1); atomicAdd(acc + 2, s2); }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && pl == 0) {
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(acc + 3), 1u);
    last_flag = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!last_flag) return;
  __threadfence();
  const double S0 = __ldcg(acc + 0), S1 = __ldcg(acc + 1), S2 = __ldcg(acc + 2);
  const float inv_bb = 1.0f / (float)bb;
  const float rms = sqrtf((float)(S1 / bb));                  // mean(...) ** 0.5
  const int tid = pl * 32 + threadIdx.x;
  if (tid == 0 && f.loss) {
    const float reg = f.mu * (float)(S0 / bb);
    const float con = (f.rho / 2.0f) * rms;
    *f.loss = (reg + con) + (float)(S2 / bb);
  }
  if (f.dLdD) {
    const float k1 = (f.rho / 2.0f) / rms * inv_bb;           // rms == 0 -> inf, and 0 * inf = NaN as autograd gives
    for (int q = tid; q < bb; q += 1024) {
      const int i = q / B, j = q - i * B;
      const float z = f.Z[(size_t)i * f.dim + j], u = f.U[(size_t)i * f.dim + j];
      const float r = __fsub_rn(__ldcg(D + q), z);
      const float sg = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
      f.dLdD[q] = k1 * r + u * sg * inv_bb;
    }
  }
}

// reduce, or reduce + ADMM loss when the caller asked for it (fused forward only)
static void launch_finish_tc(const float* partials, int nparts, int B, int64_t F, int nacc, int fused, float* G, float* D,
                             const AdmmFinish* fin, void* ws, cudaStream_t s) {
  if (fused && G && D && !fin) {              // raw sums (gram_tc_sums): G <- sum_x, D <- sum_t, not divided by F
    const int bb = B * B;
    gram_reduce_tc_kernel<<<(bb + 31) / 32, dim3(32, 32), 0, s>>>(partials, nparts, B, 1.0f, nacc, 2, G, D);
    return;
  }
  if (fin && fused) {
    const int bb = B * B;
    if (nparts <= 64)
      (void)launch_pdl(gram_finish_tc_kernel, dim3((bb + 1023) / 1024), dim3(1024), 0, s, partials, nparts, B, 1.0f / (float)F, *fin,
                       D, reinterpret_cast<double*>(ws));
    else
      (void)launch_pdl(gram_finish_tc_wide_kernel, dim3((bb + 31) / 32), dim3(32, 32), 0, s, partials, nparts, B, 1.0f / (float)F,
                       *fin, D, reinterpret_cast<double*>(ws));
  } else {
    launch_reduce_tc(partials, nparts, B, F, nacc, fused, G, D, s);
  }
}

template <int MODE, bool FUSED, bool STAGED>
static int launch_impl(const float* x, int B, int64_t F, float eps, ActQ q, float* y, float* G, float* D, void* ws,
                       size_t ws_bytes, const AdmmFinish* fin, cudaStream_t s) {
  using C = Cfg<MODE, FUSED, STAGED>;
  const int64_t ntiles = (F + KB - 1) / KB;
  float* partials = reinterpret_cast<float*>(ws) + gram_wsym_floats(B);
  const size_t head = gram_wsym_floats(B) * sizeof(float);
  if (ws_bytes <= head) return ALIGNQ_ENOSPACE;
  int64_t cap = (int64_t)((ws_bytes - head) / ((size_t)C::NACC * B * B * sizeof(float)));
  if (cap < 1) return ALIGNQ_ENOSPACE;
  // One wave of CTAs with an equal number of tiles each: tiles-per-CTA = ceil(ntiles / 148).  (Measured at B = 128:
  // F = 4096 / 8192 / 16384 are all fastest with 128 CTAs of 1 / 2 / 4 tiles; fewer, longer CTAs lose more in the
  // main loop than they save in partials.)
  const int64_t tiles_per_cta = (ntiles + ALIGNQ_NUM_SMS - 1) / ALIGNQ_NUM_SMS;
  int64_t grid = (ntiles + tiles_per_cta - 1) / tiles_per_cta;
  if (grid > ALIGNQ_NUM_SMS) grid = ALIGNQ_NUM_SMS;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  cudaError_t e = cudaFuncSetAttribute(gram_tc_kernel<MODE, FUSED, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  if (e != cudaSuccess) return (int)e;
  double* zero_acc = (fin && FUSED) ? reinterpret_cast<double*>(ws) : nullptr;
  // In-cluster split-K reduction (tf32 modes): clusters of 8 CTAs when no CTA flushes mid-way and the grid divides.
  int ncl = 1;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (C::CLUSTER_OK && grid >= 16 && (C::FLUSH_TILES == 0 || (ntiles + grid - 1) / grid <= C::FLUSH_TILES)) {
    // the largest cluster size whose clusters are ALL co-resident (one wave): a cluster needs its SMs inside one GPC,
    // and with ~220 KB of shared memory per CTA (one CTA per SM) a second wave would double the kernel's time
    static int best_ncl[2] = {0, 0};                              // per FUSED instantiation (smem differs); benign race
    int& cached = best_ncl[FUSED ? 1 : 0];
    if (cached == 0) {
      cached = 1;
      for (int cand = 8; cand >= 2; cand >>= 1) {
        cfg.gridDim = dim3((unsigned)(ALIGNQ_NUM_SMS / cand * cand));
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cand; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, gram_tc_kernel<MODE, FUSED, STAGED>, &cfg) == cudaSuccess &&
            nclusters * cand >= 128) { cached = cand; break; }
        (void)cudaGetLastError();
      }
    }
    ncl = cached;
    grid -= grid % ncl;                                           // a multiple of the cluster size; tiles are strided over the grid
  }
  int nparts = (int)grid;
  if (ncl > 1) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ncl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, gram_tc_kernel<MODE, FUSED, STAGED>, x, B, F, eps, q, y, partials, ntiles, zero_acc, ncl);
    if (e != cudaSuccess) {                                       // clusters not schedulable here: the global-partials path
      (void)cudaGetLastError();
      ncl = 1;
    } else {
      nparts = (int)(grid / ncl);
    }
  }
  if (ncl == 1) gram_tc_kernel<MODE, FUSED, STAGED><<<(unsigned)grid, NT, C::SMEM_BYTES, s>>>(x, B, F, eps, q, y, partials, ntiles, zero_acc, 1);
  ALIGNQ_LAUNCH_CHECK();
  launch_finish_tc(partials, nparts, B, F, C::NACC, FUSED ? 1 : 0, G, D, fin, ws, s);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

template <int MODE, bool FUSED>
static int launch(const float* x, int B, int64_t F, float eps, ActQ q, float* y, float* G, float* D, void* ws,
                  size_t ws_bytes, const AdmmFinish* fin, cudaStream_t s) {
  const bool staged = aligned16(x) && (F % 4 == 0);          // cp.async needs 16-byte aligned row segments
  if (B <= 32) {                                             // gram_tc_small.cu: thread-per-column variant
    constexpr int NACC = FUSED ? 2 : 1;
    float* partials = reinterpret_cast<float*>(ws) + gram_wsym_floats(B);
    const size_t head = gram_wsym_floats(B) * sizeof(float);
    if (ws_bytes <= head) return ALIGNQ_ENOSPACE;
    const int64_t cap = (int64_t)((ws_bytes - head) / ((size_t)NACC * B * B * sizeof(float)));
    if (cap < 1) return ALIGNQ_ENOSPACE;
    int nparts = 0;
    double* zero_acc = (fin && FUSED) ? reinterpret_cast<double*>(ws) : nullptr;
    const int rc = gram_tc_small_partials(x, B, F, eps, q, FUSED ? 1 : 0, y, partials, cap, MODE, &nparts, zero_acc, s);
    if (rc != ALIGNQ_OK) return rc;
    launch_finish_tc(partials, nparts, B, F, NACC, FUSED ? 1 : 0, G, D, fin, ws, s);
    ALIGNQ_LAUNCH_CHECK();
    return ALIGNQ_OK;
  }
  return staged ? launch_impl<MODE, FUSED, true>(x, B, F, eps, q, y, G, D, ws, ws_bytes, fin, s)
                : launch_impl<MODE, FUSED, false>(x, B, F, eps, q, y, G, D, ws, ws_bytes, fin, s);
}

}  // namespace tc

int gram_tc_corr(const float* x, int B, int64_t F, float eps, float* G, void* ws, size_t ws_bytes, int gram_mode,
                 cudaStream_t s) {
  if (B > 128 || B < 2) return ALIGNQ_ERANGE;      // one 128-row UMMA tile; larger batches use the fp32 path
  ActQ q{0.f, 0.f, 0.f, 0};
  if (gram_mode == ALIGNQ_GRAM_TF32X3) return tc::launch<ALIGNQ_GRAM_TF32X3, false>(x, B, F, eps, q, nullptr, G, nullptr, ws, ws_bytes, nullptr, s);
  if (gram_mode == ALIGNQ_GRAM_BF16) return tc::launch<ALIGNQ_GRAM_BF16, false>(x, B, F, eps, q, nullptr, G, nullptr, ws, ws_bytes, nullptr, s);
  return ALIGNQ_EINVAL;
}

int gram_tc_sums(const float* x, int B, int64_t F, ActQ q, float eps, float* Gx, float* Gt, void* ws, size_t ws_bytes,
                 int gram_mode, cudaStream_t s) {
  if (B > 128 || B < 2 || !Gx || !Gt) return ALIGNQ_ERANGE;
  if (gram_mode == ALIGNQ_GRAM_TF32X3) return tc::launch<ALIGNQ_GRAM_TF32X3, true>(x, B, F, eps, q, nullptr, Gx, Gt, ws, ws_bytes, nullptr, s);
  if (gram_mode == ALIGNQ_GRAM_BF16) return tc::launch<ALIGNQ_GRAM_BF16, true>(x, B, F, eps, q, nullptr, Gx, Gt, ws, ws_bytes, nullptr, s);
  return ALIGNQ_EINVAL;
}

int gram_tc_fused_fwd(const float* x, int B, int64_t F, ActQ q, float eps, float* y, float* D, void* ws, size_t ws_bytes,
                      int gram_mode, const AdmmFinish* fin, cudaStream_t s) {
  if (B > 128 || B < 2) return ALIGNQ_ERANGE;
  if (gram_mode == ALIGNQ_GRAM_TF32X3) return tc::launch<ALIGNQ_GRAM_TF32X3, true>(x, B, F, eps, q, y, nullptr, D, ws, ws_bytes, fin, s);
  if (gram_mode == ALIGNQ_GRAM_BF16) return tc::launch<ALIGNQ_GRAM_BF16, true>(x, B, F, eps, q, y, nullptr, D, ws, ws_bytes, fin, s);
  return ALIGNQ_EINVAL;
}

}  // namespace alignq
