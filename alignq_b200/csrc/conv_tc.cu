// tcgen05 (5th-gen tensor core) 3x3 convolutions for the quantized conv layers -- SURVEY.md 8(f) item 2:
//   F.conv2d(input, weight_q, None, stride 1, padding 1)       cdf_alignment/resnet-20-cifar-10/model/quantization.py:116-120
// forward, data gradient and weight gradient for NHWC (channels_last) fp32 tensors with Cin == Cout == C in {16, 32, 64},
// the shape of 16 of the 21 convolutions of resnet20_quant (and 48 of 57 of resnet56_quant).  Everything else (stem,
// strided and 1x1 projection convs, depthwise) stays on cuDNN.
//
// Implicit GEMM WITHOUT im2col.  Pixels live in a zero-padded linear space: image n occupies Hp x Wp = (H+2) x (W+2)
// positions, global position g = (n Hp + yy) Wp + xx.  Output position g (yy < H, xx < W: output pixel (n, yy, xx)) reads
// input positions g + kh Wp + kw, kh, kw in 0..2 -- a pure SHIFT.  A tile of positions is staged in shared memory as
// "planes": plane c4 holds channels 4 c4 .. 4 c4 + 3 of every position, 16 bytes per position, positions consecutive:
//       byte address = plane_base + c4 * PS + position * 16.
// That one layout is at the same time
//   * a K-major UMMA operand (rows = positions, K = channels: 8 rows x 16 B core matrices, SBO = 128 B so rows are
//     LINEAR in the position, LBO = PS), used by the forward / data-gradient GEMM  D[pos, cout] = sum_tap A_tap W_tap, and
//   * an MN-major UMMA operand (MN = channels, K = positions: 8 K-rows x 16 B core matrices, LBO = 128 B so K is LINEAR
//     in the position, SBO = PS), used by the weight-gradient GEMM  dW_tap[cout, cin] = sum_pos gy[pos, cout] x[pos + shift, cin];
// and because the position enters linearly, tap (kh, kw) is just the descriptor start address moved by (kh Wp + kw) * 16
// bytes: shifted views of ONE staged tile, no data duplication (the forward uses the three kh shifts this way and folds the
// kw taps into the MMA's N, see FwdCfg; the weight gradient stages an im2col of the x operand instead, see below).
// Pad positions (xx >= W or yy >= H) produce junk output rows that are never stored (forward) or carry zero gy (weight
// gradient).
//
// Around the GEMM: the kernels are launched as programmatic dependents of the kernel before them (common.cuh: pdl_wait);
// the forward epilogue can reduce the batch statistics of the BatchNorm that follows (STATS), the data-gradient epilogue
// the backward sums of the bn-act layer that precedes (BNRED, bn_stat.cuh) -- one launch fewer per layer either way.
//
// Numerics (mode): ALIGNQ_CONV_TF32 = one kind::tf32 MMA per k-step on operands rounded to tf32 (what cuDNN runs by
// default under torch.backends.cudnn.allow_tf32 = True, the reference's own GPU path; ~3e-4 of max|ref|; the weight
// gradient: one kind::f16 MMA on fp16 operands -- the same 11 significand bits -- with gy scaled per CTA); ALIGNQ_CONV_TF32X3 =
// operands split v = H + L (H = top 19 bits), three MMAs (H H into the main accumulator, H L + L H into a second one:
// the tensor core truncates when it adds into the fp32 accumulator), fp32-level parity with F.conv2d (tested to 1e-5).
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_small_common.cuh"
#include "bn_stat.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace ctc {

using namespace tc;

constexpr int NT = 256;                 // threads per CTA: 8 warps (warps w and w+4 share TMEM lane quarter w)

// Dev aid (tools/conv_trace.py builds a second library with -DALIGNQ_CONV_TRACE): thread 0 of every CTA of the forward /
// data-gradient kernel stamps %globaltimer at the phase boundaries; alignq_conv_trace_read copies the stamps out.
#ifdef ALIGNQ_CONV_TRACE
constexpr int TRACE_SLOTS = 16, TRACE_CTAS = 1024;
__device__ unsigned long long g_conv_trace[TRACE_CTAS * TRACE_SLOTS];
__device__ __forceinline__ void trace_stamp(int slot) {
  if (threadIdx.x == 0 && blockIdx.x < TRACE_CTAS && slot < TRACE_SLOTS) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    g_conv_trace[blockIdx.x * TRACE_SLOTS + slot] = t;
  }
}
__device__ __forceinline__ void trace_smid() {
  if (threadIdx.x == 0 && blockIdx.x < TRACE_CTAS) {
    unsigned id;
    asm volatile("mov.u32 %0, %smid;" : "=r"(id));
    g_conv_trace[blockIdx.x * TRACE_SLOTS + TRACE_SLOTS - 1] = id;
  }
}
#define TRACE(slot) trace_stamp(slot)
#else
#define TRACE(slot) ((void)0)
#endif

struct Geo {
  int N, H, W, Hp, Wp;
  int per;                              // Hp * Wp positions per image
  int npos;                             // N * per
};

__device__ __forceinline__ uint32_t cvt_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Software pipeline of the staging: FETCH (global -> registers; issued right after the MMAs of the previous tile, so
// the DRAM latency hides behind them) and DEPOSIT (registers -> converted operand planes in shared memory, once those
// MMAs have retired).  Item i of thread t: unit = t % UNITS, position j = t / UNITS + i * (NT / UNITS); QW float4 per
// item (QW = 1: 4 channels, tf32 planes; QW = 2: 8 channels, bf16 planes).  AS_OUTPUT: global position g is the OUTPUT
// pixel (yy, xx) (valid for yy < H, xx < W); otherwise the zero-padded INPUT pixel (yy - 1, xx - 1).
template <int C, int QW, bool AS_OUTPUT, int MAXI>
__device__ __forceinline__ void fetch_items(const float* __restrict__ src, const Geo& G, int g0, int npt, float4 (&v)[MAXI][QW]) {
  constexpr int UNITS = C / (4 * QW), STEP = NT / UNITS;
  const int u0 = threadIdx.x % UNITS;
  const int j0 = threadIdx.x / UNITS;
  int g = g0 + j0;
  int n = g / G.per;
  int q = g - n * G.per;
  int yy = q / G.Wp;
  int xx = q - yy * G.Wp;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = j0 + i * STEP;
    const int iy = AS_OUTPUT ? yy : yy - 1, ix = AS_OUTPUT ? xx : xx - 1;
    const bool ok = j < npt && n < G.N && iy >= 0 && iy < G.H && ix >= 0 && ix < G.W;
#pragma unroll
    for (int k = 0; k < QW; ++k) v[i][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) {
      const float4* p = reinterpret_cast<const float4*>(src + ((int64_t)(n * G.H + iy) * G.W + ix) * C + 4 * QW * u0);
#pragma unroll
      for (int k = 0; k < QW; ++k) v[i][k] = __ldg(p + k);
    }
    xx += STEP;                                              // advance the decoded position by STEP
    while (xx >= G.Wp) { xx -= G.Wp; ++yy; }
    while (yy >= G.Hp) { yy -= G.Hp; ++n; }
  }
}

// tf32 planes: NS = 1: one tile of tf32-rounded values; NS = 2: H tile and, `lo_off` bytes further, the L = v - H tile.
template <int C, int NS, int MAXI>
__device__ __forceinline__ void deposit_tf32(const float4 (&v)[MAXI][1], int npt, uint8_t* planes, int PS, int lo_off) {
  constexpr int C4 = C / 4, STEP = NT / C4;
  uint8_t* dst = planes + (threadIdx.x % C4) * PS;
  const int j0 = threadIdx.x / C4;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = j0 + i * STEP;
    if (j >= npt) continue;
    uint8_t* d = dst + j * 16;
    const float4 t = v[i][0];
    if (NS == 1) {
      *reinterpret_cast<uint4*>(d) = make_uint4(cvt_tf32(t.x), cvt_tf32(t.y), cvt_tf32(t.z), cvt_tf32(t.w));
    } else {
      const uint4 h = make_uint4(__float_as_uint(t.x) & 0xFFFFE000u, __float_as_uint(t.y) & 0xFFFFE000u,
                                 __float_as_uint(t.z) & 0xFFFFE000u, __float_as_uint(t.w) & 0xFFFFE000u);
      *reinterpret_cast<uint4*>(d) = h;
      *reinterpret_cast<float4*>(d + lo_off) = make_float4(t.x - __uint_as_float(h.x), t.y - __uint_as_float(h.y),
                                                           t.z - __uint_as_float(h.z), t.w - __uint_as_float(h.w));
    }
  }
}

// plane stride: bytes per plane, padded so that the 16-byte stores of one quarter-warp (consecutive threads =
// consecutive planes of the same position) hit distinct banks: PS = 32 (C == 16) or 16 (otherwise) modulo 128
__host__ __device__ constexpr int plane_stride(int C, int npt) {
  int ps = npt * 16;
  const int want = (C == 16) ? 32 : 16;
  while ((ps % 128) != want) ps += 16;
  return ps;
}

// ---------------------------------------------------------------------------------------------------------------
// Forward / data gradient:  out[pos, co] = sum_{tap} sum_{ci} in[pos + shift(tap), ci] * Wt[tap][co][ci]
//   FLIP = false: Wt[tap = (kh, kw)][co][ci] = w[co][kh][kw][ci]                 (forward; in = x, out = y)
//   FLIP = true : Wt[tap][ci][co] = w[co][2 - kh][2 - kw][ci], roles swapped      (data gradient; in = gy, out = gx)
// A = staged input planes (K-major, M = 128 positions per MMA), B = weights in shared memory (K-major, N = C rows),
// accumulators in TMEM: per M tile C columns (+ C columns for the cross terms of the split mode).
//
// MMA count, not MMA size, bounds this kernel: a tcgen05.mma with N = 16 occupies the tensor core for ~80 clocks whatever
// its 16 K MACs would need (measured with %globaltimer stamps, tools/conv_trace.py: 36 MMAs per 256 positions took 3.0 us
// of a 4.1 us tile).  So the three kw taps are folded into N:
//     D[q, (kw, co)] = sum_{kh} sum_{ci} in[q + kh Wp, ci] * Wt[kh][kw][co][ci]          (N = 3 C, 3 * C/8 MMAs per M tile)
//     out[p, co]     = D[p, 0, co] + D[p + 1, 1, co] + D[p + 2, 2, co]
// The kw shift has moved from the A descriptor to the epilogue, where accumulator row p + 1 is simply the NEXT LANE of the
// warp: two __shfl_down per value; the last two lanes of a warp take the rows from a small shared-memory exchange the
// next warp fills, and the last two rows of a tile belong to the next tile (tiles advance by MO = MT - 2 positions).
template <int C>
struct FwdCfg {
  static constexpr int MT = 256;                                   // accumulator rows per tile (two M = 128 MMAs)
  static constexpr int MO = MT - 2;                                // output positions per tile
  static constexpr int W_TAP = C * C * 4;                          // one tap's weight tile, dense K-major
};

template <int C, int NS, bool FLIP, int MAXI, bool STATS, bool BNRED>
__global__ void __launch_bounds__(NT)
conv3x3_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, Geo G, int ntiles,
                   int npt, int PS, BnStat bs, BnRed br) {
  static_assert(!(STATS && BNRED) && (!BNRED || (FLIP && (C == 16 || C == 32))), "epilogue reductions");
  // BNRED staging by cp.async: x, y and gy2 at C = 16 (48 KB); x alone at C = 32 (32 KB: two CTAs still share an SM), y and
  // gy2 -- just written / just read by the weight gradient, i.e. L2 hits -- are loaded in the epilogue itself
  constexpr bool STAGE_ALL = C == 16;
  using F = FwdCfg<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int W_BYTES = 9 * F::W_TAP;
  uint8_t* wsm = smem;                                             // [NS][9][C x C]
  uint8_t* planes = smem + NS * W_BYTES;                           // [NS][C/4 planes][npt positions]
  const int tile_bytes = (C / 4) * PS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(planes + NS * tile_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int stage_off = NS * W_BYTES + NS * tile_bytes + 64;      // BNRED staging behind barrier + TMEM slot (16-byte aligned)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NF = 3 * C;                                        // MMA N: (kw, co)
  constexpr int ACC = (NS == 2) ? 2 * NF : NF;                     // columns per M tile (main | cross)
  constexpr int TCOLS = 2 * ACC <= 128 ? 128 : 2 * ACC <= 256 ? 256 : 512;      // two M tiles; power of two
  static_assert(2 * ACC <= 512 && NF % 16 == 0 && NF <= 256, "forward MMA shape");
  __shared__ __align__(16) float xch[NT / 32 + 1][3 * 16];                       // warp-edge rows of the kw fold, 16 channels at a time

  TRACE(0);
#ifdef ALIGNQ_CONV_TRACE
  trace_smid();
#endif
  pdl_trigger();                  // the bn-act kernel that follows touches no memory before its own pdl_wait
  // Prologue that does not depend on the preceding kernel (it may run beside that kernel's tail, see pdl_wait in
  // common.cuh): barrier, TMEM columns, and the quantized weights, which the weight bank wrote at the start of the step.
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  // ---- weights -> shared memory (K-major N x K tiles per tap), once per CTA ---------------------------------------
  // B-operand element (row r = N index, k): (r / 8) * SBO_W + (r % 8) * 16 + (k / 4) * 128 + (k % 4) * 4, SBO_W = (C/4) * 128
  constexpr int SBO_W = (C / 4) * 128;
  {
    // item = (tap, row co, 4 consecutive ci): one 16-byte load of w[co][kh][kw][4 ci4 ..].  Loads are issued in batches
    // (one dependent round trip to L2 per batch: a loop that loaded and stored item by item spent 6 us here at C = 32)
    // and then deposited: forward -> one 16-byte row chunk (row co, k = ci); data gradient -> four scalars, transposed
    // (rows ci, k = co) into the flipped tap 8 - tap.
    constexpr int NITEM = 9 * C * (C / 4), PER = (NITEM + NT - 1) / NT, BATCH = PER < 12 ? PER : 12;
#pragma unroll 1
    for (int b0 = 0; b0 < PER; b0 += BATCH) {
      float4 t[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int idx = threadIdx.x + (b0 + i) * NT;
        t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < NITEM) t[i] = __ldg(reinterpret_cast<const float4*>(w) + idx);      // physical [co][tap][ci]: idx = (co * 9 + tap) * C/4 + ci4
      }
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int idx = threadIdx.x + (b0 + i) * NT;
        if (idx >= NITEM) continue;
        const int ci4 = idx % (C / 4), tap = (idx / (C / 4)) % 9, co = idx / (9 * (C / 4));
        const float v[4] = {t[i].x, t[i].y, t[i].z, t[i].w};
        if (!FLIP) {
          uint8_t* d = wsm + tap * F::W_TAP + (co >> 3) * SBO_W + (co & 7) * 16 + ci4 * 128;
          if (NS == 1) {
            *reinterpret_cast<uint4*>(d) = make_uint4(cvt_tf32(v[0]), cvt_tf32(v[1]), cvt_tf32(v[2]), cvt_tf32(v[3]));
          } else {
            uint32_t h[4];
            float l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { h[e] = __float_as_uint(v[e]) & 0xFFFFE000u; l[e] = v[e] - __uint_as_float(h[e]); }
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(d + W_BYTES) = make_float4(l[0], l[1], l[2], l[3]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = 4 * ci4 + e;                             // row = ci, k = co
            uint8_t* d = wsm + (8 - tap) * F::W_TAP + (r >> 3) * SBO_W + (r & 7) * 16 + (co >> 2) * 128 + (co & 3) * 4;
            if (NS == 1) {
              *reinterpret_cast<uint32_t*>(d) = cvt_tf32(v[e]);
            } else {
              const uint32_t h = __float_as_uint(v[e]) & 0xFFFFE000u;
              *reinterpret_cast<uint32_t*>(d) = h;
              *reinterpret_cast<float*>(d + W_BYTES) = v[e] - __uint_as_float(h);
            }
          }
        }
      }
    }
  }
  // the input is the preceding kernel's output: wait for it, then request the first tile's positions
  pdl_wait();
  float4 fv[MAXI][1];
  if ((int)blockIdx.x < ntiles) fetch_items<C, 1, false, MAXI>(in, G, blockIdx.x * F::MO, npt, fv);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(2u /*tf32*/, 128u, (uint32_t)NF);
  TRACE(1);

  // this thread's positions: per-channel (sum, sum of squares) [STATS] or (sum g_z, sum g_z xhat) [BNRED]
  // Every tile is reduced over the warp at once (butterfly, bn_stat.cuh) and the warp totals are kept in shared memory:
  // no per-thread accumulators (2 C registers: the C = 32 kernel had 170 and one CTA per SM)
  constexpr bool RED = STATS || BNRED;
  __shared__ float wred[RED ? NT / 32 : 1][RED ? C / 16 : 1][32];  // [warp][channel group][lane]: channel lane & 15, sum (lane >> 4)
  if constexpr (RED) {
#pragma unroll
    for (int gq = 0; gq < C / 16; ++gq) wred[warp][gq][lane] = 0.f;
  }
  // BNRED staging: x, y, gy2 of this thread's output position arrive by cp.async while the MMAs run ([12][NT] float4)
  float4* stage = reinterpret_cast<float4*>(smem + stage_off);
  __shared__ float bnp[BNRED ? 4 : 1][BNRED ? C : 1];              // mean | invstd | gamma | beta of the bn-act layer
  if constexpr (BNRED) if (threadIdx.x < C) {                                  // (forward-pass data: long complete; visible after the sync below)
    bnp[0][threadIdx.x] = br.mean[threadIdx.x];
    bnp[1][threadIdx.x] = br.invstd[threadIdx.x];
    bnp[2][threadIdx.x] = br.gamma ? br.gamma[threadIdx.x] : 1.f;
    bnp[3][threadIdx.x] = br.beta ? br.beta[threadIdx.x] : 0.f;
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int g0 = tile * F::MO;
    // ---- 1. deposit the fetched input positions g0 .. g0 + MT + 2 Wp (the planes are free: last tile's MMAs retired)
    deposit_tf32<C, NS, MAXI>(fv, npt, planes, PS, tile_bytes);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    TRACE(2 + 4 * it);
    // ---- 2. MMAs ---------------------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(planes), sw = smem_u32(wsm);
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t d_main = tmem_base + mt * ACC, d_cross = d_main + NF;
#pragma unroll 1
        for (int kh = 0; kh < 3; ++kh) {
          const uint32_t shift = (uint32_t)(kh * G.Wp + mt * 128) * 16u;
          // the three kw taps of row kh are consecutive [C x C] tiles: one K-major [3 C x C] B operand (rows (kw, co))
          const uint32_t wk = sw + kh * 3 * F::W_TAP;
#pragma unroll
          for (int ks = 0; ks < C / 8; ++ks) {
            const uint32_t acc = (kh > 0 || ks > 0) ? 1u : 0u;
            const uint64_t ah = make_desc(sa + shift + ks * 2 * PS, PS, 128);
            const uint64_t bh = make_desc(wk + ks * 2 * 128, 128, SBO_W);
            umma<true>(d_main, ah, bh, IDESC, acc);
            if (NS == 2) {
              const uint64_t al = make_desc(sa + tile_bytes + shift + ks * 2 * PS, PS, 128);
              const uint64_t bl = make_desc(wk + W_BYTES + ks * 2 * 128, 128, SBO_W);
              umma<true>(d_cross, ah, bl, IDESC, acc);
              umma<true>(d_cross, al, bh, IDESC, 1u);
            }
          }
        }
      }
      umma_commit(bar);
      TRACE(3 + 4 * it);
    }
    // ---- 3. next tile's positions -> registers while the tensor core works -----------------------------------------
    if (tile + (int)gridDim.x < ntiles) fetch_items<C, 1, false, MAXI>(in, G, (tile + gridDim.x) * F::MO, npt, fv);
    // this thread's output position (warp w -> M tile w / 4, TMEM lanes 32 (w % 4) .. +31)
    const int mt = warp >> 2;
    const int j = mt * 128 + (warp & 3) * 32 + lane;               // position inside the tile
    const int g = g0 + j;
    const int n = g / G.per;
    const int q = g - n * G.per;
    const int yy = q / G.Wp, xx = q - yy * G.Wp;
    const bool ok = j < F::MO && g < G.npos && yy < G.H && xx < G.W;
    const int64_t ooff = ((int64_t)(n * G.H + yy) * G.W + xx) * C;
    // BNRED: the bn-act layer's x, y (and the second consumer's gradient) at this position, requested before the wait
    if constexpr (BNRED) {
      if (ok) {
#pragma unroll
        for (int k = 0; k < C / 4; ++k) {
          tcsmall::cp_async16(smem_u32(stage + (0 * (C / 4) + k) * NT + threadIdx.x), reinterpret_cast<const float4*>(br.x + ooff) + k);
          if (STAGE_ALL && br.relu)
            tcsmall::cp_async16(smem_u32(stage + (1 * (C / 4) + k) * NT + threadIdx.x), reinterpret_cast<const float4*>(br.y + ooff) + k);
          if (STAGE_ALL && br.gy2)
            tcsmall::cp_async16(smem_u32(stage + (2 * (C / 4) + k) * NT + threadIdx.x), reinterpret_cast<const float4*>(br.gy2 + ooff) + k);
        }
      }
      tcsmall::cp_async_commit();
    }
    mbar_wait(bar, (uint32_t)(it & 1));
    tc_fence_after();
    TRACE(4 + 4 * it);
    // ---- 4. epilogue: one output position per thread -----------------------------------------------------------------
    if constexpr (BNRED) tcsmall::cp_async_wait_all();             // this thread's own staged x / y / gy2
    {
      float* o = out + ooff;
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + mt * ACC;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 16) {
        // accumulator row of this thread: kw = 0 | 1 | 2 blocks of C columns (+ the cross-term set NF columns further)
        float r[16], b1[16], c2[16];
        {
          uint32_t v[16], v1[16], v2[16];
          tmem_ld16_nowait(ta + c0, v);                            // three loads in flight, one wait
          tmem_ld16_nowait(ta + C + c0, v1);
          tmem_ld16_nowait(ta + 2 * C + c0, v2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) { r[e] = __uint_as_float(v[e]); b1[e] = __uint_as_float(v1[e]); c2[e] = __uint_as_float(v2[e]); }
          if (NS == 2) {
            tmem_ld16(ta + NF + c0, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) r[e] += __uint_as_float(v[e]);
            tmem_ld16(ta + NF + C + c0, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) b1[e] += __uint_as_float(v[e]);
            tmem_ld16(ta + NF + 2 * C + c0, v);
#pragma unroll
            for (int e = 0; e < 16; ++e) c2[e] += __uint_as_float(v[e]);
          }
        }
        // rows p + 1 (kw = 1) and p + 2 (kw = 2): the next lanes; across the warp edge through shared memory
        if (c0 > 0) __syncthreads();                               // the previous channel group's exchange has been read
        if (lane < 2) {
          float4* xw = reinterpret_cast<float4*>(xch[warp]);
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            if (lane == 0) xw[e / 4] = make_float4(b1[e], b1[e + 1], b1[e + 2], b1[e + 3]);
            xw[(lane == 0 ? 4 : 8) + e / 4] = make_float4(c2[e], c2[e + 1], c2[e + 2], c2[e + 3]);
          }
        }
        __syncthreads();
        {
          // branch-free: every lane reads the (broadcast) exchange rows and selects; divergent `if (lane == 31)` loads
          // compiled to 32 branch / reconverge pairs and cost 2.3 us per tile
          const float4* xb = reinterpret_cast<const float4*>(xch[warp + 1]);
          const float4* xc = xb + (lane == 31 ? 8 : 4);
          const bool e31 = lane == 31, e30 = lane >= 30;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 yb4 = xb[e / 4], yc4 = xc[e / 4];
            const float yb[4] = {yb4.x, yb4.y, yb4.z, yb4.w}, yc[4] = {yc4.x, yc4.y, yc4.z, yc4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float nb = __shfl_down_sync(0xffffffffu, b1[e + k], 1);
              const float nc = __shfl_down_sync(0xffffffffu, c2[e + k], 2);
              r[e + k] = (r[e + k] + (e31 ? yb[k] : nb)) + (e30 ? yc[k] : nc);
            }
          }
        }
        if (ok) {
          if constexpr (BNRED) {
            if (br.gy2) {                          // r becomes the bn-act output's total upstream gradient
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 g4 = STAGE_ALL ? stage[(2 * (C / 4) + (c0 + e) / 4) * NT + threadIdx.x]
                                            : __ldg(reinterpret_cast<const float4*>(br.gy2 + ooff + c0 + e));
                r[e] += g4.x; r[e + 1] += g4.y; r[e + 2] += g4.z; r[e + 3] += g4.w;
              }
            }
          }
#pragma unroll
          for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(o + c0 + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
        }
        if constexpr (STATS) {                                     // per-channel sum and sum of squares of this tile's rows
          float v[32];
#pragma unroll
          for (int e = 0; e < 16; ++e) { v[e] = ok ? r[e] : 0.f; v[16 + e] = ok ? r[e] * r[e] : 0.f; }
          wred[warp][c0 / 16][lane] += warp_reduce_transpose32(v, lane);
        }
        if constexpr (BNRED) {
          // g_z and the two BatchNorm-backward sums of this tile: v[0..15] = g_z, v[16..31] = g_z xhat per channel, summed
          // over the warp at once (lane l keeps the total of value l), accumulated in the warp's shared-memory row
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0.f;
          if (ok) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 x4 = stage[(0 * (C / 4) + (c0 + e) / 4) * NT + threadIdx.x];
              float4 y4 = make_float4(1.f, 1.f, 1.f, 1.f);
              if (br.relu) y4 = STAGE_ALL ? stage[(1 * (C / 4) + (c0 + e) / 4) * NT + threadIdx.x]
                                          : __ldg(reinterpret_cast<const float4*>(br.y + ooff + c0 + e));
              const float xx4[4] = {x4.x, x4.y, x4.z, x4.w}, yy4[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = c0 + e + k;
                const float xh = (xx4[k] - bnp[0][c]) * bnp[1][c];
                const float gz = bnq_gz_raw(fmaf(xh, bnp[2][c], bnp[3][c]), r[e + k], yy4[k], br.gscale, br.relu);
                v[e + k] = gz;
                v[16 + e + k] = gz * xh;
              }
            }
          }
          wred[warp][c0 / 16][lane] += warp_reduce_transpose32(v, lane);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                               // accumulators and planes are free again
    TRACE(5 + 4 * it);
  }
  TRACE(13);
  if constexpr (STATS) {
    // CTA totals -> fp64 atomics -> the last CTA finishes mean / invstd / running statistics (bn_stat.cuh); the planes are idle
    double* red = reinterpret_cast<double*>(planes);               // [warp][2 C]
#pragma unroll
    for (int gq = 0; gq < C / 16; ++gq) red[warp * 2 * C + 2 * (16 * gq + (lane & 15)) + (lane >> 4)] = (double)wred[warp][gq][lane];
    bn_cta_finish_from_red<C, NT>(red, bs.ws, bs.counter, [&](int c, double S, double SS) { bn_stat_finalize(bs, c, S, SS); });
  }
  if constexpr (BNRED) {
    // the preceding bn-act layer's backward sums; the last CTA leaves mean(g_z), mean(g_z xhat) and the affine gradients
    double* red = reinterpret_cast<double*>(planes);               // [warp][2 C] (the planes are idle)
#pragma unroll
    for (int gq = 0; gq < C / 16; ++gq) red[warp * 2 * C + 2 * (16 * gq + (lane & 15)) + (lane >> 4)] = (double)wred[warp][gq][lane];
    bn_cta_finish_from_red<C, NT>(red, br.ws, br.counter, [&](int c, double S, double SS) {
      if (br.gbeta) br.gbeta[c] = (float)S;
      if (br.ggamma) br.ggamma[c] = (float)SS;
      br.coef[2 * c] = (float)(S / br.count);
      br.coef[2 * c + 1] = (float)(SS / br.count);
    });
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
  TRACE(14);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient:  dW[co][kh][kw][ci] = sum_pos gy[pos, co] * x[pos + kh Wp + kw, ci]       (split-K over CTAs)
// Both operands are MN-major here (MN = channels, K = positions).  kind::tf32 does not take MN-major operands in the
// un-swizzled layout (measured: the MMA returns zeros), so the operands are bf16 TERMS of the fp32 values, 8 channels
// per 16-byte unit:   byte address = plane_base + unit * PS + position * 16.
//   NSB = 2 (mode TF32):   v = H + M, 16 mantissa bits (more than tf32's 11), products H H | H M + M H     (3 MMAs)
//   NSB = 3 (mode TF32X3): v = H + M + L, 24 bits, products H H | H M + M H + H L + L H + M M               (6 MMAs)
// (the large product in the main accumulator, the small ones in a second one: the tensor core truncates when it adds
// into the fp32 accumulator).
// A tcgen05.mma costs ~50 clocks to issue whatever its size (measured: a first version with one M = 64, N = C, K = 16
// MMA per tap and k-step spent 88 us on [128, 16, 32, 32], all of it MMA issue), so the taps are folded into N: the
// x operand is staged as an im2col-in-shared-memory of TAPS shifted copies, unit u = tap * (C/8) + c8, and ONE
// M = 64, N = TAPS * C, K = 16 MMA per k-step and product covers all the CTA's taps.  A = gy planes (M = 64 rows of
// which C are real; the units past C/8 read whatever follows -- rows of D that are never stored).  Accumulator row m
// is TMEM lane (m % 16) + 32 (m / 16), columns tap * C + ci (+ TAPS * C for the cross terms).  Every CTA writes its
// partial [C][TAPS * C] to the workspace; conv3x3_wgrad_reduce_kernel sums them in a fixed order.
template <int C>
struct WgCfg {
  static constexpr int KT = 128;                                   // positions (K) per tile
};

__host__ __device__ constexpr int plane_stride_bf16(int C, int npt) {
  int ps = npt * 16;
  const int want = (C == 16) ? 64 : (C == 32) ? 32 : 16;          // conflict-free 16-byte stores of one quarter-warp
  while ((ps % 128) != want) ps += 16;
  return ps;
}

// bf16-term planes (tiles `term_off` bytes apart) from fetched 8-channel items.
// IM2COL = false: unit c8, position j.   IM2COL = true (the x operand; npt = KT + 2 Wp + 2 halo positions were fetched):
// source position j is written to unit (tap, c8) at position j - shift(tap) for every tap of the CTA that needs it.
// NSB = 1: ONE fp16 term per value (11 significand bits = tf32's), the value multiplied by a power of two first (exact)
// so that the tile's largest magnitude sits near 2^14 -- fp16's narrow exponent range then costs nothing: elements
// more than 2^28 below the largest one lose bits, and their products are that far below the sum's own rounding error.
// Saturating: |v| scale > 65504 (only possible for the unscaled x operand) becomes +-65504, never inf.
__device__ __forceinline__ uint4 pack_fp16x8(const float (&c)[8], float scale) {
  uint32_t h[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float a = fminf(fmaxf(c[2 * p] * scale, -65504.f), 65504.f), b = fminf(fmaxf(c[2 * p + 1] * scale, -65504.f), 65504.f);
    const __half2 hh = __floats2half2_rn(a, b);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  return make_uint4(h[0], h[1], h[2], h[3]);
}
template <int NSB>
__device__ __forceinline__ void make_terms(const float (&c)[8], float scale, uint4 (&t)[NSB]) {
  if constexpr (NSB == 1) t[0] = pack_fp16x8(c, scale);
  else tcsmall::split_chunk_n<NSB>(c, t);
}

template <int C, int NSB, bool IM2COL, int TAPS, int MAXI>
__device__ __forceinline__ void deposit_bf16(const float4 (&v)[MAXI][2], const Geo& G, int npt, int kt, int tap0,
                                             uint8_t* planes, int PS, int term_off, float scale = 1.f) {
  constexpr int C8 = C / 8, STEP = NT / C8;
  const int c8 = threadIdx.x % C8;
  const int j0 = threadIdx.x / C8;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) {
    const int j = j0 + i * STEP;
    if (j >= npt) continue;
    const float c[8] = {v[i][0].x, v[i][0].y, v[i][0].z, v[i][0].w, v[i][1].x, v[i][1].y, v[i][1].z, v[i][1].w};
    uint4 t[NSB];
    make_terms<NSB>(c, scale, t);
    if (!IM2COL) {
#pragma unroll
      for (int k = 0; k < NSB; ++k) *reinterpret_cast<uint4*>(planes + c8 * PS + j * 16 + k * term_off) = t[k];
    } else {
#pragma unroll
      for (int tt = 0; tt < TAPS; ++tt) {
        const int tap = tap0 + tt;
        const int jd = j - ((tap / 3) * G.Wp + (tap % 3));
        if (jd >= 0 && jd < kt) {
          uint8_t* d = planes + (tt * C8 + c8) * PS + jd * 16;
#pragma unroll
          for (int k = 0; k < NSB; ++k) *reinterpret_cast<uint4*>(d + k * term_off) = t[k];
        }
      }
    }
  }
}

template <int C, int NSB, int TAPS, bool CROSS, int MAXG, int MAXX, int NBUF>
__global__ void __launch_bounds__(NT)
conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partials, Geo G,
                     int ntiles, int npx, int PS, float* __restrict__ inv_scales) {
  using K = WgCfg<C>;
  constexpr int C8 = C / 8;
  extern __shared__ __align__(128) uint8_t smem[];
  // NBUF operand buffers: with two, the staging of tile t + 1 runs beside the MMAs of tile t (measured per tile with one
  // buffer: deposit 0.8 us, then MMAs 0.9 us, strictly one after the other -- tools/conv_trace.py)
  const int gtile = C8 * PS, xtile = TAPS * C8 * PS;
  const int bufbytes = NSB * (gtile + xtile);
  uint8_t* gplanes0 = smem;                                        // [NSB][C8 units][KT]
  uint8_t* xplanes0 = smem + NSB * gtile;                          // [NSB][TAPS * C8 units][KT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NBUF * bufbytes);        // one barrier per buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ACC = TAPS * C;                                    // columns of one accumulator set = MMA N
  // CROSS: the small products get their own accumulator (fp32-parity mode); without it everything shares one, which
  // halves the TMEM columns so that two CTAs fit on an SM (a second 512-column allocation would wait for the first)
  constexpr int NEED = CROSS ? 2 * ACC : ACC;
  constexpr int TCOLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
  constexpr int NSPLIT = ACC > 256 ? 2 : 1;                        // an MMA's N is at most 256: wider sets take two
  constexpr int NMMA = ACC / NSPLIT;
  static_assert(NEED <= 512 && NMMA <= 256 && NMMA % 16 == 0 && (TAPS * C8) % NSPLIT == 0, "weight-gradient MMA shape");
  static_assert(NSB != 1 || !CROSS, "the single-term mode has one accumulator set");
  const int tap0 = blockIdx.y * TAPS;
  TRACE(0);
#ifdef ALIGNQ_CONV_TRACE
  trace_smid();
#endif

  pdl_trigger();                                                   // the reduce kernel launches early and waits for us
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // bf16 terms (fp16 in the single-term mode), both operands MN-major (bits 15, 16)
  constexpr uint32_t IDESC = make_idesc(NSB == 1 ? 0u /*fp16*/ : 1u /*bf16*/, 64u, (uint32_t)NMMA) | (1u << 15) | (1u << 16);

  float4 fg[MAXG][2], fx[MAXX][2];
  // Single-term mode: this CTA's power-of-two scale for gy, from the largest |gy| over ITS tiles (a first pass over gy,
  // which the main pass then finds in L2); the partial sums are scaled back by the reduce kernel, exactly.
  float gscale = 1.f, ginv = 1.f;
  if constexpr (NSB == 1) {
    __shared__ float wmax[NT / 32];
    float amax = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      fetch_items<C, 2, true, MAXG>(gy, G, tile * K::KT, K::KT, fg);
#pragma unroll
      for (int i = 0; i < MAXG; ++i)
#pragma unroll
        for (int k = 0; k < 2; ++k)
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(fg[i][k].x), fabsf(fg[i][k].y)), fmaxf(fabsf(fg[i][k].z), fabsf(fg[i][k].w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0) wmax[warp] = amax;
    __syncthreads();
#pragma unroll
    for (int wv = 0; wv < NT / 32; ++wv) amax = fmaxf(amax, wmax[wv]);
    const int e = (int)((__float_as_uint(amax) >> 23) & 0xffu);      // amax in [2^(e-127), 2^(e-126))
    if (e >= 16 && e < 255) {                                        // (zero / denormal-tiny / inf / nan: unscaled)
      gscale = __uint_as_float((uint32_t)(268 - e) << 23);           // 2^(14 - (e - 127)): scaled amax in [2^14, 2^15)
      ginv = __uint_as_float((uint32_t)(e - 14) << 23);
    }
  }
  TRACE(1);
  if ((int)blockIdx.x < ntiles) {
    fetch_items<C, 2, true, MAXG>(gy, G, blockIdx.x * K::KT, K::KT, fg);
    fetch_items<C, 2, false, MAXX>(x, G, blockIdx.x * K::KT, npx, fx);
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = NBUF == 2 ? (it & 1) : 0;
    uint8_t* gplanes = gplanes0 + b * bufbytes;
    uint8_t* xplanes = xplanes0 + b * bufbytes;
    if (NBUF == 2 && it >= 2) {                                    // this buffer's previous MMAs (tile it - 2) have retired
      mbar_wait(bar + b, (uint32_t)(((it >> 1) - 1) & 1));
      tc_fence_after();
    }
    deposit_bf16<C, NSB, false, 1, MAXG>(fg, G, K::KT, K::KT, 0, gplanes, PS, gtile, gscale);   // gy is zero at pad positions
    deposit_bf16<C, NSB, true, TAPS, MAXX>(fx, G, npx, K::KT, tap0, xplanes, PS, xtile);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (it < 2) TRACE(2 + 4 * it);
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sg = smem_u32(gplanes), sx = smem_u32(xplanes);
      const uint32_t d_main = tmem_base, d_cross = CROSS ? tmem_base + ACC : tmem_base;
#pragma unroll
      for (int ks = 0; ks < K::KT / 16; ++ks) {
        const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
        // MN-major: LBO = stride of the 8-position K groups (128 B: K is linear in the position), SBO = unit stride
        const uint32_t ga = sg + ks * 256, xa = sx + ks * 256;
        const uint64_t ah = make_desc(ga, 128, PS), bh = make_desc(xa, 128, PS);
        if constexpr (NSB == 1) {                                            // one product; N in NSPLIT pieces
#pragma unroll
          for (int h = 0; h < NSPLIT; ++h)
            umma<false>(d_main + h * NMMA, ah, make_desc(xa + h * (TAPS * C8 / NSPLIT) * PS, 128, PS), IDESC, acc);
          continue;
        }
        const uint64_t am = make_desc(ga + gtile, 128, PS), bm = make_desc(xa + xtile, 128, PS);
        umma<false>(d_main, ah, bh, IDESC, acc);                             // H H
        umma<false>(d_cross, ah, bm, IDESC, CROSS ? acc : 1u);               // the small products
        umma<false>(d_cross, am, bh, IDESC, 1u);
        if (NSB == 3) {
          const uint64_t al = make_desc(ga + 2 * gtile, 128, PS), bl = make_desc(xa + 2 * xtile, 128, PS);
          umma<false>(d_cross, ah, bl, IDESC, 1u);
          umma<false>(d_cross, al, bh, IDESC, 1u);
          umma<false>(d_cross, am, bm, IDESC, 1u);
        }
      }
      umma_commit(bar + b);
      if (it < 2) TRACE(3 + 4 * it);
    }
    if (tile + (int)gridDim.x < ntiles) {                          // next tile -> registers while the tensor core works
      fetch_items<C, 2, true, MAXG>(gy, G, (tile + gridDim.x) * K::KT, K::KT, fg);
      fetch_items<C, 2, false, MAXX>(x, G, (tile + gridDim.x) * K::KT, npx, fx);
    }
    if (NBUF == 1) {
      mbar_wait(bar, (uint32_t)(it & 1));                          // the planes are rewritten by the next deposit
      tc_fence_after();
      if (it < 2) TRACE(4 + 4 * it);
      tc_fence_before();
      __syncthreads();
      if (it < 2) TRACE(5 + 4 * it);
    }
  }
  if (NBUF == 2 && it > 0) {                                       // MMAs complete in order: the last commit covers them all
    mbar_wait(bar + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
    tc_fence_after();
  }
  TRACE(13);
  if (threadIdx.x == 0 && inv_scales) inv_scales[blockIdx.y * gridDim.x + blockIdx.x] = ginv;
  // ---- epilogue: accumulator row co = 16 w + l sits in lane l of warp w's quarter (w < C / 16, l < 16) ------------
  // (all 32 lanes of a warp must take part in tcgen05.ld: the store is predicated instead)
  if (warp < 4) {
    const bool mine = warp < C / 16 && lane < 16;
    const int co = 16 * warp + lane;
    float* o = partials + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * C + (mine ? co : 0)) * ACC;
    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < ACC; c0 += 16) {
      uint32_t v[16];
      float r[16];
      if (it > 0) {
        tmem_ld16(ta + c0, v);
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __uint_as_float(v[e]);
        if (CROSS) {
          uint32_t x2[16];
          tmem_ld16(ta + ACC + c0, x2);
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] += __uint_as_float(x2[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = 0.f;                    // a CTA without tiles contributes zeros
      }
      if (mine) {
#pragma unroll
        for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(o + c0 + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
  TRACE(14);
}

// gw[co][kh][kw][ci] (+)= sum over the CTAs' partials [grp][cta][co][TAPS * C], fixed order.  Block = 32 elements x
// 32 part lanes: lane pl sums parts pl, pl + 32, ...; the 32 lane sums are then added in a fixed order.  (A first
// version with one thread per element walking all ~300 partials took 26 us -- longer than the MMA kernel itself.)
__global__ void __launch_bounds__(1024)
conv3x3_wgrad_reduce_kernel(const float* __restrict__ partials, int nparts, int C, int taps, float* __restrict__ gw,
                            int accumulate, const float* __restrict__ inv_scales) {
  __shared__ float sm[32][33];
  pdl_wait();                                                     // launched as a programmatic dependent of the kernel above
  const int e = blockIdx.x * 32 + threadIdx.x;                    // index into [co][9][ci]
  const int pl = threadIdx.y;
  const int total = C * 9 * C;
  float s = 0.f;
  if (e < total) {
    const int ci = e % C, tap = (e / C) % 9, co = e / (9 * C);
    const int grp = tap / taps, t = tap - grp * taps;
    const int acc = taps * C;
    const float* p = partials + ((size_t)grp * nparts * C + co) * acc + t * C + ci;
    // every load of a batch is issued before the first add (a plain `s += p[..]` loop went to L2 once per part, one
    // after the other: 8 us for 296 parts); the order of the additions is fixed
    for (int k0 = pl; k0 < nparts; k0 += 32 * 8) {
      float v[8], sc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + 32 * u;
        v[u] = k < nparts ? p[(size_t)k * C * acc] : 0.f;            // (coherent loads: written by the kernel we depend on)
        sc[u] = (inv_scales && k < nparts) ? inv_scales[grp * nparts + k] : 1.f;             // per-CTA powers of two: exact
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u] * sc[u];
    }
  }
  sm[pl][threadIdx.x] = s;
  __syncthreads();
  if (pl == 0 && e < total) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) r += sm[k][threadIdx.x];
    gw[e] = accumulate ? gw[e] + r : r;
  }
}

// tuning knobs read from the environment on every launch set-up (host side, a getenv per call: nanoseconds)
inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

inline Geo make_geo(int N, int H, int W) {
  Geo G;
  G.N = N; G.H = H; G.W = W; G.Hp = H + 2; G.Wp = W + 2;
  G.per = G.Hp * G.Wp;
  G.npos = N * G.per;
  return G;
}

template <int C, int NS, bool FLIP, bool STATS = false, bool BNRED = false>
static int launch_fwd(const float* in, const float* w, float* out, int N, int H, int W, cudaStream_t s, BnStat bs = BnStat{},
                      BnRed br = BnRed{}) {
  using F = FwdCfg<C>;
  const Geo G = make_geo(N, H, W);
  const int ntiles = (G.npos + F::MO - 1) / F::MO;
  const int npt = F::MT + 2 * G.Wp + 2;
  const int PS = plane_stride(C, npt);
  const size_t smem = (size_t)NS * 9 * F::W_TAP + (size_t)NS * (C / 4) * PS + 64 + (BNRED ? (size_t)(C == 16 ? 3 : 1) * (C / 4) * NT * 16 : 0);
  if (smem > 227 * 1024) return ALIGNQ_ERANGE;
  // staging registers: items per thread = ceil(npt / (NT / (C/4))); the instantiation covers rows up to W = 32 + 2
  constexpr int STEP = NT / (C / 4);
  constexpr int MAXI = (F::MT + 2 * 34 + 2 + STEP - 1) / STEP;
  if (npt > MAXI * STEP) return ALIGNQ_ERANGE;                 // wider images: the caller's library convolution
  cudaError_t e = cudaFuncSetAttribute(conv3x3_fwd_kernel<C, NS, FLIP, MAXI, STATS, BNRED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // persistent CTAs (the weights are staged once per CTA): two per SM overlap each other's deposit / MMA / epilogue phases
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > env_int("ALIGNQ_CONV_PER_SM", 2)) per_sm = env_int("ALIGNQ_CONV_PER_SM", 2);
  int grid = ntiles < ALIGNQ_NUM_SMS * per_sm ? ntiles : ALIGNQ_NUM_SMS * per_sm;
  bs.count = (double)N * H * W;
  br.count = (double)N * H * W;
  // launched as a programmatic dependent of the preceding kernel (common.cuh: pdl_wait); ALIGNQ_PDL=0 switches it off
  e = launch_pdl(conv3x3_fwd_kernel<C, NS, FLIP, MAXI, STATS, BNRED>, dim3(grid), dim3(NT), smem, s, in, w, out, G, ntiles, npt, PS, bs, br);
  if (e != cudaSuccess) return (int)e;
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

template <int C, int NSB, int TAPS, bool CROSS, int NBUF = 1>
static int launch_wgrad(const float* x, const float* gy, float* gw, int N, int H, int W, int accumulate, float* ws,
                        size_t ws_bytes, cudaStream_t s) {
  using K = WgCfg<C>;
  const Geo G = make_geo(N, H, W);
  const int ntiles = (G.npos + K::KT - 1) / K::KT;
  const int npx = K::KT + 2 * G.Wp + 2;
  const int PS = plane_stride_bf16(C, K::KT);
  size_t smem = (size_t)NBUF * NSB * (C / 8) * PS * (1 + TAPS) + 64;
  const size_t reach = (size_t)(NSB - 1) * (C / 8) * PS + 8 * (size_t)PS + K::KT * 16;   // what the M = 64 descriptors may touch
  if (smem < reach + 64) smem = reach + 64;
  if (smem > 227 * 1024) return ALIGNQ_ERANGE;
  const int groups = 9 / TAPS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > env_int("ALIGNQ_WGRAD_PER_SM", 2)) per_sm = env_int("ALIGNQ_WGRAD_PER_SM", 2);
  if ((CROSS ? 2 : 1) * TAPS * C > 256) per_sm = 1;              // a 512-column TMEM allocation: a second CTA would wait
  if (per_sm < 1) per_sm = 1;
  int grid = ALIGNQ_NUM_SMS * per_sm / groups;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  const size_t need = (size_t)groups * grid * C * TAPS * C * sizeof(float);
  if (ws_bytes < need + (size_t)groups * grid * sizeof(float)) return ALIGNQ_ENOSPACE;
  float* inv_scales = NSB == 1 ? ws + need / sizeof(float) : nullptr;
  constexpr int STEP = NT / (C / 8);
  constexpr int MAXG = (K::KT + STEP - 1) / STEP, MAXX = (K::KT + 2 * 34 + 2 + STEP - 1) / STEP;
  if (npx > MAXX * STEP) return ALIGNQ_ERANGE;                  // wider images: the caller's library convolution
  cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<C, NSB, TAPS, CROSS, MAXG, MAXX, NBUF>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  conv3x3_wgrad_kernel<C, NSB, TAPS, CROSS, MAXG, MAXX, NBUF><<<dim3(grid, groups), NT, smem, s>>>(x, gy, ws, G, ntiles, npx, PS, inv_scales);
  ALIGNQ_LAUNCH_CHECK();
  e = launch_pdl(conv3x3_wgrad_reduce_kernel, dim3((C * 9 * C + 31) / 32), dim3(32, 32), 0, s, (const float*)ws, grid, C, TAPS, gw,
                 accumulate, (const float*)inv_scales);
  if (e != cudaSuccess) return (int)e;
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

}  // namespace ctc
}  // namespace alignq

using namespace alignq;
using namespace alignq::ctc;

static int conv_args_ok(const void* a, const void* b, const void* c, int N, int H, int W, int C, int mode) {
  if (!a || !b || !c || N < 1 || H < 1 || W < 1) return ALIGNQ_EINVAL;
  if (C != 16 && C != 32 && C != 64) return ALIGNQ_ERANGE;
  if (mode != ALIGNQ_CONV_TF32 && mode != ALIGNQ_CONV_TF32X3) return ALIGNQ_EINVAL;
  if (!aligned16(a) || !aligned16(b) || !aligned16(c)) return ALIGNQ_EALIGN;
  if ((int64_t)N * (H + 2) * (W + 2) > (int64_t)1 << 30) return ALIGNQ_ERANGE;
  return ALIGNQ_OK;
}

#define CONV_DISPATCH(C_, NS_, CALL)                                \
  do {                                                              \
    if ((C_) == 16) { if ((NS_) == 1) { CALL(16, 1); } else { CALL(16, 2); } } \
    else if ((C_) == 32) { if ((NS_) == 1) { CALL(32, 1); } else { CALL(32, 2); } } \
    else { if ((NS_) == 1) { CALL(64, 1); } else { return ALIGNQ_ERANGE; } }   /* 64 x3: TMEM holds one set only */ \
  } while (0)

extern "C" int alignq_conv3x3_fwd(const float* x, const float* w, float* y, int N, int H, int W, int C, int mode,
                                  alignq_stream_t stream) {
  int rc = conv_args_ok(x, w, y, N, H, W, C, mode);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ns = mode == ALIGNQ_CONV_TF32X3 ? 2 : 1;
#define CALL(CC, NN) return launch_fwd<CC, NN, false>(x, w, y, N, H, W, s)
  CONV_DISPATCH(C, ns, CALL);
#undef CALL
  return ALIGNQ_EINVAL;
}

extern "C" int alignq_conv3x3_fwd_bnstats(const float* x, const float* w, float* y, int N, int H, int W, int C, int mode,
                                          float* running_mean, float* running_var, float momentum, float bn_eps,
                                          float* save_mean, float* save_invstd, double* bn_ws, uint32_t* bn_counter,
                                          int64_t* num_batches_tracked, alignq_stream_t stream) {
  int rc = conv_args_ok(x, w, y, N, H, W, C, mode);
  if (rc) return rc;
  if (!save_mean || !save_invstd || !bn_ws || !bn_counter) return ALIGNQ_EINVAL;
  if (C != 16 && C != 32) return ALIGNQ_ERANGE;               // per-thread channel accumulators: narrow layers only
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  BnStat bs{bn_ws, bn_counter, running_mean, running_var, save_mean, save_invstd,
            reinterpret_cast<long long*>(num_batches_tracked), momentum, bn_eps, 0.0};
  if (mode == ALIGNQ_CONV_TF32X3) {
    if (C == 16) return launch_fwd<16, 2, false, true>(x, w, y, N, H, W, s, bs);
    return launch_fwd<32, 2, false, true>(x, w, y, N, H, W, s, bs);
  }
  if (C == 16) return launch_fwd<16, 1, false, true>(x, w, y, N, H, W, s, bs);
  return launch_fwd<32, 1, false, true>(x, w, y, N, H, W, s, bs);
}

extern "C" int alignq_conv3x3_bwd_data(const float* gy, const float* w, float* gx, int N, int H, int W, int C, int mode,
                                       alignq_stream_t stream) {
  int rc = conv_args_ok(gy, w, gx, N, H, W, C, mode);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int ns = mode == ALIGNQ_CONV_TF32X3 ? 2 : 1;
#define CALL(CC, NN) return launch_fwd<CC, NN, true>(gy, w, gx, N, H, W, s)
  CONV_DISPATCH(C, ns, CALL);
#undef CALL
  return ALIGNQ_EINVAL;
}

extern "C" int alignq_conv3x3_bwd_data_bnreduce(const float* gy, const float* w, float* gx, int N, int H, int W, int C, int mode,
                                                const float* bn_x, const float* bn_y, const float* gy2,
                                                const float* save_mean, const float* save_invstd, const float* gamma,
                                                const float* beta, float act_range, int relu, float* ggamma, float* gbeta,
                                                double* bn_ws, uint32_t* bn_counter, alignq_stream_t stream) {
  int rc = conv_args_ok(gy, w, gx, N, H, W, C, mode);
  if (rc) return rc;
  // C = 16 only.  (The kernel template also covers C = 32 -- x staged by cp.async, y / gy2 loaded in the epilogue -- but
  // the two-channel-group epilogue made the ResNet-20 step slower, 1.10-1.13 ms against 1.08, so it is not instantiated.)
  if (C != 16) return ALIGNQ_ERANGE;
  if (!bn_x || !save_mean || !save_invstd || !bn_ws || !bn_counter || (relu && !bn_y)) return ALIGNQ_EINVAL;
  if (!aligned16(bn_x) || (relu && !aligned16(bn_y)) || (gy2 && !aligned16(gy2))) return ALIGNQ_EALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // coefficients where the bn-act backward's apply pass looks for them: behind the three accumulator sets of the layer
  float* coef = reinterpret_cast<float*>(bn_ws + (size_t)3 * ALIGNQ_BN_SLOTS * C * 2);
  BnRed br{bn_x, bn_y, gy2, save_mean, save_invstd, gamma, beta, 2.0f * act_range * kInvSqrt2Pi, relu, bn_ws, bn_counter,
           coef, ggamma, gbeta, 0.0};
  if (mode == ALIGNQ_CONV_TF32X3) return launch_fwd<16, 2, true, false, true>(gy, w, gx, N, H, W, s, BnStat{}, br);
  return launch_fwd<16, 1, true, false, true>(gy, w, gx, N, H, W, s, BnStat{}, br);
}

extern "C" size_t alignq_conv3x3_ws_bytes(int C) {
  // partials of at most 2 CTAs per SM: [groups * ctas][C][taps * C] floats = ctas_total * C * 9 * C / groups... <= 2 * 148 * 9 C^2
  return (size_t)2 * ALIGNQ_NUM_SMS * 9 * C * C * sizeof(float) + (size_t)8 * ALIGNQ_NUM_SMS * sizeof(float);
}

extern "C" int alignq_conv3x3_bwd_weight(const float* x, const float* gy, float* gw, int N, int H, int W, int C, int mode,
                                         int accumulate, void* ws, size_t ws_bytes, alignq_stream_t stream) {
  int rc = conv_args_ok(x, gy, gw, N, H, W, C, mode);
  if (rc) return rc;
  if (!ws || !aligned16(ws)) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* wsf = reinterpret_cast<float*>(ws);
  // bf16 terms per value: 2 (16 bits, >= tf32) or 3 (24 bits); taps per CTA so that main + cross accumulators fit TMEM
  // TF32 mode: ONE fp16 term per value (11 significand bits = tf32's; gy scaled per CTA by a power of two), one MMA per
  // k-step -- a third of the tensor work and half the staging of the two-term bf16 split this mode used before
  // (ALIGNQ_WGRAD_TERMS=2 brings that back); TF32X3 mode: three bf16 terms (24 bits), six MMAs.
  if (mode == ALIGNQ_CONV_TF32 && env_int("ALIGNQ_WGRAD_TERMS", 1) == 1) {
    if (C == 16) return launch_wgrad<16, 1, 9, false, 2>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
    if (C == 32) return launch_wgrad<32, 1, 9, false, 2>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
    return launch_wgrad<64, 1, 3, false>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  }
  if (mode == ALIGNQ_CONV_TF32) {
    if (C == 16) return launch_wgrad<16, 2, 9, false>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
    if (C == 32) return launch_wgrad<32, 2, 3, false>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
    return launch_wgrad<64, 2, 3, false>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  }
  if (C == 16) return launch_wgrad<16, 3, 9, true>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  if (C == 32) return launch_wgrad<32, 3, 3, true>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  return launch_wgrad<64, 3, 3, true>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
}

#ifdef ALIGNQ_CONV_TRACE
extern "C" int alignq_conv_trace_reset(void) {
  void* p = nullptr;
  cudaError_t e = cudaGetSymbolAddress(&p, g_conv_trace);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(p, 0, sizeof(unsigned long long) * TRACE_CTAS * TRACE_SLOTS);
}
extern "C" int alignq_conv_trace_read(void* host_dst, size_t bytes) {
  if (bytes > sizeof(unsigned long long) * TRACE_CTAS * TRACE_SLOTS) return ALIGNQ_EINVAL;
  return (int)cudaMemcpyFromSymbol(host_dst, g_conv_trace, bytes);
}
#endif
