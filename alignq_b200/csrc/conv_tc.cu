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
// bytes: nine shifted views of ONE staged tile, no data duplication.  Pad positions (xx >= W or yy >= H) produce junk
// output rows that are never stored (forward) or carry zero gy (weight gradient).
//
// Numerics (mode): ALIGNQ_CONV_TF32 = one kind::tf32 MMA per k-step on operands rounded to tf32 (what cuDNN runs by
// default under torch.backends.cudnn.allow_tf32 = True, the reference's own GPU path; ~5e-4 relative); ALIGNQ_CONV_TF32X3 =
// operands split v = H + L (H = top 19 bits), three MMAs (H H into the main accumulator, H L + L H into a second one:
// the tensor core truncates when it adds into the fp32 accumulator), fp32-level parity with F.conv2d (tested to 1e-5).
#include "common.cuh"
#include "tc_common.cuh"
#include "tc_small_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace ctc {

using namespace tc;

constexpr int NT = 256;                 // threads per CTA: 8 warps (warps w and w+4 share TMEM lane quarter w)

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Stage `npt` positions starting at global position g0 of an NHWC tensor into planes.  AS_OUTPUT: position g is the
// OUTPUT pixel (yy, xx) (valid for yy < H, xx < W); otherwise the zero-padded INPUT pixel (yy - 1, xx - 1).
// NS = 1: one tile of tf32-rounded values; NS = 2: H tile and, `lo_off` bytes further, the L = v - H tile.
template <int C, int NS, bool AS_OUTPUT>
__device__ __forceinline__ void stage_planes(const float* __restrict__ src, const Geo& G, int g0, int npt, uint8_t* planes,
                                             int PS, int lo_off) {
  constexpr int C4 = C / 4, STEP = NT / C4;
  const int c4 = threadIdx.x % C4;
  int j = threadIdx.x / C4;
  int g = g0 + j;
  int n = g / G.per;
  int q = g - n * G.per;
  int yy = q / G.Wp;
  int xx = q - yy * G.Wp;
  uint8_t* dst = planes + c4 * PS;
  constexpr int U = 4;
  for (; j < npt; j += U * STEP) {
    float4 v[U];
    int jj[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      jj[u] = j + u * STEP;
      const int iy = AS_OUTPUT ? yy : yy - 1, ix = AS_OUTPUT ? xx : xx - 1;
      const bool ok = jj[u] < npt && n < G.N && iy >= 0 && iy < G.H && ix >= 0 && ix < G.W;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v[u] = __ldg(reinterpret_cast<const float4*>(src + ((int64_t)(n * G.H + iy) * G.W + ix) * C + 4 * c4));
      xx += STEP;                                            // advance the decoded position by STEP
      while (xx >= G.Wp) { xx -= G.Wp; ++yy; }
      while (yy >= G.Hp) { yy -= G.Hp; ++n; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (jj[u] >= npt) continue;
      uint8_t* d = dst + jj[u] * 16;
      if (NS == 1) {
        *reinterpret_cast<uint4*>(d) = make_uint4(cvt_tf32(v[u].x), cvt_tf32(v[u].y), cvt_tf32(v[u].z), cvt_tf32(v[u].w));
      } else {
        const uint4 h = make_uint4(__float_as_uint(v[u].x) & 0xFFFFE000u, __float_as_uint(v[u].y) & 0xFFFFE000u,
                                   __float_as_uint(v[u].z) & 0xFFFFE000u, __float_as_uint(v[u].w) & 0xFFFFE000u);
        *reinterpret_cast<uint4*>(d) = h;
        *reinterpret_cast<float4*>(d + lo_off) = make_float4(v[u].x - __uint_as_float(h.x), v[u].y - __uint_as_float(h.y),
                                                             v[u].z - __uint_as_float(h.z), v[u].w - __uint_as_float(h.w));
      }
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
template <int C>
struct FwdCfg {
  static constexpr int MT = 256;                                   // output positions per tile (two M = 128 MMAs)
  static constexpr int W_TAP = C * C * 4;                          // one tap's weight tile, dense K-major
};

template <int C, int NS, bool FLIP>
__global__ void __launch_bounds__(NT)
conv3x3_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out, Geo G, int ntiles,
                   int npt, int PS) {
  using F = FwdCfg<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int W_BYTES = 9 * F::W_TAP;
  uint8_t* wsm = smem;                                             // [NS][9][C x C]
  uint8_t* planes = smem + NS * W_BYTES;                           // [NS][C/4 planes][npt positions]
  const int tile_bytes = (C / 4) * PS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(planes + NS * tile_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ACC = (NS == 2) ? 2 * C : C;                       // columns per M tile
  constexpr int TCOLS = (2 * ACC < 32) ? 32 : 2 * ACC;             // two M tiles; power of two >= 32

  // ---- weights -> shared memory (K-major N x K tiles per tap), once per CTA ---------------------------------------
  // B-operand element (row r = N index, k): (r / 8) * SBO_W + (r % 8) * 16 + (k / 4) * 128 + (k % 4) * 4, SBO_W = (C/4) * 128
  constexpr int SBO_W = (C / 4) * 128;
  for (int idx = threadIdx.x; idx < 9 * C * (C / 4); idx += NT) {
    const int k4 = idx % (C / 4), r = (idx / (C / 4)) % C, tap = idx / (C * (C / 4));
    const int kh = tap / 3, kw = tap % 3;
    float v[4];
    if (!FLIP) {                                                   // row = co, k = ci: 4 consecutive ci of w[co][kh][kw][:]
      const float4 t = __ldg(reinterpret_cast<const float4*>(w + ((size_t)(r * 3 + kh) * 3 + kw) * C + 4 * k4));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {                                                       // row = ci, k = co: w[co][2-kh][2-kw][ci] for 4 consecutive co
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = __ldg(w + ((size_t)((4 * k4 + e) * 3 + (2 - kh)) * 3 + (2 - kw)) * C + r);
    }
    uint8_t* d = wsm + tap * F::W_TAP + (r >> 3) * SBO_W + (r & 7) * 16 + k4 * 128;
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
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t IDESC = make_idesc(2u /*tf32*/, 128u, (uint32_t)C);

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int g0 = tile * F::MT;
    // ---- 1. stage the input positions g0 .. g0 + MT + 2 Wp + 2 --------------------------------------------------
    stage_planes<C, NS, false>(in, G, g0, npt, planes, PS, tile_bytes);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- 2. MMAs ---------------------------------------------------------------------------------------------------
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sa = smem_u32(planes), sw = smem_u32(wsm);
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t d_main = tmem_base + mt * ACC, d_cross = d_main + C;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t shift = (uint32_t)((tap / 3) * G.Wp + (tap % 3) + mt * 128) * 16u;
#pragma unroll
          for (int ks = 0; ks < C / 8; ++ks) {
            const uint32_t acc = (tap > 0 || ks > 0) ? 1u : 0u;
            const uint64_t ah = make_desc(sa + shift + ks * 2 * PS, PS, 128);
            const uint64_t bh = make_desc(sw + tap * F::W_TAP + ks * 2 * 128, 128, SBO_W);
            umma<true>(d_main, ah, bh, IDESC, acc);
            if (NS == 2) {
              const uint64_t al = make_desc(sa + tile_bytes + shift + ks * 2 * PS, PS, 128);
              const uint64_t bl = make_desc(sw + W_BYTES + tap * F::W_TAP + ks * 2 * 128, 128, SBO_W);
              umma<true>(d_cross, ah, bl, IDESC, acc);
              umma<true>(d_cross, al, bh, IDESC, 1u);
            }
          }
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, (uint32_t)(it & 1));
    tc_fence_after();
    // ---- 3. epilogue: warp w -> M tile w / 4, TMEM lanes 32 (w % 4) .. +31; one output position per thread ---------
    {
      const int mt = warp >> 2;
      const int j = mt * 128 + (warp & 3) * 32 + lane;             // position inside the tile
      const int g = g0 + j;
      const int n = g / G.per;
      const int q = g - n * G.per;
      const int yy = q / G.Wp, xx = q - yy * G.Wp;
      const bool ok = g < G.npos && yy < G.H && xx < G.W;
      float* o = out + ((int64_t)(n * G.H + yy) * G.W + xx) * C;
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + mt * ACC;
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(ta + c0, v);
        float r[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __uint_as_float(v[e]);
        if (NS == 2) {
          uint32_t x2[16];
          tmem_ld16(ta + C + c0, x2);
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] += __uint_as_float(x2[e]);
        }
        if (ok) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(o + c0 + e) = make_float4(r[e], r[e + 1], r[e + 2], r[e + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                               // accumulators and planes are free again
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TCOLS);
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient:  dW[co][kh][kw][ci] = sum_pos gy[pos, co] * x[pos + kh Wp + kw, ci]       (split-K over CTAs)
// Both operands are MN-major here (MN = channels, K = positions).  kind::tf32 does not take MN-major operands in the
// un-swizzled layout (measured: the MMA returns zeros), so the operands are bf16 TERMS of the fp32 values, 8 channels
// per 16-byte unit, plane c8 = channels 8 c8 .. 8 c8 + 7:   byte address = plane_base + c8 * PS + position * 16.
//   NSB = 2 (mode TF32):   v = H + M, 16 mantissa bits (more than tf32's 11), products H H | H M + M H     (3 MMAs)
//   NSB = 3 (mode TF32X3): v = H + M + L, 24 bits, products H H | H M + M H + H L + L H + M M               (6 MMAs)
// (the large product in the main accumulator, the small ones in a second one: the tensor core truncates when it adds
// into the fp32 accumulator).  A = gy planes (M = 64 rows of which C are real), B = x planes with the start address
// moved by the tap shift, M = 64, N = C, K = 16 positions per MMA.  Tap t accumulates in TMEM columns t * C (+ TAPS * C
// for the cross terms); accumulator row m is TMEM lane (m % 16) + 32 (m / 16).  Every CTA writes its partial
// [C][TAPS * C] to the workspace; conv3x3_wgrad_reduce_kernel sums them in a fixed order.
template <int C>
struct WgCfg {
  static constexpr int KT = 256;                                   // positions (K) per tile
};

__host__ __device__ constexpr int plane_stride_bf16(int C, int npt) {
  int ps = npt * 16;
  const int want = (C == 16) ? 64 : (C == 32) ? 32 : 16;          // conflict-free 16-byte stores of one quarter-warp
  while ((ps % 128) != want) ps += 16;
  return ps;
}

// like stage_planes, but 8 channels per thread and position, converted to NSB bf16 terms (tiles `term_off` bytes apart)
template <int C, int NSB, bool AS_OUTPUT>
__device__ __forceinline__ void stage_planes_bf16(const float* __restrict__ src, const Geo& G, int g0, int npt,
                                                  uint8_t* planes, int PS, int term_off) {
  constexpr int C8 = C / 8, STEP = NT / C8;
  const int c8 = threadIdx.x % C8;
  int j = threadIdx.x / C8;
  int g = g0 + j;
  int n = g / G.per;
  int q = g - n * G.per;
  int yy = q / G.Wp;
  int xx = q - yy * G.Wp;
  uint8_t* dst = planes + c8 * PS;
  constexpr int U = 2;
  for (; j < npt; j += U * STEP) {
    float4 v[U][2];
    int jj[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      jj[u] = j + u * STEP;
      const int iy = AS_OUTPUT ? yy : yy - 1, ix = AS_OUTPUT ? xx : xx - 1;
      const bool ok = jj[u] < npt && n < G.N && iy >= 0 && iy < G.H && ix >= 0 && ix < G.W;
      v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float4* p = reinterpret_cast<const float4*>(src + ((int64_t)(n * G.H + iy) * G.W + ix) * C + 8 * c8);
        v[u][0] = __ldg(p);
        v[u][1] = __ldg(p + 1);
      }
      xx += STEP;
      while (xx >= G.Wp) { xx -= G.Wp; ++yy; }
      while (yy >= G.Hp) { yy -= G.Hp; ++n; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (jj[u] >= npt) continue;
      const float c[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
      tcsmall::store_chunk_n<NSB>(dst + jj[u] * 16, term_off, c);
    }
  }
}

template <int C, int NSB, int TAPS>
__global__ void __launch_bounds__(NT)
conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ partials, Geo G,
                     int ntiles, int npx, int PSX, int PSG) {
  using K = WgCfg<C>;
  extern __shared__ __align__(128) uint8_t smem[];
  // gy planes first: the M = 64 descriptor walks 8 MN units of PSG bytes; units >= C/8 read whatever follows (rows of D
  // that are never stored), which must stay inside this CTA's allocation -- the launcher sizes it accordingly
  uint8_t* gplanes = smem;                                         // [NSB][C/8][KT]
  const int gtile = (C / 8) * PSG, xtile = (C / 8) * PSX;
  uint8_t* xplanes = smem + NSB * gtile;                           // [NSB][C/8][npx]
  uint64_t* bar = reinterpret_cast<uint64_t*>(xplanes + NSB * xtile);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ACC = TAPS * C;                                    // columns of one accumulator set
  constexpr int NEED = 2 * ACC;                                    // main + cross
  constexpr int TCOLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
  static_assert(NEED <= 512, "weight-gradient accumulators exceed TMEM");
  const int tap0 = blockIdx.y * TAPS;

  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // bf16 operands, both MN-major (bits 15, 16)
  constexpr uint32_t IDESC = make_idesc(1u /*bf16*/, 64u, (uint32_t)C) | (1u << 15) | (1u << 16);

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int g0 = tile * K::KT;
    stage_planes_bf16<C, NSB, true>(gy, G, g0, K::KT, gplanes, PSG, gtile);     // zero at pad positions: they add nothing
    stage_planes_bf16<C, NSB, false>(x, G, g0, npx, xplanes, PSX, xtile);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t sg = smem_u32(gplanes), sx = smem_u32(xplanes);
#pragma unroll 1
      for (int t = 0; t < TAPS; ++t) {
        const int tap = tap0 + t;
        const uint32_t shift = (uint32_t)((tap / 3) * G.Wp + (tap % 3)) * 16u;
        const uint32_t d_main = tmem_base + t * C, d_cross = tmem_base + ACC + t * C;
#pragma unroll 2
        for (int ks = 0; ks < K::KT / 16; ++ks) {
          const uint32_t acc = (it > 0 || ks > 0) ? 1u : 0u;
          // MN-major: LBO = stride of the 8-position K groups (128 B: K is linear in the position), SBO = plane stride
          const uint32_t ga = sg + ks * 256, xa = sx + shift + ks * 256;
          const uint64_t ah = make_desc(ga, 128, PSG), bh = make_desc(xa, 128, PSX);
          const uint64_t am = make_desc(ga + gtile, 128, PSG), bm = make_desc(xa + xtile, 128, PSX);
          umma<false>(d_main, ah, bh, IDESC, acc);                           // H H
          umma<false>(d_cross, ah, bm, IDESC, acc);                          // the small products
          umma<false>(d_cross, am, bh, IDESC, 1u);
          if (NSB == 3) {
            const uint64_t al = make_desc(ga + 2 * gtile, 128, PSG), bl = make_desc(xa + 2 * xtile, 128, PSX);
            umma<false>(d_cross, ah, bl, IDESC, 1u);
            umma<false>(d_cross, al, bh, IDESC, 1u);
            umma<false>(d_cross, am, bm, IDESC, 1u);
          }
        }
      }
      umma_commit(bar);
    }
    mbar_wait(bar, (uint32_t)(it & 1));                            // the planes are rewritten by the next tile
    tc_fence_after();
    tc_fence_before();
    __syncthreads();
  }
  // ---- epilogue: accumulator row co = 16 w + l sits in lane l of warp w's quarter (w < C / 16, l < 16) ------------
  // (all 32 lanes of a warp must take part in tcgen05.ld: the store is predicated instead)
  if (warp < 4) {
    const bool mine = warp < C / 16 && lane < 16;
    const int co = 16 * warp + lane;
    float* o = partials + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * C + (mine ? co : 0)) * ACC;
    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < ACC; c0 += 16) {
      uint32_t v[16], x2[16];
      float r[16];
      if (it > 0) {
        tmem_ld16(ta + c0, v);
        tmem_ld16(ta + ACC + c0, x2);
#pragma unroll
        for (int e = 0; e < 16; ++e) r[e] = __uint_as_float(v[e]) + __uint_as_float(x2[e]);
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
}

// gw[co][kh][kw][ci] (+)= sum over the CTAs' partials [grp][cta][co][TAPS * C], fixed order
__global__ void __launch_bounds__(256)
conv3x3_wgrad_reduce_kernel(const float* __restrict__ partials, int nparts, int C, int taps, float* __restrict__ gw,
                            int accumulate) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;            // index into [co][9][ci]
  if (e >= C * 9 * C) return;
  const int ci = e % C, tap = (e / C) % 9, co = e / (9 * C);
  const int grp = tap / taps, t = tap - grp * taps;
  const int acc = taps * C;
  const float* p = partials + ((size_t)grp * nparts * C + co) * acc + t * C + ci;
  float s = 0.f;
  for (int k = 0; k < nparts; ++k) s += p[(size_t)k * C * acc];
  gw[e] = accumulate ? gw[e] + s : s;
}

inline Geo make_geo(int N, int H, int W) {
  Geo G;
  G.N = N; G.H = H; G.W = W; G.Hp = H + 2; G.Wp = W + 2;
  G.per = G.Hp * G.Wp;
  G.npos = N * G.per;
  return G;
}

template <int C, int NS, bool FLIP>
static int launch_fwd(const float* in, const float* w, float* out, int N, int H, int W, cudaStream_t s) {
  using F = FwdCfg<C>;
  const Geo G = make_geo(N, H, W);
  const int ntiles = (G.npos + F::MT - 1) / F::MT;
  const int npt = F::MT + 2 * G.Wp + 2;
  const int PS = plane_stride(C, npt);
  const size_t smem = (size_t)NS * 9 * F::W_TAP + (size_t)NS * (C / 4) * PS + 64;
  if (smem > 227 * 1024) return ALIGNQ_ERANGE;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_fwd_kernel<C, NS, FLIP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = ntiles < ALIGNQ_NUM_SMS * per_sm ? ntiles : ALIGNQ_NUM_SMS * per_sm;
  conv3x3_fwd_kernel<C, NS, FLIP><<<grid, NT, smem, s>>>(in, w, out, G, ntiles, npt, PS);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

template <int C, int NSB, int TAPS>
static int launch_wgrad(const float* x, const float* gy, float* gw, int N, int H, int W, int accumulate, float* ws,
                        size_t ws_bytes, cudaStream_t s) {
  using K = WgCfg<C>;
  const Geo G = make_geo(N, H, W);
  const int ntiles = (G.npos + K::KT - 1) / K::KT;
  const int npx = K::KT + 2 * G.Wp + 2;
  const int PSX = plane_stride_bf16(C, npx), PSG = plane_stride_bf16(C, K::KT);
  size_t smem = (size_t)NSB * (C / 8) * PSG + (size_t)NSB * (C / 8) * PSX + 64;
  const size_t reach = (size_t)(NSB - 1) * (C / 8) * PSG + 8 * (size_t)PSG;   // what the M = 64 descriptors may touch
  if (smem < reach + 64) smem = reach + 64;
  if (smem > 227 * 1024) return ALIGNQ_ERANGE;
  const int groups = 9 / TAPS;
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  int grid = ALIGNQ_NUM_SMS * per_sm / groups;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  const size_t need = (size_t)groups * grid * C * TAPS * C * sizeof(float);
  if (ws_bytes < need) return ALIGNQ_ENOSPACE;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<C, NSB, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  conv3x3_wgrad_kernel<C, NSB, TAPS><<<dim3(grid, groups), NT, smem, s>>>(x, gy, ws, G, ntiles, npx, PSX, PSG);
  ALIGNQ_LAUNCH_CHECK();
  conv3x3_wgrad_reduce_kernel<<<(C * 9 * C + 255) / 256, 256, 0, s>>>(ws, grid, C, TAPS, gw, accumulate);
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
    else { if ((NS_) == 1) { CALL(64, 1); } else { CALL(64, 2); } } \
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

extern "C" size_t alignq_conv3x3_ws_bytes(int C) {
  // partials of at most 2 CTAs per SM: [groups * ctas][C][taps * C] floats = ctas_total * C * 9 * C / groups... <= 2 * 148 * 9 C^2
  return (size_t)2 * ALIGNQ_NUM_SMS * 9 * C * C * sizeof(float);
}

extern "C" int alignq_conv3x3_bwd_weight(const float* x, const float* gy, float* gw, int N, int H, int W, int C, int mode,
                                         int accumulate, void* ws, size_t ws_bytes, alignq_stream_t stream) {
  int rc = conv_args_ok(x, gy, gw, N, H, W, C, mode);
  if (rc) return rc;
  if (!ws || !aligned16(ws)) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* wsf = reinterpret_cast<float*>(ws);
  // bf16 terms per value: 2 (16 bits, >= tf32) or 3 (24 bits); taps per CTA so that main + cross accumulators fit TMEM
  if (mode == ALIGNQ_CONV_TF32) {
    if (C == 16) return launch_wgrad<16, 2, 9>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
    if (C == 32) return launch_wgrad<32, 2, 3>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
    return launch_wgrad<64, 2, 3>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  }
  if (C == 16) return launch_wgrad<16, 3, 9>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  if (C == 32) return launch_wgrad<32, 3, 3>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
  return launch_wgrad<64, 3, 3>(x, gy, gw, N, H, W, accumulate, wsf, ws_bytes, s);
}
