// First-layer ("stem") 3x3 convolution of the CIFAR-style networks -- SURVEY.md 8(f) item 2, the Cin = 3 case:
//   F.conv2d(image, weight_q, None, stride 1, padding 1)        cdf_alignment/resnet-20-cifar-10/model/quantization.py:116-120
//   as called by  self.conv0 = Conv2d(3, 16, kernel_size=3, stride=1, padding=1, bias=False)   .../model/resnet.py:92
// forward (optionally with the batch statistics of the following BatchNorm in the epilogue) and weight gradient, NHWC fp32.
// (No data gradient: the image does not require one; callers that want it use cuDNN.)
//
// Why not a tensor-core kernel: K = 27 and the whole layer is 57 MFLOP on 10 MB of traffic -- it is bound by HBM and by
// launch latency, not by math, so it is a direct fp32 FFMA convolution (exact fp32 products, no tf32 rounding).  What it
// replaces in the step (B = 128, 32 x 32; CUPTI timeline of one graph replay): cuDNN's forward for this shape is an NCHW
// kernel wrapped in two layout-conversion launches (9.2 + 7.8 us) followed by the statistics launch (9.1 us), and its weight
// gradient (`wgrad_alg0_engine_NHWC`, 40 us) is the exposed tail of the whole backward pass: nothing is left to overlap it.
#include "common.cuh"
#include "bn_stat.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace stem {

constexpr int CIN = 3, TAPS = 9, KW = TAPS * CIN;       // 27 weights per output channel
constexpr int NT = 256, TW = 32, IW = TW + 2;

// ---- forward: one 16 x 32 pixel tile per CTA, thread = column tx, rows ty and ty + 8, all COUT channels ------------
constexpr int FTH = 16, FIH = FTH + 2;

template <int COUT, bool STATS>
__global__ void __launch_bounds__(NT)
stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, int H, int W, int tiles_x,
                int tiles_y, BnStat bs) {
  __shared__ __align__(16) float wsm[KW * COUT];        // [tap][ci][co]
  __shared__ float xin[FIH * IW * CIN];
  __shared__ double red[STATS ? (NT / 32) * 2 * COUT : 1];
  const int tile = blockIdx.x;
  const int n = tile / (tiles_x * tiles_y), ty0 = ((tile / tiles_x) % tiles_y) * FTH, tx0 = (tile % tiles_x) * TW;
  for (int i = threadIdx.x; i < KW * COUT; i += NT) {   // w is [co][kh][kw][ci] (channels_last)
    const int co = i % COUT, k = i / COUT;
    wsm[i] = w[co * KW + k];
  }
  for (int i = threadIdx.x; i < FIH * IW * CIN; i += NT) {
    const int r = i / (IW * CIN), rem = i % (IW * CIN), col = rem / CIN, ci = rem % CIN;
    const int yy = ty0 - 1 + r, xx = tx0 - 1 + col;
    xin[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[(((int64_t)n * H + yy) * W + xx) * CIN + ci] : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float a0[COUT], a1[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
  const float4* w4 = reinterpret_cast<const float4*>(wsm);
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float v0 = xin[((ty + kh) * IW + tx + kw) * CIN + ci];
        const float v1 = xin[((ty + 8 + kh) * IW + tx + kw) * CIN + ci];
        const int k = (kh * 3 + kw) * CIN + ci;
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 wv = w4[k * (COUT / 4) + c4];
          a0[4 * c4 + 0] = fmaf(v0, wv.x, a0[4 * c4 + 0]); a0[4 * c4 + 1] = fmaf(v0, wv.y, a0[4 * c4 + 1]);
          a0[4 * c4 + 2] = fmaf(v0, wv.z, a0[4 * c4 + 2]); a0[4 * c4 + 3] = fmaf(v0, wv.w, a0[4 * c4 + 3]);
          a1[4 * c4 + 0] = fmaf(v1, wv.x, a1[4 * c4 + 0]); a1[4 * c4 + 1] = fmaf(v1, wv.y, a1[4 * c4 + 1]);
          a1[4 * c4 + 2] = fmaf(v1, wv.z, a1[4 * c4 + 2]); a1[4 * c4 + 3] = fmaf(v1, wv.w, a1[4 * c4 + 3]);
        }
      }
  const int ox = tx0 + tx, oy0 = ty0 + ty, oy1 = ty0 + ty + 8;
  const bool ok0 = ox < W && oy0 < H, ok1 = ox < W && oy1 < H;
  if (ok0) {
    float4* o = reinterpret_cast<float4*>(y + (((int64_t)n * H + oy0) * W + ox) * COUT);
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) o[c4] = make_float4(a0[4 * c4], a0[4 * c4 + 1], a0[4 * c4 + 2], a0[4 * c4 + 3]);
  }
  if (ok1) {
    float4* o = reinterpret_cast<float4*>(y + (((int64_t)n * H + oy1) * W + ox) * COUT);
#pragma unroll
    for (int c4 = 0; c4 < COUT / 4; ++c4) o[c4] = make_float4(a1[4 * c4], a1[4 * c4 + 1], a1[4 * c4 + 2], a1[4 * c4 + 3]);
  }
  if (STATS) {
    float s[COUT], ss[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
      const float u = ok0 ? a0[c] : 0.f, v = ok1 ? a1[c] : 0.f;
      s[c] = u + v;
      ss[c] = fmaf(u, u, v * v);
    }
    bn_stat_cta_finish<COUT, NT>(s, ss, red, bs);
  }
}

// ---- weight gradient: dW[co][kh][kw][ci] = sum_{n,y,x} gy[n,y,x,co] x[n,y+kh-1,x+kw-1,ci] -------------------------------
// Persistent CTAs over 8 x 32 pixel tiles.  Thread = (channel quad cg, position lane pl): 4 x 27 accumulators in registers
// over every position the lane visits; per position one 128-bit shared load of gy and 27 broadcast loads of x feed 108
// FFMAs.  The next tile's global loads are issued before the current tile is computed (registers), so the loop never
// waits on HBM.  CTA totals go through shuffles and shared memory in a fixed order, then ONE fp64 atomic per value into a
// few accumulator copies (the order across CTAs is not fixed: ~1e-16 relative in fp64, invisible after rounding to fp32);
// the last CTA (ticket) writes dW and re-arms the accumulators, so no reduce launch and no memset follow.
constexpr int GTH = 8, GIH = GTH + 2, GPX = GTH * TW, WG_SLOTS = 4;
constexpr int XI = (GIH * IW * CIN + NT - 1) / NT;      // x-tile floats per thread (4)

template <int COUT>
struct WgFetch {
  static constexpr int GI = GPX * (COUT / 4) / NT;      // gy float4 per thread
  float4 g[GI];
  float xv[XI];
};

template <int COUT>
__device__ __forceinline__ void wg_fetch(const float* __restrict__ x, const float* __restrict__ gy, int H, int W, int tiles_x,
                                         int tiles_y, int tile, WgFetch<COUT>& f) {
  const int n = tile / (tiles_x * tiles_y), ty0 = ((tile / tiles_x) % tiles_y) * GTH, tx0 = (tile % tiles_x) * TW;
#pragma unroll
  for (int u = 0; u < WgFetch<COUT>::GI; ++u) {
    const int i = threadIdx.x + u * NT, p = i / (COUT / 4), c4 = i % (COUT / 4);
    const int yy = ty0 + p / TW, xx = tx0 + p % TW;
    f.g[u] = (yy < H && xx < W) ? __ldg(reinterpret_cast<const float4*>(gy + (((int64_t)n * H + yy) * W + xx) * COUT) + c4)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < XI; ++u) {
    const int i = threadIdx.x + u * NT;
    const int r = i / (IW * CIN), rem = i % (IW * CIN), col = rem / CIN, ci = rem % CIN;
    const int yy = ty0 - 1 + r, xx = tx0 - 1 + col;
    f.xv[u] = (i < GIH * IW * CIN && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (((int64_t)n * H + yy) * W + xx) * CIN + ci) : 0.f;
  }
}

template <int COUT>
__global__ void __launch_bounds__(NT, 1)
stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gw, int H, int W, int tiles_x,
                  int tiles_y, int ntiles, double* __restrict__ acc_ws, unsigned* __restrict__ counter) {
  constexpr int CG = COUT / 4, PL = NT / CG, PPT = GPX / PL;            // position lanes, positions per lane and tile
  constexpr int GYF = GPX * COUT;                                        // floats of the gy tile
  constexpr int REDF = (NT / 32) * CG * 4 * KW;                          // floats of the cross-warp reduction
  __shared__ __align__(16) float gys[GYF > REDF ? GYF : REDF];
  __shared__ float xin[XI * NT];
  __shared__ unsigned last_flag;
  const int cg = threadIdx.x % CG, pl = threadIdx.x / CG;
  float acc[4][KW];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KW; ++k) acc[j][k] = 0.f;

  WgFetch<COUT> f;
  int tile = blockIdx.x;
  if (tile < ntiles) wg_fetch<COUT>(x, gy, H, W, tiles_x, tiles_y, tile, f);
  for (; tile < ntiles; tile += gridDim.x) {
    __syncthreads();                                                     // previous tile fully consumed
#pragma unroll
    for (int u = 0; u < WgFetch<COUT>::GI; ++u) reinterpret_cast<float4*>(gys)[threadIdx.x + u * NT] = f.g[u];
#pragma unroll
    for (int u = 0; u < XI; ++u) xin[threadIdx.x + u * NT] = f.xv[u];
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) wg_fetch<COUT>(x, gy, H, W, tiles_x, tiles_y, tile + gridDim.x, f);
#pragma unroll 1
    for (int q = 0; q < PPT; ++q) {
      const int p = pl + q * PL, py = p / TW, px = p % TW;
      const float4 g = reinterpret_cast<const float4*>(gys)[p * CG + cg];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float v = xin[((py + kh) * IW + px + kw) * CIN + ci];
            const int k = (kh * 3 + kw) * CIN + ci;
            acc[0][k] = fmaf(g.x, v, acc[0][k]); acc[1][k] = fmaf(g.y, v, acc[1][k]);
            acc[2][k] = fmaf(g.z, v, acc[2][k]); acc[3][k] = fmaf(g.w, v, acc[3][k]);
          }
    }
  }
  // ---- CTA total: lanes of a warp that share cg (xor offsets CG .. 16), then the warps through shared memory ----------
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      float v = acc[j][k];
#pragma unroll
      for (int o = CG; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < CG) gys[(warp * CG + lane) * 4 * KW + j * KW + k] = v;  // lane == cg here (CG divides 32)
    }
  __syncthreads();
  for (int o = threadIdx.x; o < COUT * KW; o += NT) {                    // o = co * 27 + k  = (cg * 4 + j) * 27 + k
    const int g4 = o / (4 * KW), r = o % (4 * KW);
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < NT / 32; ++wv) v += (double)gys[(wv * CG + g4) * 4 * KW + r];
    atomicAdd(acc_ws + (size_t)(blockIdx.x % WG_SLOTS) * COUT * KW + o, v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_flag = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last_flag) {
    __threadfence();
    for (int o = threadIdx.x; o < COUT * KW; o += NT) {
      double v = 0.0;
#pragma unroll
      for (int sl = 0; sl < WG_SLOTS; ++sl) {
        double* a = acc_ws + (size_t)sl * COUT * KW + o;
        v += __ldcg(a);
        *a = 0.0;                                                        // re-arm
      }
      gw[o] = (float)v;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

static int stem_args_ok(const void* a, const void* b, const void* c, int N, int H, int W, int Cout) {
  if (!a || !b || !c || N < 1 || H < 1 || W < 1) return ALIGNQ_EINVAL;
  if (Cout != 16 && Cout != 32) return ALIGNQ_ERANGE;
  if ((int64_t)N * H * W > (int64_t)1 << 30) return ALIGNQ_ERANGE;
  return ALIGNQ_OK;
}

}  // namespace stem
}  // namespace alignq

using namespace alignq;

extern "C" size_t alignq_conv3x3_stem_ws_bytes(int Cout) {
  return (size_t)stem::WG_SLOTS * Cout * stem::KW * sizeof(double) + 16;
}

extern "C" int alignq_conv3x3_stem_fwd(const float* x, const float* w, float* y, int N, int H, int W, int Cout,
                                       float* running_mean, float* running_var, float momentum, float bn_eps,
                                       float* save_mean, float* save_invstd, double* bn_ws, uint32_t* bn_counter,
                                       int64_t* num_batches_tracked, alignq_stream_t stream) {
  int rc = stem::stem_args_ok(x, w, y, N, H, W, Cout);
  if (rc) return rc;
  if ((uintptr_t)y % 16) return ALIGNQ_EALIGN;
  const bool stats = save_mean != nullptr;
  if (stats && (!save_invstd || !bn_ws || !bn_counter)) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int tiles_x = (W + stem::TW - 1) / stem::TW, tiles_y = (H + stem::FTH - 1) / stem::FTH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;
  if (ntiles > 0x7fffffff) return ALIGNQ_ERANGE;
  BnStat bs{bn_ws, bn_counter, running_mean, running_var, save_mean, save_invstd,
            reinterpret_cast<long long*>(num_batches_tracked), momentum, bn_eps, (double)N * H * W};
  if (Cout == 16) {
    if (stats) stem::stem_fwd_kernel<16, true><<<(int)ntiles, stem::NT, 0, s>>>(x, w, y, H, W, tiles_x, tiles_y, bs);
    else stem::stem_fwd_kernel<16, false><<<(int)ntiles, stem::NT, 0, s>>>(x, w, y, H, W, tiles_x, tiles_y, bs);
  } else {
    if (stats) stem::stem_fwd_kernel<32, true><<<(int)ntiles, stem::NT, 0, s>>>(x, w, y, H, W, tiles_x, tiles_y, bs);
    else stem::stem_fwd_kernel<32, false><<<(int)ntiles, stem::NT, 0, s>>>(x, w, y, H, W, tiles_x, tiles_y, bs);
  }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_conv3x3_stem_bwd_weight(const float* x, const float* gy, float* gw, int N, int H, int W, int Cout,
                                              void* ws, size_t ws_bytes, alignq_stream_t stream) {
  int rc = stem::stem_args_ok(x, gy, gw, N, H, W, Cout);
  if (rc) return rc;
  if (!ws) return ALIGNQ_EINVAL;
  if ((uintptr_t)gy % 16 || (uintptr_t)ws % 8) return ALIGNQ_EALIGN;
  if (ws_bytes < alignq_conv3x3_stem_ws_bytes(Cout)) return ALIGNQ_ENOSPACE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int tiles_x = (W + stem::TW - 1) / stem::TW, tiles_y = (H + stem::GTH - 1) / stem::GTH;
  const int64_t ntiles = (int64_t)N * tiles_x * tiles_y;
  if (ntiles > 0x7fffffff) return ALIGNQ_ERANGE;
  double* acc = reinterpret_cast<double*>(ws);                   // ZERO before first use; re-armed by the kernel
  unsigned* counter = reinterpret_cast<unsigned*>(acc + (size_t)stem::WG_SLOTS * Cout * stem::KW);
  const int grid = (int)(ntiles < ALIGNQ_NUM_SMS ? ntiles : ALIGNQ_NUM_SMS);
  if (Cout == 16) stem::stem_wgrad_kernel<16><<<grid, stem::NT, 0, s>>>(x, gy, gw, H, W, tiles_x, tiles_y, (int)ntiles, acc, counter);
  else stem::stem_wgrad_kernel<32><<<grid, stem::NT, 0, s>>>(x, gy, gw, H, W, tiles_x, tiles_y, (int)ntiles, acc, counter);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
