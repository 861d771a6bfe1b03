// fp32 (CUDA-core FFMA) correlation / Gram kernels: the bit-conservative numerics mode
// (ALIGNQ_GRAM_FP32) of the ADMM correlation-preservation term, and the checker the tcgen05
// modes are validated against on the GPU.
//
// Replaces corr() (cdf_alignment_admm/resnet-56-cifar-10/model/quantization.py:134-137, eps = 0;
// cdf_alignment_admm/dann_office/model/quantization.py:158-161, eps = 1e-5) and, fused with the
// activation map, the body of activation_quantize_fn.forward (quantization.py:109-123): one pass
// over x yields y, corr(x, x) and corr(t, t) partials.  Backward follows SURVEY.md Appendix A.4.
//
// Layout: x is [B, F] row-major (NCHW activation viewed as [B, C*H*W]).  A CTA owns a slab of
// feature columns and walks it in tiles of KT = 32 columns; all B rows of a tile sit in shared
// memory, so the per-column batch statistics never leave the SM.  Split-K partials
// [which][slab][B][B] go to the caller's workspace and are reduced by gram_reduce_kernel
// (deterministic, no atomics).
#include "common.cuh"
#include "gram_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {

constexpr int KT = 32;           // feature columns per tile
constexpr int GT = 256;          // threads per CTA

// Load one [B x KT] tile of src (columns f0..f0+KT) into S[k*LD + i], zero-filling rows >= B (up to
// Bpad) and columns >= F.  TRANSFORM applies the activation map; y (nullable) receives the quantized
// activation for the loaded elements.
template <bool TRANSFORM>
__device__ __forceinline__ void load_tile_T(const float* __restrict__ src, int B, int Bpad, int64_t F, int64_t f0,
                                            float* __restrict__ S, int LD, ActQ q, float* __restrict__ y) {
  const bool vec = ((F & 3) == 0) && aligned16(src) && (!y || aligned16(y));
  const int k4 = threadIdx.x & 7;
  for (int r = threadIdx.x >> 3; r < Bpad; r += GT / 8) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t f = f0 + 4 * k4;
    if (r < B) {
      const float* p = src + (int64_t)r * F + f;
      if (vec && f + 3 < F) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (f + c < F) v[c] = p[c];
      }
      if (TRANSFORM) {
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { v[c] = act_map_t(v[c], q.ar); o[c] = act_quant_from_t(v[c], q); }
        if (y) {
          float* py = y + (int64_t)r * F + f;
          if (vec && f + 3 < F) *reinterpret_cast<float4*>(py) = make_float4(o[0], o[1], o[2], o[3]);
          else {
#pragma unroll
            for (int c = 0; c < 4; ++c) if (f + c < F) py[c] = o[c];
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) if (f + c >= F) v[c] = 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) S[(4 * k4 + c) * LD + r] = v[c];
  }
}

// Per-column batch statistics of the tile in S (two-pass, fp32): mean_k and std_k (unbiased).
// 8 threads per column.  Results to smem arrays mu[KT], sd[KT].
__device__ __forceinline__ void tile_col_stats(const float* __restrict__ S, int LD, int B,
                                               float* __restrict__ mu, float* __restrict__ sd) {
  const int k = threadIdx.x >> 3, l = threadIdx.x & 7;
  const float* col = S + k * LD;
  float s = 0.f;
  for (int i = l; i < B; i += 8) s += col[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float m = s / (float)B;
  float ss = 0.f;
  for (int i = l; i < B; i += 8) { const float d = col[i] - m; ss = fmaf(d, d, ss); }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  if (l == 0) { mu[k] = m; sd[k] = sqrtf(ss / (float)(B - 1)); }
}

// S[k][i] <- (S[k][i] - mu_k) / (sd_k + eps) for i < B; columns >= ncols_valid are zeroed.
__device__ __forceinline__ void tile_standardise(float* __restrict__ S, int LD, int B, int ncols_valid,
                                                 const float* __restrict__ mu, const float* __restrict__ sd, float eps) {
  for (int e = threadIdx.x; e < KT * B; e += GT) {
    const int k = e / B, i = e - k * B;
    float v = 0.f;
    if (k < ncols_valid) v = __fdiv_rn(__fsub_rn(S[k * LD + i], mu[k]), __fadd_rn(sd[k], eps));
    S[k * LD + i] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Forward: split-K Gram partials.
//   grid = (nslabs, nb*nb, nz)   z = 0: corr(xa, xb) operand pair as given
//                                 z = 1: the activation-mapped operand t (fused mode, SAME only)
//   TB = output block edge (32 / 64 / 128), nb = Bpad / TB.
template <int TB, bool SAME>
__global__ void __launch_bounds__(GT)
gram_ffma_fwd_kernel(const float* __restrict__ xa, const float* __restrict__ xb, int B, int Bpad, int64_t F,
                     float eps, int tiles_per_slab, int fused, ActQ q, float* __restrict__ y,
                     float* __restrict__ partials) {
  extern __shared__ __align__(16) float smem[];
  constexpr int MT = TB / 16;
  const int LD = Bpad + 4;
  float* SA = smem;
  float* SB = SAME ? SA : SA + KT * LD;
  float* mu = SB + KT * LD;
  float* sd = mu + KT;

  const int nb = Bpad / TB;
  const int bi = blockIdx.y / nb, bj = blockIdx.y % nb;
  const int z = blockIdx.z;
  const bool transform = fused && z == 1;
  const bool write_y = transform && y != nullptr && blockIdx.y == 0;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  float acc[MT][MT];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < MT; ++b) acc[a][b] = 0.f;

  const int64_t ntiles = (F + KT - 1) / KT;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_slab;
  const int64_t t1 = (t0 + tiles_per_slab < ntiles) ? t0 + tiles_per_slab : ntiles;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t f0 = t * KT;
    const int valid = (int)((F - f0 < KT) ? (F - f0) : KT);
    __syncthreads();
    if (transform) load_tile_T<true>(xa, B, Bpad, F, f0, SA, LD, q, write_y ? y : nullptr);
    else           load_tile_T<false>(xa, B, Bpad, F, f0, SA, LD, q, nullptr);
    __syncthreads();
    tile_col_stats(SA, LD, B, mu, sd);
    __syncthreads();
    tile_standardise(SA, LD, B, valid, mu, sd, eps);
    if (!SAME) {
      __syncthreads();
      load_tile_T<false>(xb, B, Bpad, F, f0, SB, LD, q, nullptr);
      __syncthreads();
      tile_col_stats(SB, LD, B, mu, sd);
      __syncthreads();
      tile_standardise(SB, LD, B, valid, mu, sd, eps);
    }
    __syncthreads();
    const float* pa = SA + bi * TB + ty * MT;
    const float* pb = SB + bj * TB + tx * MT;
#pragma unroll 4
    for (int k = 0; k < KT; ++k) {
      float a[MT], b[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) { a[m] = pa[k * LD + m]; b[m] = pb[k * LD + m]; }
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < MT; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
    }
  }
  float* out = partials + ((size_t)z * gridDim.x + blockIdx.x) * (size_t)B * B;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const int i = bi * TB + ty * MT + m;
    if (i >= B) continue;
#pragma unroll
    for (int n = 0; n < MT; ++n) {
      const int j = bj * TB + tx * MT + n;
      if (j < B) out[(size_t)i * B + j] = acc[m][n];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward (fused activation + ADMM term), one CTA per slab of columns, tiles of KT columns:
//   gS = W Xs / F with W = s (dLdD + dLdD^T), s = -gloss for the x path and +gloss for the t path,
//   then the standardise backward per column, and
//   gx = bx + (bt + gy) * 2 ar phi(x).
// smem: R[B][KTP] raw x, S[B][KTP] standardised operand, G[B][KTP] gS; row-major, KTP = KT + 1.
constexpr int KTP = KT + 1;

__device__ __forceinline__ void bwd_one_path(const float* __restrict__ W, int B, int Bp, float wscale, float invF, float eps,
                                             float* __restrict__ S, float* __restrict__ G, float* __restrict__ mu,
                                             float* __restrict__ sd, float* __restrict__ red, int valid) {
  // S holds the raw operand [B][KTP].  Column stats (8 threads per column).
  {
    const int k = threadIdx.x >> 3, l = threadIdx.x & 7;
    float s = 0.f;
    for (int i = l; i < B; i += 8) s += S[i * KTP + k];
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float m = s / (float)B;
    float ss = 0.f;
    for (int i = l; i < B; i += 8) { const float d = S[i * KTP + k] - m; ss = fmaf(d, d, ss); }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    if (l == 0) { mu[k] = m; sd[k] = sqrtf(ss / (float)(B - 1)); }
  }
  __syncthreads();
  // centre in place: S <- c = x - mu  (Xs = c / (sd + eps) applied on the fly)
  for (int e = threadIdx.x; e < B * KT; e += GT) {
    const int i = e / KT, k = e - i * KT;
    S[i * KTP + k] = (k < valid) ? (S[i * KTP + k] - mu[k]) : 0.f;
  }
  __syncthreads();
  // gS[i][k] = wscale * invF / (sd_k + eps) * sum_j Wsym[j][i] * c[j][k]   (Wsym = dLdD + dLdD^T, ld = Bp)
  // thread -> 4 columns (kq) x 4 rows (ig): 1 LDG.128 (L1-resident Wsym) + 4 LDS per 16 FMA
  {
    const int kq = threadIdx.x & 7, ig = threadIdx.x >> 3;
    float rs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) rs[c] = wscale * invF / (sd[4 * kq + c] + eps);
    for (int ib = 0; ib < Bp; ib += 128) {
      const int i = ib + 4 * ig;
      if (i >= Bp) continue;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
      for (int j = 0; j < B; ++j) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)j * Bp + i));
        const float* sp = S + j * KTP + 4 * kq;
        const float s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3];
        acc[0][0] = fmaf(w.x, s0, acc[0][0]); acc[0][1] = fmaf(w.x, s1, acc[0][1]); acc[0][2] = fmaf(w.x, s2, acc[0][2]); acc[0][3] = fmaf(w.x, s3, acc[0][3]);
        acc[1][0] = fmaf(w.y, s0, acc[1][0]); acc[1][1] = fmaf(w.y, s1, acc[1][1]); acc[1][2] = fmaf(w.y, s2, acc[1][2]); acc[1][3] = fmaf(w.y, s3, acc[1][3]);
        acc[2][0] = fmaf(w.z, s0, acc[2][0]); acc[2][1] = fmaf(w.z, s1, acc[2][1]); acc[2][2] = fmaf(w.z, s2, acc[2][2]); acc[2][3] = fmaf(w.z, s3, acc[2][3]);
        acc[3][0] = fmaf(w.w, s0, acc[3][0]); acc[3][1] = fmaf(w.w, s1, acc[3][1]); acc[3][2] = fmaf(w.w, s2, acc[3][2]); acc[3][3] = fmaf(w.w, s3, acc[3][3]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (i + r < B)
#pragma unroll
          for (int c = 0; c < 4; ++c) G[(i + r) * KTP + 4 * kq + c] = acc[r][c] * rs[c];
    }
  }
  __syncthreads();
  // per column: mean_i gS and sum_i gS*c  -> red[k], red[KT + k]
  {
    const int k = threadIdx.x >> 3, l = threadIdx.x & 7;
    float a = 0.f, b = 0.f;
    for (int i = l; i < B; i += 8) { const float g = G[i * KTP + k]; a += g; b = fmaf(g, S[i * KTP + k], b); }
    a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2); a += __shfl_xor_sync(0xffffffffu, a, 4);
    b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2); b += __shfl_xor_sync(0xffffffffu, b, 4);
    if (l == 0) { red[k] = a / (float)B; red[KT + k] = b; }
  }
  __syncthreads();
  // G <- (gS - mean gS)/(sd+eps) + dsd * c / ((B-1) sd),  dsd = -sum(gS c)/(sd+eps)^2
  for (int e = threadIdx.x; e < B * KT; e += GT) {
    const int i = e / KT, k = e - i * KT;
    const float se = sd[k] + eps;
    const float dsd = -red[KT + k] / (se * se);
    // torch's std backward masks the 0/0 of a constant column to 0 (std_backward: masked_fill_(result == 0, 0))
    const float through_sd = (sd[k] > 0.f) ? dsd * S[i * KTP + k] / ((float)(B - 1) * sd[k]) : 0.f;
    G[i * KTP + k] = (G[i * KTP + k] - red[k]) / se + through_sd;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(GT)
gram_ffma_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ Wsym,
                     const float* __restrict__ gloss, int B, int Bp, int64_t F, float ar, float eps, int tiles_per_slab,
                     float* __restrict__ gx) {
  extern __shared__ __align__(16) float smem[];
  float* R = smem;                 // raw x tile
  float* S = R + B * KTP;          // operand tile (x, then t)
  float* G = S + B * KTP;          // path gradient
  float* A = G + B * KTP;          // accumulated x-path gradient
  float* mu = A + B * KTP;
  float* sd = mu + KT;
  float* red = sd + KT;
  const float gl = gloss ? __ldg(gloss) : 1.0f;
  const float invF = 1.0f / (float)F;
  const float gscale = 2.0f * ar * kInvSqrt2Pi;
  const int64_t ntiles = (F + KT - 1) / KT;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_slab;
  const int64_t t1 = (t0 + tiles_per_slab < ntiles) ? t0 + tiles_per_slab : ntiles;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t f0 = t * KT;
    const int valid = (int)((F - f0 < KT) ? (F - f0) : KT);
    __syncthreads();
    for (int e = threadIdx.x; e < B * KT; e += GT) {
      const int i = e / KT, k = e - i * KT;
      const float v = (k < valid) ? x[(int64_t)i * F + f0 + k] : 0.f;
      R[i * KTP + k] = v;
      S[i * KTP + k] = v;
    }
    __syncthreads();
    bwd_one_path(Wsym, B, Bp, -gl, invF, eps, S, G, mu, sd, red, valid);         // corr_bwd(X, -dD)
    for (int e = threadIdx.x; e < B * KT; e += GT) {
      const int i = e / KT, k = e - i * KT;
      A[i * KTP + k] = G[i * KTP + k];
      S[i * KTP + k] = (k < valid) ? act_map_t(R[i * KTP + k], ar) : 0.f;
    }
    __syncthreads();
    bwd_one_path(Wsym, B, Bp, gl, invF, eps, S, G, mu, sd, red, valid);          // corr_bwd(T, +dD)
    for (int e = threadIdx.x; e < B * KT; e += GT) {
      const int i = e / KT, k = e - i * KT;
      if (k >= valid) continue;
      const int64_t idx = (int64_t)i * F + f0 + k;
      const float xv = R[i * KTP + k];
      const float v = __fmul_rn(xv, kInvSqrt2);
      const float dphi = gscale * gauss_kernel_from_v(v);
      const float up = G[i * KTP + k] + (gy ? gy[idx] : 0.f);
      gx[idx] = A[i * KTP + k] + up * dphi;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of the stand-alone corr(a, b) (quantization.py:134-137) with respect to ONE operand a, for an upstream
// dG:  gAs = W Bs / F  with  Wt[j][i] = W[i][j]  (W = dG for a = x, dG^T for a = y, dG + dG^T when x is y),
// then the standardise backward of a.  SAME: b aliases a (one tile, one set of statistics).
__device__ __forceinline__ void col_stats_rows(const float* __restrict__ S, int B, float* __restrict__ mu, float* __restrict__ sd) {
  const int k = threadIdx.x >> 3, l = threadIdx.x & 7;
  float s = 0.f;
  for (int i = l; i < B; i += 8) s += S[i * KTP + k];
  s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float m = s / (float)B;
  float ss = 0.f;
  for (int i = l; i < B; i += 8) { const float d = S[i * KTP + k] - m; ss = fmaf(d, d, ss); }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  if (l == 0) { mu[k] = m; sd[k] = sqrtf(ss / (float)(B - 1)); }
}

template <bool SAME>
__global__ void __launch_bounds__(GT)
corr_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ Wt, int B, int Bp,
                int64_t F, float eps, int tiles_per_slab, float* __restrict__ ga) {
  extern __shared__ __align__(16) float smem[];
  float* CA = smem;                                   // a tile: raw -> centred
  float* CB = SAME ? CA : CA + B * KTP;               // b tile: raw -> centred
  float* G = CB + B * KTP;
  float* muA = G + B * KTP;
  float* sdA = muA + KT;
  float* muB = sdA + KT;
  float* sdB = muB + KT;
  float* red = sdB + KT;                              // [2 KT]
  const float invF = 1.0f / (float)F;
  const int64_t ntiles = (F + KT - 1) / KT;
  const int64_t t0 = (int64_t)blockIdx.x * tiles_per_slab;
  const int64_t t1 = (t0 + tiles_per_slab < ntiles) ? t0 + tiles_per_slab : ntiles;
  for (int64_t t = t0; t < t1; ++t) {
    const int64_t f0 = t * KT;
    const int valid = (int)((F - f0 < KT) ? (F - f0) : KT);
    __syncthreads();
    for (int e = threadIdx.x; e < B * KT; e += GT) {
      const int i = e / KT, k = e - i * KT;
      CA[i * KTP + k] = (k < valid) ? a[(int64_t)i * F + f0 + k] : 0.f;
      if (!SAME) CB[i * KTP + k] = (k < valid) ? b[(int64_t)i * F + f0 + k] : 0.f;
    }
    __syncthreads();
    col_stats_rows(CA, B, muA, sdA);
    if (!SAME) col_stats_rows(CB, B, muB, sdB);
    __syncthreads();
    for (int e = threadIdx.x; e < B * KT; e += GT) {
      const int i = e / KT, k = e - i * KT;
      CA[i * KTP + k] = (k < valid) ? (CA[i * KTP + k] - muA[k]) : 0.f;
      if (!SAME) CB[i * KTP + k] = (k < valid) ? (CB[i * KTP + k] - muB[k]) : 0.f;
    }
    __syncthreads();
    const float* sdb = SAME ? sdA : sdB;
    {
      const int kq = threadIdx.x & 7, ig = threadIdx.x >> 3;
      float rs[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) rs[c] = invF / (sdb[4 * kq + c] + eps);
      for (int ib = 0; ib < Bp; ib += 128) {
        const int i = ib + 4 * ig;
        if (i >= Bp) continue;
        float acc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
        for (int j = 0; j < B; ++j) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)j * Bp + i));
          const float w[4] = {w4.x, w4.y, w4.z, w4.w};
          const float* sp = CB + j * KTP + 4 * kq;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(w[r], sp[c], acc[r][c]);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (i + r < B)
#pragma unroll
            for (int c = 0; c < 4; ++c) G[(i + r) * KTP + 4 * kq + c] = acc[r][c] * rs[c];
      }
    }
    __syncthreads();
    {
      const int k = threadIdx.x >> 3, l = threadIdx.x & 7;
      float u = 0.f, v = 0.f;
      for (int i = l; i < B; i += 8) { const float g = G[i * KTP + k]; u += g; v = fmaf(g, CA[i * KTP + k], v); }
      u += __shfl_xor_sync(0xffffffffu, u, 1); u += __shfl_xor_sync(0xffffffffu, u, 2); u += __shfl_xor_sync(0xffffffffu, u, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
      if (l == 0) { red[k] = u / (float)B; red[KT + k] = v; }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < B * KT; e += GT) {
      const int i = e / KT, k = e - i * KT;
      if (k >= valid) continue;
      const float se = sdA[k] + eps;
      const float dsd = -red[KT + k] / (se * se);
      const float through_sd = (sdA[k] > 0.f) ? dsd * CA[i * KTP + k] / ((float)(B - 1) * sdA[k]) : 0.f;   // torch masks std == 0
      ga[(int64_t)i * F + f0 + k] = (G[i * KTP + k] - red[k]) / se + through_sd;
    }
  }
}

// Wt[j][i] (leading dimension Bp, zero padded) from dG [B][B]: mode 0: dG[i][j] (transpose), 1: dG[j][i], 2: both added
__global__ void corr_wt_kernel(const float* __restrict__ dG, int B, int Bp, int mode, float* __restrict__ Wt) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * Bp) return;
  const int j = e / Bp, i = e - j * Bp;
  float v = 0.f;
  if (i < B) v = (mode == 0) ? dG[(size_t)i * B + j] : (mode == 1) ? dG[(size_t)j * B + i] : dG[(size_t)i * B + j] + dG[(size_t)j * B + i];
  Wt[e] = v;
}

int corr_ffma_backward(const float* x, const float* y, const float* dG, int B, int64_t F, float eps, float* gx, float* gy,
                       float* wt0, float* wt1, cudaStream_t s) {
  const int Bp = gram_bp(B);
  const bool same = (x == y);
  const int64_t ntiles = (F + KT - 1) / KT;
  int64_t nslabs = ntiles < 4 * ALIGNQ_NUM_SMS ? ntiles : 4 * ALIGNQ_NUM_SMS;
  const int tiles_per_slab = (int)((ntiles + nslabs - 1) / nslabs);
  nslabs = (ntiles + tiles_per_slab - 1) / tiles_per_slab;
  const size_t smem = ((size_t)(same ? 2 : 3) * B * KTP + 6 * KT) * sizeof(float);
  if (smem > 227 * 1024) return ALIGNQ_ERANGE;
  const int nb = (B * Bp + 255) / 256;
  if (same) {
    if (!gx) return ALIGNQ_OK;
    corr_wt_kernel<<<nb, 256, 0, s>>>(dG, B, Bp, 2, wt0);
    ALIGNQ_LAUNCH_CHECK();
    cudaError_t e = cudaFuncSetAttribute(corr_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    corr_bwd_kernel<true><<<(unsigned)nslabs, GT, smem, s>>>(x, x, wt0, B, Bp, F, eps, tiles_per_slab, gx);
    ALIGNQ_LAUNCH_CHECK();
    return ALIGNQ_OK;
  }
  cudaError_t e = cudaFuncSetAttribute(corr_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (gx) {                                     // gXs = dG Ys / F: Wt[j][i] = dG[i][j]
    corr_wt_kernel<<<nb, 256, 0, s>>>(dG, B, Bp, 0, wt0);
    ALIGNQ_LAUNCH_CHECK();
    corr_bwd_kernel<false><<<(unsigned)nslabs, GT, smem, s>>>(x, y, wt0, B, Bp, F, eps, tiles_per_slab, gx);
    ALIGNQ_LAUNCH_CHECK();
  }
  if (gy) {                                     // gYs = dG^T Xs / F: Wt[j][i] = dG[j][i]
    corr_wt_kernel<<<nb, 256, 0, s>>>(dG, B, Bp, 1, wt1);
    ALIGNQ_LAUNCH_CHECK();
    corr_bwd_kernel<false><<<(unsigned)nslabs, GT, smem, s>>>(y, x, wt1, B, Bp, F, eps, tiles_per_slab, gy);
    ALIGNQ_LAUNCH_CHECK();
  }
  return ALIGNQ_OK;
}

// ---------------------------------------------------------------------------------------------
static int pick_tb(int B) { return B <= 32 ? 32 : (B <= 64 ? 64 : 128); }

int gram_ffma_forward(const float* xa, const float* xb, int B, int64_t F, float eps, int fused, ActQ q, float* y,
                      float* partials, int* nslabs_out, size_t ws_bytes, cudaStream_t s) {
  const bool same = (xa == xb);
  if (fused && !same) return ALIGNQ_EINVAL;
  const int TB = pick_tb(B);
  const int Bpad = (B + TB - 1) / TB * TB;
  const int nb = Bpad / TB, nz = fused ? 2 : 1;
  const int64_t ntiles = (F + KT - 1) / KT;
  int64_t want = (2 * ALIGNQ_NUM_SMS) / (nb * nb * nz);
  if (want < 1) want = 1;
  int64_t nslabs = ntiles < want ? ntiles : want;
  const int64_t cap = gram_ws_slab_cap(ws_bytes, B);
  if (cap < 1) return ALIGNQ_ENOSPACE;
  if (nslabs > cap) nslabs = cap;
  const int tiles_per_slab = (int)((ntiles + nslabs - 1) / nslabs);
  nslabs = (ntiles + tiles_per_slab - 1) / tiles_per_slab;
  *nslabs_out = (int)nslabs;
  const int LD = Bpad + 4;
  const size_t smem = ((size_t)(same ? 1 : 2) * KT * LD + 2 * KT) * sizeof(float);
  dim3 grid((unsigned)nslabs, (unsigned)(nb * nb), (unsigned)nz);
#define GRAM_LAUNCH(TBV, SAMEV)                                                                              \
  do {                                                                                                       \
    cudaError_t e = cudaFuncSetAttribute(gram_ffma_fwd_kernel<TBV, SAMEV>,                                   \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e != cudaSuccess) return (int)e;                                                                     \
    gram_ffma_fwd_kernel<TBV, SAMEV><<<grid, GT, smem, s>>>(xa, xb, B, Bpad, F, eps, tiles_per_slab, fused,  \
                                                            q, y, partials);                                 \
  } while (0)
  if (TB == 32)      { if (same) GRAM_LAUNCH(32, true);  else GRAM_LAUNCH(32, false); }
  else if (TB == 64) { if (same) GRAM_LAUNCH(64, true);  else GRAM_LAUNCH(64, false); }
  else               { if (same) GRAM_LAUNCH(128, true); else GRAM_LAUNCH(128, false); }
#undef GRAM_LAUNCH
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

int gram_ffma_backward(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B,
                       int64_t F, float ar, float eps, float* gx, cudaStream_t s) {
  const int64_t ntiles = (F + KT - 1) / KT;
  int64_t nslabs = ntiles < 4 * ALIGNQ_NUM_SMS ? ntiles : 4 * ALIGNQ_NUM_SMS;
  const int tiles_per_slab = (int)((ntiles + nslabs - 1) / nslabs);
  nslabs = (ntiles + tiles_per_slab - 1) / tiles_per_slab;
  const size_t smem = ((size_t)4 * B * KTP + 4 * KT) * sizeof(float);
  if (smem > 227 * 1024) return ALIGNQ_ERANGE;
  cudaError_t e = cudaFuncSetAttribute(gram_ffma_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  gram_ffma_bwd_kernel<<<(unsigned)nslabs, GT, smem, s>>>(x, gy, Wsym, gloss, B, Bp, F, ar, eps, tiles_per_slab, gx);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

}  // namespace alignq
