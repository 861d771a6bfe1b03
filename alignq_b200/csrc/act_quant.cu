// Activation quantizer: fused Gaussian-CDF map + k-bit rounding + dequant (forward) and the
// single-pass straight-through backward.
//
// Replaces, per element, the ~35 ATen elementwise kernels of
//   activation_quantize_fn.forward   cdf_alignment/*/model/quantization.py:91-103        (variant A)
//   activation_quantize_fn[2].forward cdf_alignment_admm/*/model/quantization.py:102-132 (variants B/C)
// and their autograd backward (uniform_quantize.backward, quantization.py:29-32, chained with
// erf's derivative).  Both kernels are pure HBM streams: 8 B/elem forward, 12 B/elem backward.
#include "common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {

enum ActMode { kQuant = 0, kSign = 1, kCdf = 2 };

struct ActParams {
  float n;          // 2^k - 1
  float inv_n;      // fp32 1/n  (ATen: tensor / python_scalar -> multiply by reciprocal)
  float ar;         // act_range
};

template <int VARIANT, int MODE>
__device__ __forceinline__ float act_fwd_one(float x, const ActParams& p, int& code) {
  float c = normal_cdf_std(x);
  if (VARIANT != 0) c = __fmul_rn(sym_map(c), p.ar);            // QB:53-56
  float q;
  if (MODE == kCdf) { code = 0; return c; }                     // a_bit==32, stage=='align'
  if (MODE == kSign) {
    q = (c > 0.0f) ? 1.0f : ((c < 0.0f) ? -1.0f : c);           // torch.sign (NaN stays NaN)
    code = (int)q;
  } else {
    float r = quant_code(c, p.n);                               // round(c*n)
    code = (int)r;
    q = __fmul_rn(r, p.inv_n);                                  // / n
  }
  if (VARIANT == 0) q = __fmul_rn(sym_map(q), p.ar);            // QA:98  (q*2-1)*act_range
  return q;
}

template <int VARIANT, int MODE, bool CODES, int UNROLL>
__global__ void __launch_bounds__(256, 4)
act_fwd_vec_kernel(const float4* __restrict__ x, float4* __restrict__ y, short4* __restrict__ codes,
                   int64_t n4, ActParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ld_stream(x + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int c0, c1, c2, c3;
      float4 o;
      o.x = act_fwd_one<VARIANT, MODE>(v[u].x, p, c0);
      o.y = act_fwd_one<VARIANT, MODE>(v[u].y, p, c1);
      o.z = act_fwd_one<VARIANT, MODE>(v[u].z, p, c2);
      o.w = act_fwd_one<VARIANT, MODE>(v[u].w, p, c3);
      y[i + u * stride] = o;
      if (CODES) codes[i + u * stride] = make_short4((short)c0, (short)c1, (short)c2, (short)c3);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = ld_stream(x + i), o;
    int c0, c1, c2, c3;
    o.x = act_fwd_one<VARIANT, MODE>(v.x, p, c0);
    o.y = act_fwd_one<VARIANT, MODE>(v.y, p, c1);
    o.z = act_fwd_one<VARIANT, MODE>(v.z, p, c2);
    o.w = act_fwd_one<VARIANT, MODE>(v.w, p, c3);
    y[i] = o;
    if (CODES) codes[i] = make_short4((short)c0, (short)c1, (short)c2, (short)c3);
  }
}

// scalar path: ragged tails and unaligned views
template <int VARIANT, int MODE, bool CODES>
__global__ void __launch_bounds__(256)
act_fwd_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, int16_t* __restrict__ codes,
                      int64_t begin, int64_t numel, ActParams p) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    int c;
    y[i] = act_fwd_one<VARIANT, MODE>(x[i], p, c);
    if (CODES) codes[i] = (int16_t)c;
  }
}

// backward: gx = gy * gscale * exp(-v^2), v = x/sqrt(2); gscale = 2*ar/sqrt(2*pi) (or 1/sqrt(2*pi)
// when variant A returns the bare CDF).  Argument of the exponential is formed exactly as erf's
// ATen backward forms it (exp(-(v*v)) with v the forward's fp32 v), so the result stays within a
// few ulp of the reference chain for every |x|.
__device__ __forceinline__ float act_bwd_one(float x, float g, float gscale) {
  float v = __fmul_rn(x, kInvSqrt2);
  return __fmul_rn(g, __fmul_rn(gscale, gauss_kernel_from_v(v)));
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
act_bwd_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, float4* __restrict__ gx,
                   int64_t n4, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 a[UNROLL], b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) { a[u] = ld_stream(x + i + u * stride); b[u] = ld_stream(gy + i + u * stride); }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float4 o;
      o.x = act_bwd_one(a[u].x, b[u].x, gscale);
      o.y = act_bwd_one(a[u].y, b[u].y, gscale);
      o.z = act_bwd_one(a[u].z, b[u].z, gscale);
      o.w = act_bwd_one(a[u].w, b[u].w, gscale);
      gx[i + u * stride] = o;
    }
  }
  for (; i < n4; i += stride) {
    float4 a = ld_stream(x + i), b = ld_stream(gy + i), o;
    o.x = act_bwd_one(a.x, b.x, gscale);
    o.y = act_bwd_one(a.y, b.y, gscale);
    o.z = act_bwd_one(a.z, b.z, gscale);
    o.w = act_bwd_one(a.w, b.w, gscale);
    gx[i] = o;
  }
}

__global__ void __launch_bounds__(256)
act_bwd_scalar_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                      int64_t begin, int64_t numel, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    gx[i] = act_bwd_one(x[i], gy[i], gscale);
}

// gx = gadd + gy * gscale * exp(-v^2): the straight-through term added to a gradient that arrived by another route
// (data-parallel feature-sharded ADMM term: the all-to-all'ed Gram gradient, utils/dp_gram.py)
__global__ void __launch_bounds__(256)
act_bwd_add_kernel(const float* __restrict__ x, const float* __restrict__ gy, const float* __restrict__ gadd,
                   float* __restrict__ gx, int64_t numel, int vec, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = vec ? numel / 4 : 0;
  for (int64_t k = i; k < n4; k += stride) {
    const float4 a = ld_stream(reinterpret_cast<const float4*>(x) + k), b = ld_stream(reinterpret_cast<const float4*>(gy) + k);
    const float4 c = ld_stream(reinterpret_cast<const float4*>(gadd) + k);
    float4 o;
    o.x = c.x + act_bwd_one(a.x, b.x, gscale);
    o.y = c.y + act_bwd_one(a.y, b.y, gscale);
    o.z = c.z + act_bwd_one(a.z, b.z, gscale);
    o.w = c.w + act_bwd_one(a.w, b.w, gscale);
    reinterpret_cast<float4*>(gx)[k] = o;
  }
  for (int64_t k = n4 * 4 + i; k < numel; k += stride) gx[k] = gadd[k] + act_bwd_one(x[k], gy[k], gscale);
}

// grid sizing: whole multiples of the SM count, 8 resident 256-thread CTAs per SM at most
static inline int stream_grid(int64_t work_items, int per_block) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)ALIGNQ_NUM_SMS * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <int VARIANT, int MODE, bool CODES>
static int launch_act_fwd(const float* x, float* y, int16_t* codes, int64_t numel, ActParams p, cudaStream_t s) {
  constexpr int UNROLL = 4;
  const bool vec = aligned16(x) && aligned16(y) && (!CODES || (reinterpret_cast<uintptr_t>(codes) & 7u) == 0);
  int64_t n4 = vec ? numel / 4 : 0;
  if (n4 > 0) {
    int grid = stream_grid(n4, 256 * UNROLL);
    act_fwd_vec_kernel<VARIANT, MODE, CODES, UNROLL><<<grid, 256, 0, s>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), reinterpret_cast<short4*>(codes), n4, p);
    ALIGNQ_LAUNCH_CHECK();
  }
  if (n4 * 4 < numel) {
    int grid = stream_grid(numel - n4 * 4, 256);
    act_fwd_scalar_kernel<VARIANT, MODE, CODES><<<grid, 256, 0, s>>>(x, y, codes, n4 * 4, numel, p);
    ALIGNQ_LAUNCH_CHECK();
  }
  return ALIGNQ_OK;
}

template <int VARIANT, int MODE>
static int dispatch_codes(const float* x, float* y, int16_t* codes, int64_t numel, ActParams p, cudaStream_t s) {
  return codes ? launch_act_fwd<VARIANT, MODE, true>(x, y, codes, numel, p, s)
               : launch_act_fwd<VARIANT, MODE, false>(x, y, nullptr, numel, p, s);
}

}  // namespace alignq

using namespace alignq;

extern "C" int alignq_act_fwd(const float* x, float* y, int16_t* codes, int64_t numel, int a_bit, float act_range,
                              int variant, int return_cdf, alignq_stream_t stream) {
  if (numel < 0 || a_bit < 1 || a_bit > 32 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (a_bit == 32 && !return_cdf) return ALIGNQ_EINVAL;            // identity is the caller's job (QA:92-95)
  if (numel == 0) return ALIGNQ_OK;
  if (!x || !y) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ActParams p;
  p.n = (a_bit == 32) ? 1.0f : (float)((1ull << a_bit) - 1);
  p.inv_n = 1.0f / p.n;
  p.ar = act_range;
  const int mode = (a_bit == 32) ? kCdf : (a_bit == 1 ? kSign : kQuant);
  if (codes && mode == kQuant) {
    double maxcode = (variant == 0 ? 1.0 : (double)fabsf(act_range)) * (double)p.n;
    if (maxcode > 32767.0) return ALIGNQ_ERANGE;
  }
  const int v = variant == 0 ? 0 : 1;
  if (v == 0) {
    if (mode == kQuant) return dispatch_codes<0, kQuant>(x, y, codes, numel, p, s);
    if (mode == kSign) return dispatch_codes<0, kSign>(x, y, codes, numel, p, s);
    return dispatch_codes<0, kCdf>(x, y, nullptr, numel, p, s);
  }
  if (mode == kQuant) return dispatch_codes<1, kQuant>(x, y, codes, numel, p, s);
  if (mode == kSign) return dispatch_codes<1, kSign>(x, y, codes, numel, p, s);
  return dispatch_codes<1, kCdf>(x, y, nullptr, numel, p, s);
}

extern "C" float alignq_act_grad_scale(int a_bit, float act_range, int variant, int return_cdf) {
  // d y / d Phi-argument: variant A returning the bare CDF has slope phi(x); everything else 2*ar*phi(x)
  const float base = (variant == 0 && a_bit == 32 && return_cdf) ? 1.0f : 2.0f * act_range;
  return base * kInvSqrt2Pi;
}

extern "C" int alignq_act_bwd(const float* x, const float* gy, float* gx, int64_t numel, int a_bit, float act_range,
                              int variant, int return_cdf, alignq_stream_t stream) {
  if (numel < 0 || a_bit < 1 || a_bit > 32 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (a_bit == 32 && !return_cdf) return ALIGNQ_EINVAL;
  if (numel == 0) return ALIGNQ_OK;
  if (!x || !gy || !gx) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float gscale = alignq_act_grad_scale(a_bit, act_range, variant, return_cdf);
  constexpr int UNROLL = 4;
  const bool vec = aligned16(x) && aligned16(gy) && aligned16(gx);
  int64_t n4 = vec ? numel / 4 : 0;
  if (n4 > 0) {
    int grid = stream_grid(n4, 256 * UNROLL);
    act_bwd_vec_kernel<UNROLL><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(x),
                                                    reinterpret_cast<const float4*>(gy),
                                                    reinterpret_cast<float4*>(gx), n4, gscale);
    ALIGNQ_LAUNCH_CHECK();
  }
  if (n4 * 4 < numel) {
    int grid = stream_grid(numel - n4 * 4, 256);
    act_bwd_scalar_kernel<<<grid, 256, 0, s>>>(x, gy, gx, n4 * 4, numel, gscale);
    ALIGNQ_LAUNCH_CHECK();
  }
  return ALIGNQ_OK;
}

extern "C" int alignq_act_bwd_add(const float* x, const float* gy, const float* gadd, float* gx, int64_t numel, int a_bit,
                                  float act_range, int variant, int return_cdf, alignq_stream_t stream) {
  if (numel < 0 || a_bit < 1 || a_bit > 32 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (a_bit == 32 && !return_cdf) return ALIGNQ_EINVAL;
  if (numel == 0) return ALIGNQ_OK;
  if (!x || !gy || !gadd || !gx) return ALIGNQ_EINVAL;
  const float gscale = alignq_act_grad_scale(a_bit, act_range, variant, return_cdf);
  const int vec = aligned16(x) && aligned16(gy) && aligned16(gx) && aligned16(gadd);
  act_bwd_add_kernel<<<stream_grid((numel + 3) / 4, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, gy, gadd, gx, numel, vec, gscale);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone L1 pieces of the reference surface (the hot path uses the fused kernels above).
//   uniform_quantize(k).forward     quantization.py:15-27
//   cdf(m, s, src).forward          quantization.py:45-50 (A) / 49-59 (B, C); m, s are device scalars
namespace alignq {

__global__ void __launch_bounds__(256)
uniform_q_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t numel, int k, float n, float inv_n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const float v = x[i];
    y[i] = (k == 1) ? ((v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : v)) : __fmul_rn(rintf(__fmul_rn(v, n)), inv_n);
  }
}

// mode 0: forward (out0 = mapped cdf, out1 = pdf);  mode 1: backward wrt x (out0 = gx), in1 = g_cdf, in2 = g_pdf
__global__ void __launch_bounds__(256)
cdf_kernel(const float* __restrict__ x, const float* __restrict__ m_ptr, const float* __restrict__ s_ptr, int sym,
           float post_scale, int mode, const float* __restrict__ g_cdf, const float* __restrict__ g_pdf,
           float* __restrict__ out0, float* __restrict__ out1, int64_t numel) {
  const float m = __ldg(m_ptr), s = __ldg(s_ptr);
  const float r = __frcp_rn(s);
  const float var = __fmul_rn(s, s), two_var = __fmul_rn(2.0f, var), log_s = logf(s);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const float xv = x[i];
    const float d = __fsub_rn(xv, m);
    float lp = __fdiv_rn(-__fmul_rn(d, d), two_var);
    lp = __fsub_rn(__fsub_rn(lp, log_s), kLogSqrt2Pi);
    const float pdf = __fmul_rn(expf(lp), 2.0f);
    if (mode == 0) {
      float c = normal_cdf(xv, m, r);
      if (sym) { c = sym_map(c); if (post_scale != 1.0f) c = __fmul_rn(c, post_scale); }
      out0[i] = c;
      if (out1) out1[i] = pdf;
    } else {
      // d cdf/dx = 0.5 pdf (pdf = 2 N(x; m, s)), times 2*post_scale in the symmetric map; d pdf/dx = -pdf d / var
      float g = 0.f;
      if (g_cdf) g = g_cdf[i] * (sym ? pdf * post_scale : 0.5f * pdf);
      if (g_pdf) g += g_pdf[i] * (-pdf * d / var);
      out0[i] = g;
    }
  }
}

}  // namespace alignq

extern "C" int alignq_uniform_q_fwd(const float* x, float* y, int64_t numel, int k, alignq_stream_t stream) {
  if (numel < 0 || k < 1 || k >= 32) return ALIGNQ_EINVAL;
  if (numel == 0) return ALIGNQ_OK;
  if (!x || !y) return ALIGNQ_EINVAL;
  const float n = (float)((1ull << k) - 1);
  uniform_q_kernel<<<stream_grid(numel, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, numel, k, n, 1.0f / n);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_cdf_fwd(const float* x, const float* m, const float* s, int variant, int src_is_act,
                              float act_range, float* cdf_out, float* pdf_out, int64_t numel, alignq_stream_t stream) {
  if (numel < 0 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (numel == 0) return ALIGNQ_OK;
  if (!x || !m || !s || !cdf_out) return ALIGNQ_EINVAL;
  const int sym = variant != 0;
  const float post = (sym && src_is_act) ? act_range : 1.0f;
  cdf_kernel<<<stream_grid(numel, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, m, s, sym, post, 0, nullptr, nullptr, cdf_out, pdf_out, numel);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_cdf_bwd(const float* x, const float* m, const float* s, int variant, int src_is_act,
                              float act_range, const float* g_cdf, const float* g_pdf, float* gx, int64_t numel,
                              alignq_stream_t stream) {
  if (numel < 0 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (numel == 0) return ALIGNQ_OK;
  if (!x || !m || !s || !gx) return ALIGNQ_EINVAL;
  const int sym = variant != 0;
  const float post = (sym && src_is_act) ? act_range : 1.0f;
  cdf_kernel<<<stream_grid(numel, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, m, s, sym, post, 1, g_cdf, g_pdf, gx, nullptr, numel);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
