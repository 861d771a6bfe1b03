// Weight quantizer, multi-tensor: per-tensor mean/std reduction, CDF map + rounding + dequant, and
// the backward through mean and std -- 2 launches forward, 2 launches backward for ANY number of
// tensors (the reference runs ~105 ATen kernels per tensor per iteration).
//
// Replaces weight_quantize_fn.forward
//   cdf_alignment/*/model/quantization.py:62-78 (variant A), cdf_alignment_admm/*/model/quantization.py:71-85 (B/C)
// and its autograd backward.
//
// Work decomposition: every tensor (segment of one flat buffer) is cut into ALIGNQ_CHUNK-element
// chunks; one 256-thread CTA per chunk.  Pass 1 writes one fp64 partial pair per chunk; pass 2's
// prologue re-reduces the (few) partials of its own segment, so there are no atomics, no memset
// and the result is bit-reproducible run to run.
#include "common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {

constexpr int kChunk = ALIGNQ_CHUNK;
constexpr int kThreads = 256;
constexpr int kPerThread = kChunk / kThreads;   // 16 elements = 4 float4

struct ChunkRange {
  int seg;
  int64_t begin, end;   // element range in the flat buffer
  int64_t seg_numel;
  int chunk0, nchunks;  // this segment's chunk span
};

__device__ __forceinline__ ChunkRange chunk_range(const int64_t* seg_off, const int32_t* chunk_seg,
                                                  const int32_t* seg_chunk0) {
  ChunkRange r;
  const int c = blockIdx.x;
  r.seg = chunk_seg[c];
  r.chunk0 = seg_chunk0[r.seg];
  r.nchunks = seg_chunk0[r.seg + 1] - r.chunk0;
  const int64_t s0 = seg_off[r.seg], s1 = seg_off[r.seg + 1];
  r.seg_numel = s1 - s0;
  r.begin = s0 + (int64_t)(c - r.chunk0) * kChunk;
  r.end = (r.begin + kChunk < s1) ? r.begin + kChunk : s1;
  return r;
}

// Visit the chunk's elements: f(index, value...) with 128-bit loads when the chunk is 16B aligned.
template <typename F>
__device__ __forceinline__ void for_each_1(const float* __restrict__ a, const ChunkRange& r, F f) {
  const int64_t n = r.end - r.begin;
  const float* pa = a + r.begin;
  if (aligned16(pa) && n == kChunk) {
#pragma unroll
    for (int u = 0; u < kPerThread / 4; ++u) {
      const int i = (u * kThreads + threadIdx.x) * 4;
      float4 v = *reinterpret_cast<const float4*>(pa + i);
      f(r.begin + i, v.x); f(r.begin + i + 1, v.y); f(r.begin + i + 2, v.z); f(r.begin + i + 3, v.w);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += kThreads) f(r.begin + i, pa[i]);
  }
}
template <typename F>
__device__ __forceinline__ void for_each_2(const float* __restrict__ a, const float* __restrict__ b,
                                           const ChunkRange& r, F f) {
  const int64_t n = r.end - r.begin;
  const float* pa = a + r.begin;
  const float* pb = b + r.begin;
  if (aligned16(pa) && aligned16(pb) && n == kChunk) {
#pragma unroll
    for (int u = 0; u < kPerThread / 4; ++u) {
      const int i = (u * kThreads + threadIdx.x) * 4;
      float4 v = *reinterpret_cast<const float4*>(pa + i);
      float4 w = *reinterpret_cast<const float4*>(pb + i);
      f(r.begin + i, v.x, w.x); f(r.begin + i + 1, v.y, w.y); f(r.begin + i + 2, v.z, w.z); f(r.begin + i + 3, v.w, w.w);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += kThreads) f(r.begin + i, pa[i], pb[i]);
  }
}

// Sum this segment's partial pairs; every thread gets the totals.
__device__ __forceinline__ void segment_totals(const double* __restrict__ partials, const ChunkRange& r,
                                               double& s0, double& s1, double* scratch) {
  s0 = 0.0; s1 = 0.0;
  for (int i = threadIdx.x; i < r.nchunks; i += kThreads) {
    s0 += partials[2 * (size_t)(r.chunk0 + i)];
    s1 += partials[2 * (size_t)(r.chunk0 + i) + 1];
  }
  block_sum2(s0, s1, scratch);
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
wq_stats_kernel(const float* __restrict__ flat, const int64_t* __restrict__ seg_off,
                const int32_t* __restrict__ chunk_seg, const int32_t* __restrict__ seg_chunk0,
                double* __restrict__ partials) {
  pdl_trigger();                      // the apply pass launches early and waits for this grid (common.cuh)
  __shared__ double scratch[64];
  const ChunkRange r = chunk_range(seg_off, chunk_seg, seg_chunk0);
  double s = 0.0, ss = 0.0;
  for_each_1(flat, r, [&](int64_t, float v) { const double d = (double)v; s += d; ss = fma(d, d, ss); });
  block_sum2(s, ss, scratch);
  if (threadIdx.x == 0) { partials[2 * (size_t)blockIdx.x] = s; partials[2 * (size_t)blockIdx.x + 1] = ss; }
}

struct SegStats { float mean, std, rstd, var; };

// per-segment upstream-gradient pointers, passed BY VALUE as a kernel parameter (no table upload, and
// therefore nothing but kernel nodes when the step is captured in a CUDA graph)
constexpr int kMaxPtrs = 128;
struct GPtrTable { const float* p[kMaxPtrs]; };

__device__ __forceinline__ SegStats finalize_stats(double s, double ss, int64_t N) {
  SegStats st;
  const double mean = s / (double)N;
  const double var = (ss - s * mean) / (double)(N - 1);     // N == 1 -> 0/0 = NaN, as torch.std
  st.mean = (float)mean;
  st.std = (float)sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
  st.rstd = __frcp_rn(st.std);                              // scale.reciprocal()
  st.var = __fmul_rn(st.std, st.std);                       // scale ** 2
  return st;
}

template <int VARIANT, bool SIGN>
__global__ void __launch_bounds__(kThreads)
wq_fwd_kernel(const float* __restrict__ flat, const int64_t* __restrict__ seg_off,
              const int32_t* __restrict__ chunk_seg, const int32_t* __restrict__ seg_chunk0,
              const double* __restrict__ partials, float n, float inv_n,
              float* __restrict__ wq, float* __restrict__ w_cdf, float* __restrict__ w_pdf,
              int16_t* __restrict__ codes, float* __restrict__ stats) {
  pdl_wait();                         // programmatic dependent of wq_stats_kernel: nothing is touched before this
  __shared__ double scratch[64];
  const ChunkRange r = chunk_range(seg_off, chunk_seg, seg_chunk0);
  double s, ss;
  segment_totals(partials, r, s, ss, scratch);
  const SegStats st = finalize_stats(s, ss, r.seg_numel);
  if (blockIdx.x == r.chunk0 && threadIdx.x == 0) {
    stats[4 * r.seg + 0] = st.mean; stats[4 * r.seg + 1] = st.std;
    stats[4 * r.seg + 2] = st.rstd; stats[4 * r.seg + 3] = (float)r.seg_numel;
  }
  const float two_var = __fmul_rn(2.0f, st.var);
  const float log_scale = logf(st.std);
  for_each_1(flat, r, [&](int64_t i, float w) {
    float c = normal_cdf(w, st.mean, st.rstd);
    if (VARIANT != 0) c = sym_map(c);                                        // QB:53
    float code = SIGN ? ((c > 0.0f) ? 1.0f : ((c < 0.0f) ? -1.0f : c)) : quant_code(c, n);
    float q = SIGN ? code : __fmul_rn(code, inv_n);
    if (VARIANT == 0) q = sym_map(q);                                        // QA:72
    wq[i] = q;
    if (w_cdf) w_cdf[i] = c;
    if (w_pdf) {                                                             // exp(log_prob) * 2  (QA:49)
      const float d = __fsub_rn(w, st.mean);
      float lp = __fdiv_rn(-__fmul_rn(d, d), two_var);
      lp = __fsub_rn(__fsub_rn(lp, log_scale), kLogSqrt2Pi);
      w_pdf[i] = __fmul_rn(expf(lp), 2.0f);
    }
    if (codes) codes[i] = (int16_t)code;
  });
}

// a = 2 g phi(z) (both variants: d wq / d z = 2 phi(z) under the straight-through rounding)
__device__ __forceinline__ void bwd_terms(float w, float g, float mean, float rstd, float& a, float& z) {
  z = __fmul_rn(__fsub_rn(w, mean), rstd);
  const float v = __fmul_rn(z, kInvSqrt2);
  a = __fmul_rn(g, __fmul_rn(2.0f * kInvSqrt2Pi, gauss_kernel_from_v(v)));
}

__global__ void __launch_bounds__(kThreads)
wq_bwd_partial_kernel(const float* __restrict__ flat, const float* g_wq,
                      const int64_t* __restrict__ seg_off, const int32_t* __restrict__ chunk_seg,
                      const int32_t* __restrict__ seg_chunk0, const float* __restrict__ stats,
                      double* __restrict__ partials, const __grid_constant__ GPtrTable gt, int use_gt, int seg_base) {
  pdl_trigger();
  __shared__ double scratch[64];
  const ChunkRange r = chunk_range(seg_off, chunk_seg, seg_chunk0);
  const float mean = stats[4 * r.seg], rstd = stats[4 * r.seg + 2];
  double sa = 0.0, saz = 0.0;
  if (use_gt) {                                   // per-segment upstream gradients (NULL = segment not in this pass)
    if (r.seg < seg_base || r.seg >= seg_base + kMaxPtrs) return;
    const float* gp = gt.p[r.seg - seg_base];
    if (!gp) return;
    g_wq = gp - seg_off[r.seg];
  }
  for_each_2(flat, g_wq, r, [&](int64_t, float w, float g) {
    float a, z;
    bwd_terms(w, g, mean, rstd, a, z);
    sa += (double)a;
    saz = fma((double)a, (double)z, saz);
  });
  block_sum2(sa, saz, scratch);
  if (threadIdx.x == 0) { partials[2 * (size_t)blockIdx.x] = sa; partials[2 * (size_t)blockIdx.x + 1] = saz; }
}

__global__ void __launch_bounds__(kThreads)
wq_bwd_apply_kernel(const float* __restrict__ flat, const float* g_wq,
                    const int64_t* __restrict__ seg_off, const int32_t* __restrict__ chunk_seg,
                    const int32_t* __restrict__ seg_chunk0, const float* __restrict__ stats,
                    const double* __restrict__ partials, float* __restrict__ g_w,
                    const __grid_constant__ GPtrTable gt, int use_gt, int seg_base, int accumulate) {
  pdl_trigger();                      // (the optimizer kernel may follow)
  pdl_wait();                         // programmatic dependent of wq_bwd_partial_kernel
  __shared__ double scratch[64];
  const ChunkRange r = chunk_range(seg_off, chunk_seg, seg_chunk0);
  const float mean = stats[4 * r.seg], rstd = stats[4 * r.seg + 2];
  if (use_gt) {
    if (r.seg < seg_base || r.seg >= seg_base + kMaxPtrs) return;
    const float* gp = gt.p[r.seg - seg_base];
    if (!gp) return;
    g_wq = gp - seg_off[r.seg];
  }
  double sa, saz;
  segment_totals(partials, r, sa, saz, scratch);
  const float mean_a = (float)(sa / (double)r.seg_numel);
  const float kz = (float)(saz / (double)(r.seg_numel - 1));
  for_each_2(flat, g_wq, r, [&](int64_t i, float w, float g) {
    float a, z;
    bwd_terms(w, g, mean, rstd, a, z);
    const float gw = rstd * (a - mean_a - z * kz);
    g_w[i] = accumulate ? g_w[i] + gw : gw;
  });
}

}  // namespace alignq

using namespace alignq;

extern "C" int64_t alignq_wq_plan(const int64_t* seg_off, int nseg, int32_t* chunk_seg, int32_t* seg_chunk0) {
  if (!seg_off || nseg < 0) return ALIGNQ_EINVAL;
  int64_t c = 0;
  for (int t = 0; t < nseg; ++t) {
    const int64_t n = seg_off[t + 1] - seg_off[t];
    if (n < 0) return ALIGNQ_EINVAL;
    const int64_t k = (n + kChunk - 1) / kChunk;
    if (seg_chunk0) seg_chunk0[t] = (int32_t)c;
    if (chunk_seg) for (int64_t i = 0; i < k; ++i) chunk_seg[c + i] = t;
    c += k;
    if (c > 0x7fffffff) return ALIGNQ_ERANGE;
  }
  if (seg_chunk0) seg_chunk0[nseg] = (int32_t)c;
  return c;
}

extern "C" int alignq_wq_forward(const float* flat, const int64_t* seg_off, const int32_t* chunk_seg,
                                 const int32_t* seg_chunk0, int nseg, int64_t nchunks, int w_bit, int variant,
                                 float* wq, float* w_cdf, float* w_pdf, int16_t* codes, float* stats, double* ws,
                                 alignq_stream_t stream) {
  if (nseg < 0 || nchunks < 0 || w_bit < 1 || w_bit >= 32 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (nseg == 0 || nchunks == 0) return ALIGNQ_OK;
  if (!flat || !seg_off || !chunk_seg || !seg_chunk0 || !wq || !stats || !ws) return ALIGNQ_EINVAL;
  if (codes && w_bit > 15) return ALIGNQ_ERANGE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const float n = (float)((1ull << w_bit) - 1);
  const float inv_n = 1.0f / n;
  wq_stats_kernel<<<(unsigned)nchunks, kThreads, 0, s>>>(flat, seg_off, chunk_seg, seg_chunk0, ws);
  ALIGNQ_LAUNCH_CHECK();
#define WQ_FWD(V, S) (void)launch_pdl(wq_fwd_kernel<V, S>, dim3((unsigned)nchunks), dim3(kThreads), 0, s, \
      flat, seg_off, chunk_seg, seg_chunk0, (const double*)ws, n, inv_n, wq, w_cdf, w_pdf, codes, stats)
  if (variant == 0) { if (w_bit == 1) WQ_FWD(0, true); else WQ_FWD(0, false); }
  else              { if (w_bit == 1) WQ_FWD(1, true); else WQ_FWD(1, false); }
#undef WQ_FWD
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_wq_backward(const float* flat, const float* g_wq, const float* const* g_ptrs,
                                  const int64_t* seg_off, const int32_t* chunk_seg, const int32_t* seg_chunk0,
                                  int nseg, int64_t nchunks, int w_bit, const float* stats, float* g_w,
                                  int accumulate, double* ws, alignq_stream_t stream) {
  if (nseg < 0 || nchunks < 0 || w_bit < 1 || w_bit >= 32) return ALIGNQ_EINVAL;
  if (nseg == 0 || nchunks == 0) return ALIGNQ_OK;
  if (!flat || (!g_wq && !g_ptrs) || !seg_off || !chunk_seg || !seg_chunk0 || !stats || !g_w || !ws) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  GPtrTable gt;
  for (int base = 0; base < (g_ptrs ? nseg : 1); base += kMaxPtrs) {
    for (int i = 0; i < kMaxPtrs; ++i) gt.p[i] = (g_ptrs && base + i < nseg) ? g_ptrs[base + i] : nullptr;
    wq_bwd_partial_kernel<<<(unsigned)nchunks, kThreads, 0, s>>>(flat, g_wq, seg_off, chunk_seg, seg_chunk0, stats, ws, gt,
                                                                 g_ptrs ? 1 : 0, base);
    ALIGNQ_LAUNCH_CHECK();
    (void)launch_pdl(wq_bwd_apply_kernel, dim3((unsigned)nchunks), dim3(kThreads), 0, s, flat, g_wq, seg_off, chunk_seg, seg_chunk0,
                     stats, (const double*)ws, g_w, gt, g_ptrs ? 1 : 0, base, accumulate);
    ALIGNQ_LAUNCH_CHECK();
  }
  return ALIGNQ_OK;
}
