// Multi-tensor SGD with AlignQ's quantization-aware gradient surrogate: one launch for every
// parameter of the model (the reference loops over parameters in Python, ~8 ATen kernels each).
//
// Replaces SGD.step(idx, w_cdf, w_pdf, lam, lam2) (utils/optimizer.py:196-262) and its helpers
// sigmoid / sigmoid_d / transform (utils/optimizer.py:6-13).
#include "common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {

constexpr int kSgdThreads = 256;

struct SgdScalars { float lam, two_lam2, levels, grad_scale; };

__device__ __forceinline__ float surrogate(float w_cdf, float w_pdf, const SgdScalars& k) {
  // transform(w, lam2) = (((w + 0.5) * (2^bitW - 1)) % 1) * lam2 * 2   (torch.remainder: result in [0, 1))
  float a = __fmul_rn(__fadd_rn(w_cdf, 0.5f), k.levels);
  float r = fmodf(a, 1.0f);
  if (r != 0.0f && r < 0.0f) r = __fadd_rn(r, 1.0f);
  const float t = __fmul_rn(r, k.two_lam2);
  const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-t)));           // sigmoid
  return __fmul_rn(__fmul_rn(__fmul_rn(sg, __fsub_rn(1.0f, sg)), k.lam), w_pdf);
}

__device__ __forceinline__ void sgd_one(float& p, float& g, float& buf, const alignq_sgd_tensor_t& t, bool has_buf,
                                        bool has_sur, float wc, float wp, const SgdScalars& k) {
  float d = (k.grad_scale == 1.0f) ? g : __fmul_rn(g, k.grad_scale);      // 1 / world_size after a summing all-reduce
  if (t.weight_decay != 0.0f) d = fmaf(t.weight_decay, p, d);            // d_p.add_(weight_decay, p)
  if (has_buf) {
    buf = t.first_step ? d : fmaf(1.0f - t.dampening, d, __fmul_rn(t.momentum, buf));
    d = t.nesterov ? fmaf(t.momentum, buf, d) : buf;
  }
  p = fmaf(-t.lr, d, p);                                                 // the update uses d_p (OPT:249-251)
  g = has_sur ? __fmul_rn(d, surrogate(wc, wp, k)) : d;                  // what the reference leaves in p.grad
}

__global__ void __launch_bounds__(kSgdThreads)
sgd_kernel(const alignq_sgd_tensor_t* __restrict__ tensors, const int32_t* __restrict__ chunk_tensor,
           const int32_t* __restrict__ tensor_chunk0, SgdScalars k) {
  pdl_wait();                         // programmatic dependent of whatever produced the gradients: nothing is touched before this
  const int ti = chunk_tensor[blockIdx.x];
  const alignq_sgd_tensor_t t = tensors[ti];
  const int64_t begin = (int64_t)(blockIdx.x - tensor_chunk0[ti]) * ALIGNQ_CHUNK;
  const int64_t n = (t.numel - begin < ALIGNQ_CHUNK) ? (t.numel - begin) : ALIGNQ_CHUNK;
  const bool has_buf = t.buf != nullptr && t.momentum != 0.0f;
  const bool has_sur = t.w_cdf != nullptr;
  float* p = t.p + begin;
  float* g = t.g + begin;
  float* b = has_buf ? t.buf + begin : nullptr;
  const float* wc = has_sur ? t.w_cdf + begin : nullptr;
  const float* wp = has_sur ? t.w_pdf + begin : nullptr;
  const bool vec = n == ALIGNQ_CHUNK && aligned16(p) && aligned16(g) && (!has_buf || aligned16(b)) &&
                   (!has_sur || (aligned16(wc) && aligned16(wp)));
  if (vec) {
#pragma unroll
    for (int u = 0; u < ALIGNQ_CHUNK / (4 * kSgdThreads); ++u) {
      const int i = (u * kSgdThreads + threadIdx.x) * 4;
      float4 P = *reinterpret_cast<float4*>(p + i), G = *reinterpret_cast<float4*>(g + i);
      float4 Bf = has_buf ? *reinterpret_cast<float4*>(b + i) : make_float4(0, 0, 0, 0);
      float4 C = has_sur ? *reinterpret_cast<const float4*>(wc + i) : make_float4(0, 0, 0, 0);
      float4 D = has_sur ? *reinterpret_cast<const float4*>(wp + i) : make_float4(0, 0, 0, 0);
      sgd_one(P.x, G.x, Bf.x, t, has_buf, has_sur, C.x, D.x, k);
      sgd_one(P.y, G.y, Bf.y, t, has_buf, has_sur, C.y, D.y, k);
      sgd_one(P.z, G.z, Bf.z, t, has_buf, has_sur, C.z, D.z, k);
      sgd_one(P.w, G.w, Bf.w, t, has_buf, has_sur, C.w, D.w, k);
      *reinterpret_cast<float4*>(p + i) = P;
      *reinterpret_cast<float4*>(g + i) = G;
      if (has_buf) *reinterpret_cast<float4*>(b + i) = Bf;
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += kSgdThreads) {
      float P = p[i], G = g[i], Bf = has_buf ? b[i] : 0.f;
      sgd_one(P, G, Bf, t, has_buf, has_sur, has_sur ? wc[i] : 0.f, has_sur ? wp[i] : 0.f, k);
      p[i] = P; g[i] = G;
      if (has_buf) b[i] = Bf;
    }
  }
}

}  // namespace alignq

using namespace alignq;

extern "C" int alignq_sgd_step(const alignq_sgd_tensor_t* tensors, const int32_t* chunk_tensor,
                               const int32_t* tensor_chunk0, int ntensors, int64_t nchunks, float lam, float lam2,
                               int bitW, float grad_scale, alignq_stream_t stream) {
  if (ntensors < 0 || nchunks < 0 || bitW < 1 || bitW > 32) return ALIGNQ_EINVAL;
  if (ntensors == 0 || nchunks == 0) return ALIGNQ_OK;
  if (!tensors || !chunk_tensor || !tensor_chunk0) return ALIGNQ_EINVAL;
  SgdScalars k;
  k.lam = lam;
  k.two_lam2 = lam2 * 2.0f;
  k.levels = (float)((1ull << bitW) - 1);
  k.grad_scale = grad_scale;
  (void)launch_pdl(sgd_kernel, dim3((unsigned)nchunks), dim3(kSgdThreads), 0, reinterpret_cast<cudaStream_t>(stream), tensors,
                   chunk_tensor, tensor_chunk0, k);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
