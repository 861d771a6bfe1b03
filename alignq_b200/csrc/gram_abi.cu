// C-ABI entry points of the correlation / fused activation+ADMM path (dispatch over gram_mode).
#include "common.cuh"
#include "gram_common.cuh"
#include "../../include/alignq_b200.h"

using namespace alignq;

extern "C" size_t alignq_gram_ws_bytes(int B, int64_t F) {
  if (B < 1 || F < 1) return 0;
  // room for up to 4 accumulator partials (tf32x3 fused: HH^T, HL^T for x and t) from up to 2 * 148 CTAs
  const int64_t ntiles = (F + 31) / 32;
  int64_t slabs = ntiles < 2 * ALIGNQ_NUM_SMS ? ntiles : 2 * ALIGNQ_NUM_SMS;
  const size_t head = gram_wsym_floats(B) * sizeof(float);
  size_t bytes = head + 4 * (size_t)slabs * B * B * sizeof(float);
  const size_t floor_bytes = head + 4 * (size_t)B * B * sizeof(float);
  if (bytes > kGramWsCapBytes) bytes = kGramWsCapBytes > floor_bytes ? kGramWsCapBytes : floor_bytes;
  return bytes;
}

static inline float* ws_partials(void* ws, int B) { return reinterpret_cast<float*>(ws) + gram_wsym_floats(B); }

extern "C" int alignq_corr_fwd(const float* x, const float* y, int B, int64_t F, float eps, float* G, void* ws,
                               size_t ws_bytes, int gram_mode, alignq_stream_t stream) {
  if (B < 1 || F < 1 || !x || !y || !G || !ws) return ALIGNQ_EINVAL;
  if (B > 1024) return ALIGNQ_ERANGE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (gram_mode == ALIGNQ_GRAM_FP32 || x != y || B > 128 || B < 2) {      // tcgen05 paths: one batch tile, 2 <= B <= 128
    ActQ q{0.f, 0.f, 0.f, 0};
    int nslabs = 0;
    int rc = gram_ffma_forward(x, y, B, F, eps, 0, q, nullptr, ws_partials(ws, B), &nslabs, ws_bytes, s);
    if (rc) return rc;
    return launch_gram_reduce(ws_partials(ws, B), nslabs, B, F, 0, G, nullptr, s);
  }
  return gram_tc_corr(x, B, F, eps, G, ws, ws_bytes, gram_mode, s);
}

extern "C" int alignq_corr_bwd(const float* x, const float* y, const float* dG, int B, int64_t F, float eps, float* gx,
                               float* gy, void* ws, size_t ws_bytes, alignq_stream_t stream) {
  if (B < 2 || F < 1 || !x || !y || !dG || !ws) return ALIGNQ_EINVAL;
  if (B > 1024) return ALIGNQ_ERANGE;
  if (x == y && gy && gy != gx) return ALIGNQ_EINVAL;            // one operand, one gradient
  const size_t w = gram_wsym_floats(B);
  if (ws_bytes < 2 * w * sizeof(float)) return ALIGNQ_ENOSPACE;
  float* wt = reinterpret_cast<float*>(ws);
  return corr_ffma_backward(x, y, dG, B, F, eps, gx, gy, wt, wt + w, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int alignq_act_admm_fwd(const float* x, int B, int64_t F, int a_bit, float act_range, float eps,
                                   const float* Z, const float* U, int dim, float mu, float rho, float* y, float* D,
                                   float* loss, float* dLdD, void* ws, size_t ws_bytes, int gram_mode,
                                   alignq_stream_t stream) {
  if (B < 1 || F < 1 || a_bit < 1 || a_bit > 32 || dim < B) return ALIGNQ_EINVAL;
  if (!x || !Z || !U || !y || !D || !loss || !ws) return ALIGNQ_EINVAL;
  if (B > 1024) return ALIGNQ_ERANGE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ActQ q;
  q.ar = act_range;
  q.a_bit = a_bit;
  q.n = (a_bit == 32) ? 1.0f : (float)((1ull << a_bit) - 1);
  q.inv_n = 1.0f / q.n;
  int rc;
  if (gram_mode == ALIGNQ_GRAM_FP32 || B > 128 || B < 2) {          // larger batches: the fp32 FFMA kernels
    int nslabs = 0;
    rc = gram_ffma_forward(x, x, B, F, eps, 1, q, y, ws_partials(ws, B), &nslabs, ws_bytes, s);
    if (rc) return rc;
    rc = launch_gram_reduce(ws_partials(ws, B), nslabs, B, F, 1, nullptr, D, s);
  } else {
    // tensor-core modes: the ADMM loss and dL/dD ride on the split-K reduction (one launch instead of two)
    const AdmmFinish fin{Z, U, dim, mu, rho, loss, dLdD};
    return gram_tc_fused_fwd(x, B, F, q, eps, y, D, ws, ws_bytes, gram_mode, &fin, s);
  }
  if (rc) return rc;
  return alignq_admm_loss(D, B, Z, U, dim, 1, mu, rho, nullptr, 0, loss, dLdD, nullptr, nullptr, stream);
}

extern "C" int alignq_gram_sums_fwd(const float* x, int B, int64_t F, float act_range, float eps, float* sums, void* ws,
                                    size_t ws_bytes, int gram_mode, alignq_stream_t stream) {
  if (B < 2 || F < 1 || !x || !sums || !ws) return ALIGNQ_EINVAL;
  if (B > 1024) return ALIGNQ_ERANGE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  ActQ q{act_range, 1.0f, 1.0f, 32};                           // only the map t = (2 Phi(x) - 1) ar is needed, no rounding
  float* Gx = sums;
  float* Gt = sums + (size_t)B * B;
  if (gram_mode == ALIGNQ_GRAM_FP32 || B > 128) {
    int nslabs = 0;
    int rc = gram_ffma_forward(x, x, B, F, eps, 1, q, nullptr, ws_partials(ws, B), &nslabs, ws_bytes, s);
    if (rc) return rc;
    return launch_gram_reduce_raw(ws_partials(ws, B), nslabs, B, Gx, Gt, s);
  }
  return gram_tc_sums(x, B, F, q, eps, Gx, Gt, ws, ws_bytes, gram_mode, s);
}

extern "C" int alignq_act_admm_bwd(const float* x, const float* gy, const float* dLdD, const float* gloss, int B,
                                   int64_t F, int a_bit, float act_range, float eps, float* gx, void* ws,
                                   size_t ws_bytes, int gram_mode, alignq_stream_t stream) {
  if (B < 2 || F < 1 || a_bit < 1 || a_bit > 32) return ALIGNQ_EINVAL;
  if (!x || !dLdD || !gx || !ws) return ALIGNQ_EINVAL;
  if (ws_bytes < gram_wsym_floats(B) * sizeof(float)) return ALIGNQ_ENOSPACE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* Wsym = reinterpret_cast<float*>(ws);
  int rc = launch_wsym(dLdD, B, Wsym, s);
  if (rc) return rc;
  if (gram_mode != ALIGNQ_GRAM_FP32 && (B <= 32 || (B <= 128 && aligned16(x) && (F % 4) == 0)))      // tensor-core products
    return gram_tc_backward(x, gy, Wsym, gram_bp(B), gloss, B, F, act_range, eps, gx,
                            gram_mode == ALIGNQ_GRAM_TF32X3 ? 1 : 0, s);
  return gram_ffma_backward(x, gy, Wsym, gram_bp(B), gloss, B, F, act_range, eps, gx, s);
}
