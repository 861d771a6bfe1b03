// Shared declarations of the Gram / ADMM kernels (workspace layout, cross-file launchers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "common.cuh"

namespace alignq {

// Workspace layout (floats), all offsets 256-byte aligned:
//   [0, Bp*Bp)                      Wsym = dLdD + dLdD^T, leading dimension Bp = roundup(B, 4) (backward)
//   [wsym_floats, ...)              split-K partials [2][nslabs][B][B]                            (forward)
constexpr size_t kGramWsCapBytes = 256ull << 20;

__host__ __device__ inline int gram_bp(int B) { return (B + 3) & ~3; }
inline size_t gram_wsym_floats(int B) { size_t n = (size_t)gram_bp(B) * gram_bp(B); return (n + 63) & ~(size_t)63; }
inline int64_t gram_ws_slab_cap(size_t ws_bytes, int B) {
  const size_t head = gram_wsym_floats(B) * sizeof(float);
  if (ws_bytes <= head) return 0;
  return (int64_t)((ws_bytes - head) / (2 * (size_t)B * B * sizeof(float)));
}

struct ActQ {                    // activation map / quantizer parameters for the T operand
  float ar, n, inv_n;
  int a_bit;
};
int gram_ffma_forward(const float* xa, const float* xb, int B, int64_t F, float eps, int fused, ActQ q, float* y,
                      float* partials, int* nslabs_out, size_t ws_bytes, cudaStream_t s);
int gram_ffma_backward(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B,
                       int64_t F, float ar, float eps, float* gx, cudaStream_t s);

int corr_ffma_backward(const float* x, const float* y, const float* dG, int B, int64_t F, float eps, float* gx, float* gy,
                       float* wt0, float* wt1, cudaStream_t s);

#ifdef __CUDACC__
__device__ __forceinline__ float act_map_t(float x, float ar) {           // QB:49-56
  return __fmul_rn(sym_map(normal_cdf_std(x)), ar);
}
__device__ __forceinline__ float act_quant_from_t(float t, const ActQ& q) {   // QB:110 (uniform_q)
  if (q.a_bit == 32) return t;
  if (q.a_bit == 1) return (t > 0.0f) ? 1.0f : ((t < 0.0f) ? -1.0f : t);
  return __fmul_rn(rintf(__fmul_rn(t, q.n)), q.inv_n);
}
#endif

// ADMM loss fused behind the split-K reduction of the tensor-core forward (gram_tc.cu: gram_finish_tc_kernel)
struct AdmmFinish {
  const float* Z;      // alterD [dim, dim]
  const float* U;      // gamma  [dim, dim]
  int dim;
  float mu, rho;
  float* loss;         // scalar
  float* dLdD;         // [B, B] or null
};

// gram_tc.cu (tcgen05 modes)
int gram_tc_corr(const float* x, int B, int64_t F, float eps, float* G, void* ws, size_t ws_bytes, int gram_mode,
                 cudaStream_t s);
int gram_tc_fused_fwd(const float* x, int B, int64_t F, ActQ q, float eps, float* y, float* D, void* ws,
                      size_t ws_bytes, int gram_mode, const AdmmFinish* fin, cudaStream_t s);
// un-normalised column sums  sum_f Xs Xs^T -> Gx,  sum_f Ts Ts^T -> Gt  (data-parallel feature-sharded partials)
int gram_tc_sums(const float* x, int B, int64_t F, ActQ q, float eps, float* Gx, float* Gt, void* ws, size_t ws_bytes,
                 int gram_mode, cudaStream_t s);
// gram_tc_bwd.cu
int gram_tc_backward(const float* x, const float* gy, const float* Wsym, int Bp, const float* gloss, int B, int64_t F,
                     float ar, float eps, float* gx, int split, cudaStream_t s);
// admm.cu
int launch_gram_reduce(const float* partials, int nslabs, int B, int64_t F, int fused, float* Gx_or_G, float* D,
                       cudaStream_t s);
int launch_gram_reduce_raw(const float* partials, int nslabs, int B, float* Gx, float* Gt, cudaStream_t s);
int launch_wsym(const float* dLdD, int B, float* Wsym, cudaStream_t s);

}  // namespace alignq
