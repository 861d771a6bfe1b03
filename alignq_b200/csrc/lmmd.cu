// LMMD (local maximum mean discrepancy) loss of the DSAN head -- SURVEY.md 8(f) item 4:
//   guassian_kernel / lmmd          cdf_alignment_admm/dsan_office/utils/mmd.py:9-41
// total = cat(source, target) [n = 2B, d];  L2[i,j] = sum_k (total_i[k] - total_j[k])^2;
// bandwidth = sum(L2) / (n^2 - n) (a constant: `.data`), divided by kernel_mul^(kernel_num // 2);
// K[i,j] = sum_{m < kernel_num} exp(-L2[i,j] / (bandwidth * kernel_mul^m));
// loss = sum(w_ss * K[:B,:B] + w_tt * K[B:,B:] - 2 * w_st * K[:B,B:]) = sum_ij W[i,j] K[i,j] with the signed full-size
// weight matrix W = [[w_ss, -w_st], [-w_st^T, w_tt]] (K is symmetric); loss = 0 when any K is NaN (mmd.py:33-34).
// Another pairwise (Gram-like) contraction, but tiny (n <= 256, d = 256 ... 2048): three latency-bound launches, fp32
// arithmetic in the reference's op order where it matters (squared differences summed per pair, exp per bandwidth).
// Backward (bandwidth is a constant): d loss / d total_i = 2 sum_j (W_ij + W_ji) K'_ij (total_i - total_j),
// K'_ij = dK/dL2 = -sum_m exp(-L2_ij / bw_m) / bw_m.
#include "common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {

constexpr int LMMD_MAXN = 1024;

// ws doubles: [0] sum(L2), [1] NaN flag (as double), [2] loss accumulator;  L2 -> scratch [n, n]
__global__ void __launch_bounds__(256)
lmmd_l2_kernel(const float* __restrict__ total, int n, int d, float* __restrict__ L2, double* __restrict__ ws) {
  extern __shared__ float row[];                 // total_i
  __shared__ double part[8];
  const int i = blockIdx.x;
  for (int k = threadIdx.x; k < d; k += blockDim.x) row[k] = total[(size_t)i * d + k];
  __syncthreads();
  double local = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float* tj = total + (size_t)j * d;
    float s = 0.f;
    for (int k = 0; k < d; ++k) { const float df = __fsub_rn(row[k], tj[k]); s = __fadd_rn(s, __fmul_rn(df, df)); }
    L2[(size_t)i * n + j] = s;
    local += (double)s;
  }
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(ws, t);
  }
}

__global__ void __launch_bounds__(256)
lmmd_loss_kernel(const float* __restrict__ L2, const float* __restrict__ W, int n, float kernel_mul, int kernel_num,
                 float fix_sigma, float* __restrict__ coef, double* __restrict__ ws) {
  __shared__ double part[8];
  __shared__ int nanp[8];
  const int i = blockIdx.x;
  float bw = fix_sigma > 0.f ? fix_sigma : (float)(ws[0] / ((double)n * n - (double)n));
  for (int m = 0; m < kernel_num / 2; ++m) bw = __fdiv_rn(bw, kernel_mul);        // bandwidth /= kernel_mul ** (kernel_num // 2)
  double local = 0.0;
  int anynan = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float l2 = L2[(size_t)i * n + j];
    float kv = 0.f, kd = 0.f, b = bw;
    for (int m = 0; m < kernel_num; ++m) {                                        // bandwidth * kernel_mul ** m
      const float e = expf(__fdiv_rn(-l2, b));
      kv = __fadd_rn(kv, e);
      kd -= e / b;
      b = __fmul_rn(b, kernel_mul);
    }
    anynan |= (kv != kv);
    const float wij = W[(size_t)i * n + j], wji = W[(size_t)j * n + i];
    local += (double)__fmul_rn(wij, kv);
    coef[(size_t)i * n + j] = (wij + wji) * kd;
  }
  local = warp_sum(local);
  anynan = __any_sync(0xffffffffu, anynan);
  if ((threadIdx.x & 31) == 0) { part[threadIdx.x >> 5] = local; nanp[threadIdx.x >> 5] = anynan; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int f = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t += part[w]; f |= nanp[w]; }
    atomicAdd(ws + 2, t);
    if (f) ws[1] = 1.0;
  }
}

__global__ void lmmd_finish_kernel(double* __restrict__ ws, float* __restrict__ loss) {
  *loss = (ws[1] != 0.0) ? 0.f : (float)ws[2];          // NaN anywhere in the kernel matrix: the reference returns 0
}

// g_total[i, k] = gloss * 2 * sum_j coef[i, j] (total_i[k] - total_j[k]); zero when the forward hit the NaN branch
__global__ void __launch_bounds__(256)
lmmd_bwd_kernel(const float* __restrict__ total, int n, int d, const float* __restrict__ coef, const float* __restrict__ gloss,
                const double* __restrict__ ws, float* __restrict__ g_total) {
  extern __shared__ float crow[];                // coef[i, :]
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < n; j += blockDim.x) crow[j] = coef[(size_t)i * n + j];
  __syncthreads();
  const bool dead = ws[1] != 0.0;
  const float g = 2.0f * (gloss ? __ldg(gloss) : 1.0f);
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float ti = total[(size_t)i * d + k];
    float s = 0.f;
    for (int j = 0; j < n; ++j) s = fmaf(crow[j], ti - total[(size_t)j * d + k], s);
    g_total[(size_t)i * d + k] = dead ? 0.f : g * s;
  }
}

}  // namespace alignq

using namespace alignq;

extern "C" size_t alignq_lmmd_ws_bytes(int n) { return 4 * sizeof(double) + (size_t)n * n * sizeof(float); }

extern "C" int alignq_lmmd_fwd(const float* total, int n, int d, const float* W, float kernel_mul, int kernel_num,
                               float fix_sigma, float* loss, float* coef, void* ws, size_t ws_bytes, alignq_stream_t stream) {
  if (!total || !W || !loss || !coef || !ws || n < 2 || d < 1 || kernel_num < 1 || kernel_num > 16) return ALIGNQ_EINVAL;
  if (n > LMMD_MAXN || d > 12288) return ALIGNQ_ERANGE;
  if (ws_bytes < alignq_lmmd_ws_bytes(n)) return ALIGNQ_ENOSPACE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  double* acc = reinterpret_cast<double*>(ws);                 // must be ZERO on entry (4 doubles)
  float* L2 = reinterpret_cast<float*>(acc + 4);
  lmmd_l2_kernel<<<n, 256, (size_t)d * sizeof(float), s>>>(total, n, d, L2, acc);
  ALIGNQ_LAUNCH_CHECK();
  lmmd_loss_kernel<<<n, 256, 0, s>>>(L2, W, n, kernel_mul, kernel_num, fix_sigma, coef, acc);
  ALIGNQ_LAUNCH_CHECK();
  lmmd_finish_kernel<<<1, 1, 0, s>>>(acc, loss);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_lmmd_bwd(const float* total, int n, int d, const float* coef, const float* gloss, const void* ws,
                               float* g_total, alignq_stream_t stream) {
  if (!total || !coef || !ws || !g_total || n < 2 || d < 1) return ALIGNQ_EINVAL;
  if (n > LMMD_MAXN) return ALIGNQ_ERANGE;
  lmmd_bwd_kernel<<<n, 256, (size_t)n * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>(
      total, n, d, coef, gloss, reinterpret_cast<const double*>(ws), g_total);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
