// Classifier head + loss of the CIFAR-style networks in two launches instead of sixteen:
//   out = F.avg_pool2d(out, out.size()[3]); out = out.view(out.size(0), -1); out = self.linear(out)     model/resnet.py:108-110
//   loss = criterion(output, targets)            (nn.CrossEntropyLoss, mean)                              main.py:283-286
// of cdf_alignment/resnet-20-cifar-10.  In the QAT step (B = 128) that tail is 16 tiny library kernels (mean, sgemm + bias,
// log-softmax, nll, their backwards, the bias-gradient reduction, the broadcast of the pooled gradient), ~55 us of pure launch
// latency between the forward and the backward pass (CUPTI timeline, profiles/r02_timeline_resnet20.txt).
//
// forward  (one CTA per sample): pooled = mean_hw feat[i] -> logits = W pooled + b -> log-sum-exp -> per-sample loss;
//          saves pooled [B, C] and g_logits = (softmax - onehot) / B; the last CTA (ticket) sums the per-sample losses in a
//          fixed order.
// backward (B + K CTAs): CTA i < B: g_feat[i, hw, :] = up / HW * (g_logits[i] W) for every hw; CTA B + k: row k of
//          g_W = up * g_logits^T pooled and g_b[k] = up * sum_i g_logits[i, k]   (up = the upstream gradient of the loss, read
//          on the device).  Every sum runs in a fixed order: results are deterministic.
// feat is NHWC ([B, HW, C], channels_last), fp32; targets int64 in [0, K).
#include "common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
namespace head {

constexpr int NT = 256, MAXC = 1024, MAXK = 1024, GLK_MAX = 4096;

__global__ void __launch_bounds__(NT)
head_ce_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ Wt, const float* __restrict__ bias,
                   const long long* __restrict__ target, int B, int HW, int C, int K, float* __restrict__ logits,
                   float* __restrict__ loss, float* __restrict__ pooled_out, float* __restrict__ glogits,
                   float* __restrict__ sample_loss, unsigned* __restrict__ counter) {
  extern __shared__ __align__(16) float sm[];          // [RL][C] partial sums | logits [K] | pooled [C]
  __shared__ float red[2];
  __shared__ unsigned last_flag;
  const int i = blockIdx.x, C4 = C >> 2, RL = NT / C4;
  float* part = sm;
  float* lg = sm + (size_t)RL * C;
  const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (rl < RL) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* f = reinterpret_cast<const float4*>(feat + (size_t)i * HW * C) + c4;
    for (int r = rl; r < HW; r += RL) {
      const float4 v = __ldg(f + (size_t)r * C4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    reinterpret_cast<float4*>(part)[rl * C4 + c4] = a;
  }
  __syncthreads();
  float* pooled = lg + K;
  for (int c = threadIdx.x; c < C; c += NT) {
    float s = 0.f;
    for (int r = 0; r < RL; ++r) s += part[r * C + c];
    const float pv = s / (float)HW;
    pooled[c] = pv;
    pooled_out[(size_t)i * C + c] = pv;
  }
  __syncthreads();
  for (int k = warp; k < K; k += NT / 32) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(__ldg(Wt + (size_t)k * C + c), pooled[c], s);
    s = warp_sum(s);
    if (lane == 0) {
      s += bias ? bias[k] : 0.f;
      lg[k] = s;
      logits[(size_t)i * K + k] = s;
    }
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, lg[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int k = lane; k < K; k += 32) se += expf(lg[k] - mx);
    se = warp_sum(se);
    if (lane == 0) { red[0] = mx; red[1] = logf(se); }
  }
  __syncthreads();
  const float mx = red[0], lse = red[1];
  const int t = (int)target[i];
  const float invB = 1.0f / (float)B;
  for (int k = threadIdx.x; k < K; k += NT) {
    const float lsm = lg[k] - mx - lse;                  // log_softmax
    glogits[(size_t)i * K + k] = (expf(lsm) - (k == t ? 1.f : 0.f)) * invB;
    if (k == t) sample_loss[i] = -lsm;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_flag = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last_flag && warp == 0) {
    __threadfence();
    float s = 0.f;
    for (int j = lane; j < B; j += 32) s += __ldcg(sample_loss + j);
    s = warp_sum(s);
    if (lane == 0) { *loss = s * invB; *counter = 0u; }
  }
}

__global__ void __launch_bounds__(NT)
head_ce_bwd_kernel(const float* __restrict__ Wt, const float* __restrict__ pooled, const float* __restrict__ glogits,
                   const float* __restrict__ gloss, int B, int HW, int C, int K, float* __restrict__ g_feat,
                   float* __restrict__ g_W, float* __restrict__ g_b) {
  extern __shared__ __align__(16) float sm[];           // sample CTAs: g_pooled [C]
  const float up = gloss ? __ldg(gloss) : 1.0f;
  if ((int)blockIdx.x < B) {
    const int i = blockIdx.x;
    if (!g_feat) return;
    const float sc = up / (float)HW;
    for (int c = threadIdx.x; c < C; c += NT) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = fmaf(__ldg(glogits + (size_t)i * K + k), __ldg(Wt + (size_t)k * C + c), s);
      sm[c] = s * sc;
    }
    __syncthreads();
    const int C4 = C >> 2;
    float4* o = reinterpret_cast<float4*>(g_feat + (size_t)i * HW * C);
    for (int idx = threadIdx.x; idx < HW * C4; idx += NT) o[idx] = reinterpret_cast<const float4*>(sm)[idx % C4];
    return;
  }
  // class CTA k: thread = (channel c, sample group sg); the column g_logits[:, k] is staged in shared memory, every group
  // walks its samples with independent loads (a single serial chain over B samples cost ~20 us of L2 latency), and the
  // groups are added in a fixed order
  const int k = blockIdx.x - B;
  float* glk = sm + (C > NT ? C : NT);                   // [min(B, GLK_MAX)]
  const int nst = B < GLK_MAX ? B : GLK_MAX;
  for (int i = threadIdx.x; i < nst; i += NT) glk[i] = __ldg(glogits + (size_t)i * K + k);
  __syncthreads();
  if (g_W) {
    const int SG = C >= NT ? 1 : NT / C;
    for (int c0 = 0; c0 < C; c0 += NT) {
      const int c = c0 + (SG == 1 ? threadIdx.x : threadIdx.x % C), sg = SG == 1 ? 0 : threadIdx.x / C;
      float s = 0.f;
      if (c < C && sg < SG) {
#pragma unroll 8
        for (int i = sg; i < B; i += SG) {
          const float gl = i < GLK_MAX ? glk[i] : __ldg(glogits + (size_t)i * K + k);
          s = fmaf(gl, __ldg(pooled + (size_t)i * C + c), s);
        }
      }
      if (SG == 1) {
        if (c < C) g_W[(size_t)k * C + c] = s * up;
      } else {
        __syncthreads();
        if (sg < SG) sm[sg * C + c] = s;                 // SG * C <= NT floats
        __syncthreads();
        if (threadIdx.x < C) {
          float tot = 0.f;
          for (int g = 0; g < SG; ++g) tot += sm[g * C + threadIdx.x];
          g_W[(size_t)k * C + threadIdx.x] = tot * up;
        }
      }
    }
  }
  if (g_b && threadIdx.x < 32) {
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += 32) s += i < GLK_MAX ? glk[i] : __ldg(glogits + (size_t)i * K + k);
    s = warp_sum(s);
    if (threadIdx.x == 0) g_b[k] = s * up;
  }
}

static int head_args_ok(int B, int HW, int C, int K) {
  if (B < 1 || HW < 1 || C < 4 || K < 1) return ALIGNQ_EINVAL;
  if ((C & 3) || C > MAXC || K > MAXK || B > (1 << 20)) return ALIGNQ_ERANGE;
  return ALIGNQ_OK;
}

}  // namespace head
}  // namespace alignq

using namespace alignq;

extern "C" size_t alignq_head_ce_ws_bytes(int B) { return (size_t)B * sizeof(float) + 16; }

extern "C" int alignq_head_ce_fwd(const float* feat, const float* W, const float* bias, const int64_t* target, int B, int HW,
                                  int C, int K, float* logits, float* loss, float* pooled, float* glogits, void* ws,
                                  size_t ws_bytes, alignq_stream_t stream) {
  int rc = head::head_args_ok(B, HW, C, K);
  if (rc) return rc;
  if (!feat || !W || !target || !logits || !loss || !pooled || !glogits || !ws) return ALIGNQ_EINVAL;
  if ((uintptr_t)feat % 16) return ALIGNQ_EALIGN;
  if (ws_bytes < alignq_head_ce_ws_bytes(B)) return ALIGNQ_ENOSPACE;
  unsigned* counter = reinterpret_cast<unsigned*>(ws);            // ZERO before first use (re-armed by the kernel)
  float* sample_loss = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 16);
  const int RL = head::NT / (C / 4);
  const size_t smem = ((size_t)RL * C + K + C) * sizeof(float);
  head::head_ce_fwd_kernel<<<B, head::NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      feat, W, bias, reinterpret_cast<const long long*>(target), B, HW, C, K, logits, loss, pooled, glogits, sample_loss, counter);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_head_ce_bwd(const float* W, const float* pooled, const float* glogits, const float* gloss, int B, int HW,
                                  int C, int K, float* g_feat, float* g_W, float* g_b, alignq_stream_t stream) {
  int rc = head::head_args_ok(B, HW, C, K);
  if (rc) return rc;
  if (!W || !pooled || !glogits) return ALIGNQ_EINVAL;
  if (g_feat && (uintptr_t)g_feat % 16) return ALIGNQ_EALIGN;
  const size_t smem = ((size_t)(C > head::NT ? C : head::NT) + (B < head::GLK_MAX ? B : head::GLK_MAX)) * sizeof(float);
  head::head_ce_bwd_kernel<<<B + K, head::NT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      W, pooled, glogits, gloss, B, HW, C, K, g_feat, g_W, g_b);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
