// Batch statistics of the FOLLOWING BatchNorm from a convolution's epilogue (the conv output is in registers there):
// per-channel sum and sum of squares over the valid output positions, added to the BatchNorm layer's fp64 accumulator
// copies with the layout of bn_act.cu (ws[(slot * C + c) * 2 + {0: sum, 1: sum of squares}]); the LAST CTA (ticket)
// finishes mean / invstd / running statistics exactly as bnq_stats_kernel's last block does and re-arms everything.
// The statistics launch of the fused bn-act forward (~8 us per layer, latency-bound) disappears.
#pragma once
#include "common.cuh"

namespace alignq {

struct BnStat {
  double* ws;                  // nullptr: no statistics
  unsigned* counter;
  float* running_mean;
  float* running_var;
  float* save_mean;
  float* save_invstd;
  long long* num_batches_tracked;
  float momentum, eps;
  double count;                // N * H * W
};

// Sum 32 per-lane values over the warp so that lane l ends up with the total of value l: a butterfly that halves the
// live values at every step (16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5 for one xor-reduction per value).
__device__ __forceinline__ float warp_reduce_transpose32(float (&v)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float keep = upper ? v[i + step] : v[i];
      const float send = upper ? v[i] : v[i + step];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

// CTA totals of two per-thread channel accumulators a[C], b[C] (C a multiple of 16): the warp sums by the butterfly
// above, the warps through shared memory in fp64 (`red`: NTHREADS / 32 * 2 * C doubles), one fp64 atomic per value into
// ws[(slot * C + c) * 2 + {0: a, 1: b}] (the layout of bn_act.cu); the LAST CTA of the grid (ticket) sums the accumulator
// copies in a fixed order, calls fin(c, sum a, sum b) for its channels and re-arms accumulators and ticket.
// Must be called by every thread of the CTA.
template <int C, int NTHREADS, typename Fin>
__device__ __forceinline__ void bn_cta_finish_from_red(double* red, double* ws, unsigned* counter, Fin fin) {
  // red[warp][2 C] holds the warps' totals (written by every warp, not yet synchronised)
  __shared__ unsigned last_flag;
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < NTHREADS / 32; ++wv) v += red[wv * 2 * C + threadIdx.x];
    atomicAdd(ws + (size_t)(blockIdx.x % ALIGNQ_BN_SLOTS) * C * 2 + threadIdx.x, v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    last_flag = (t == gridDim.x * gridDim.y * gridDim.z - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (last_flag) {
    __threadfence();
    for (int c = threadIdx.x; c < C; c += NTHREADS) {
      double S = 0.0, SS = 0.0;
#pragma unroll
      for (int sl = 0; sl < ALIGNQ_BN_SLOTS; ++sl) {               // fixed order over the accumulator copies
        double* acc = ws + ((size_t)sl * C + c) * 2;
        S += __ldcg(acc); SS += __ldcg(acc + 1);
        acc[0] = 0.0; acc[1] = 0.0;                                // re-arm the accumulators
      }
      fin(c, S, SS);
    }
    if (threadIdx.x == 0) *counter = 0u;                           // re-arm for the next launch
  }
}

template <int C, int NTHREADS, typename Fin>
__device__ __forceinline__ void bn_cta_sums_finish(const float (&a)[C], const float (&b)[C], double* red, double* ws,
                                                   unsigned* counter, Fin fin) {
  static_assert(C % 16 == 0, "channel groups of 16");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += 16) {
    float v[32];
#pragma unroll
    for (int e = 0; e < 16; ++e) { v[e] = a[c0 + e]; v[16 + e] = b[c0 + e]; }
    const float t = warp_reduce_transpose32(v, lane);            // lane l: channel c0 + (l & 15), value l >> 4
    red[warp * 2 * C + 2 * (c0 + (lane & 15)) + (lane >> 4)] = (double)t;
  }
  bn_cta_finish_from_red<C, NTHREADS>(red, ws, counter, fin);
}

// (sum, sum of squares) of channel c -> mean / invstd / running statistics, exactly as bnq_stats_kernel's last block does
__device__ __forceinline__ void bn_stat_finalize(const BnStat& bs, int c, double S, double SS) {
  const double mean = S / bs.count;
  double var = SS / bs.count - mean * mean;                        // biased: what BN normalises with
  var = var < 0.0 ? 0.0 : var;
  bs.save_mean[c] = (float)mean;
  bs.save_invstd[c] = (float)(1.0 / sqrt(var + (double)bs.eps));
  if (bs.running_mean) {
    const double unbiased = bs.count > 1.0 ? var * bs.count / (bs.count - 1.0) : var;
    bs.running_mean[c] = (float)((1.0 - bs.momentum) * bs.running_mean[c] + bs.momentum * mean);
    bs.running_var[c] = (float)((1.0 - bs.momentum) * bs.running_var[c] + bs.momentum * unbiased);
  }
  if (c == 0 && bs.num_batches_tracked) *bs.num_batches_tracked += 1;
}
template <int C, int NTHREADS>
__device__ __forceinline__ void bn_stat_cta_finish(const float (&st_s)[C], const float (&st_ss)[C], double* red, const BnStat& bs) {
  bn_cta_sums_finish<C, NTHREADS>(st_s, st_ss, red, bs.ws, bs.counter, [&](int c, double S, double SS) { bn_stat_finalize(bs, c, S, SS); });
}

// ---- the fused bn-act BACKWARD's reduce pass from a data-gradient convolution's epilogue ---------------------------
// The gradient that convolution produces is the upstream gradient gy of the PRECEDING bn-act layer (its output is the
// convolution's input), so the epilogue can form g_z = gy [y > 0] 2 ar phi(z) and the two BatchNorm-backward sums
// (sum g_z, sum g_z xhat) while gy is still in registers: bnq_bwd_reduce_kernel -- a full pass over x, y, gy -- disappears
// and the bn-act backward is its apply pass alone.
struct BnRed {
  const float* x;              // the bn-act layer's input (nullptr: no reduce)
  const float* y;              // its output (the convolution's forward input): the ReLU mask; unused when relu == 0
  const float* gy2;            // nullable: second consumer's gradient, added to the convolution's (see bn_act.cu: ld4g)
  const float* mean;
  const float* invstd;
  const float* gamma;          // nullable
  const float* beta;           // nullable
  float gscale;                // 2 ar / sqrt(2 pi)
  int relu;
  double* ws;
  unsigned* counter;
  float* coef;                 // [C][2]: mean(g_z), mean(g_z xhat) -- what bnq_bwd_apply_kernel reads
  float* ggamma;               // nullable
  float* gbeta;                // nullable
  double count;                // N * H * W
};

// g_z = gy * [y > 0 if relu] * gscale * exp(-(z/sqrt2)^2): the straight-through quantizer + ReLU backward (bn_act.cu)
__device__ __forceinline__ float bnq_gz_raw(float z, float gy, float yv, float gscale, int relu) {
  const float v = __fmul_rn(z, kInvSqrt2);
  const float g = __fmul_rn(gy, __fmul_rn(gscale, gauss_kernel_from_v(v)));
  return (relu && !(yv > 0.f)) ? 0.f : g;
}

}  // namespace alignq
