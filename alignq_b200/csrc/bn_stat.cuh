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

// CTA totals of per-thread (sum, sum of squares) channel accumulators: warp shuffles, then the warps through shared
// memory in fp64 (`red`: NTHREADS / 32 * 2 * C doubles), one fp64 atomic per value; the last CTA of the grid finishes.
// Must be called by every thread of the CTA.
template <int C, int NTHREADS>
__device__ __forceinline__ void bn_stat_cta_finish(const float (&st_s)[C], const float (&st_ss)[C], double* red, const BnStat& bs) {
  __shared__ unsigned last_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float a = st_s[c], b = st_ss[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { red[warp * 2 * C + 2 * c] = (double)a; red[warp * 2 * C + 2 * c + 1] = (double)b; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < NTHREADS / 32; ++wv) v += red[wv * 2 * C + threadIdx.x];
    atomicAdd(bs.ws + (size_t)(blockIdx.x % ALIGNQ_BN_SLOTS) * C * 2 + threadIdx.x, v);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(bs.counter, 1u);
    last_flag = (t == gridDim.x * gridDim.y * gridDim.z - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (last_flag) {
    __threadfence();
    for (int c = threadIdx.x; c < C; c += NTHREADS) {
      double S = 0.0, SS = 0.0;
#pragma unroll
      for (int sl = 0; sl < ALIGNQ_BN_SLOTS; ++sl) {               // fixed order over the accumulator copies
        double* a = bs.ws + ((size_t)sl * C + c) * 2;
        S += __ldcg(a); SS += __ldcg(a + 1);
        a[0] = 0.0; a[1] = 0.0;                                    // re-arm the accumulators
      }
      const double mean = S / bs.count;
      double var = SS / bs.count - mean * mean;                    // biased: what BN normalises with
      var = var < 0.0 ? 0.0 : var;
      bs.save_mean[c] = (float)mean;
      bs.save_invstd[c] = (float)(1.0 / sqrt(var + (double)bs.eps));
      if (bs.running_mean) {
        const double unbiased = bs.count > 1.0 ? var * bs.count / (bs.count - 1.0) : var;
        bs.running_mean[c] = (float)((1.0 - bs.momentum) * bs.running_mean[c] + bs.momentum * mean);
        bs.running_var[c] = (float)((1.0 - bs.momentum) * bs.running_var[c] + bs.momentum * unbiased);
      }
    }
    if (threadIdx.x == 0) {
      *bs.counter = 0u;                                            // re-arm for the next launch
      if (bs.num_batches_tracked) *bs.num_batches_tracked += 1;
    }
  }
}

}  // namespace alignq
