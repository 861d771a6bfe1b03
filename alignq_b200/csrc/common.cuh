// Shared device helpers for the alignq_b200 kernels (sm_100a only).
//
// Numerics contract: the reference's forward is a chain of
// separately rounded fp32 ATen ops.  Every step of that chain is written here with the
// non-contracting intrinsics (__fmul_rn/__fadd_rn/__fsub_rn), so ptxas can never fuse two
// of them into an FMA, and with the same libdevice erff/expf ATen's CUDA kernels call.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>

#define ALIGNQ_NUM_SMS 148
#define ALIGNQ_BN_SLOTS 16      // copies of the per-channel fp64 accumulators of the fused BatchNorm kernels (bn_act.cu, conv_tc.cu)

enum { ALIGNQ_OK = 0, ALIGNQ_EINVAL = -1, ALIGNQ_EALIGN = -2, ALIGNQ_ERANGE = -3, ALIGNQ_ENOSPACE = -4 };

// every kernel launch of the library passes through here: error check + launch statistics
// (alignq_launch_count(); a relaxed atomic, the only process-global state of the library)
extern "C" void alignq_count_launch_(void);
#define ALIGNQ_LAUNCH_CHECK()                         \
  do {                                                \
    cudaError_t e__ = cudaGetLastError();             \
    if (e__ != cudaSuccess) return (int)e__;          \
    alignq_count_launch_();                           \
  } while (0)

namespace alignq {

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// The QAT step is a chain of 5-15 us kernels; each spends 1.5-4 us on launch ramp and a prologue that does not depend on
// its predecessor (TMEM allocation, barrier set-up, staging the quantized weights, per-channel parameters).  A kernel
// launched with cudaLaunchAttributeProgrammaticStreamSerialization may start as soon as every block of its predecessor
// has executed pdl_trigger() (or exited): it runs that prologue beside the predecessor's tail and then blocks in
// pdl_wait() until the predecessor has COMPLETED and its writes are visible (and, by induction along the chain, every
// kernel before it: a dependent cannot complete its own wait earlier).  Rules kept here: (1) before pdl_wait() a kernel
// touches no global memory at all -- except the convolutions, which stage the quantized weights; (2) those weights are
// written by the weight bank at the start of the step, and at least one normally launched kernel (the stem convolution,
// the statistics kernels, every library kernel: full barriers of the stream) lies between that write and the first
// dependent launch, so "my predecessor has started" implies "the weights are complete".
// Both instructions are no-ops in a kernel launched without the attribute / with nothing depending on it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// host side: launch `kernel` as a programmatic dependent of whatever precedes it on the stream (ALIGNQ_PDL=0: plain launch)
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ALIGNQ_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr float kInvSqrt2 = 0.70710677f;        // fp32(1/fp32(sqrt(2))): ATen multiplies by the reciprocal of a CPU scalar divisor
constexpr float kTwoOverSqrtPi = 1.1283791f;    // erf'(v) = 2/sqrt(pi) exp(-v^2)
constexpr float kInvSqrt2Pi = 0.3989423f;
constexpr float kLogSqrt2Pi = 0.9189385f;       // math.log(math.sqrt(2*math.pi))

// ---- the reference's fp32 op chain ------------------------------------------------------
// Normal(loc, scale).cdf(x) = 0.5 * (1 + erf((x - loc) * (1/scale) / sqrt(2)))
// The reference rounds (1 + e) and then halves it; halving is exact in binary floating point (no subnormals
// here: 1 + e >= 2^-24), so ONE fused multiply-add round(0.5 e + 0.5) returns the identical bits.  Likewise
// c*2 is exact, so fma(c, 2, -1) == round(round(c*2) - 1).  (Two instructions saved per element on kernels
// that are instruction-issue bound.)
__device__ __forceinline__ float half_one_plus(float e) { return __fmaf_rn(e, 0.5f, 0.5f); }
__device__ __forceinline__ float normal_cdf_std(float x) {             // loc = 0, scale = 1 (QA:97)
  float v = __fmul_rn(x, kInvSqrt2);                                   // (x-0)*1 is exact
  return half_one_plus(erff(v));
}
__device__ __forceinline__ float normal_cdf(float x, float loc, float rscale) {
  float u = __fmul_rn(__fsub_rn(x, loc), rscale);
  float v = __fmul_rn(u, kInvSqrt2);
  return half_one_plus(erff(v));
}
// variant A: c ; variant B/C: (c*2-1) [* act_range for activations]
__device__ __forceinline__ float sym_map(float c) { return __fmaf_rn(c, 2.0f, -1.0f); }

// round(p*n)/n with the divide done as ATen does it on CUDA (multiply by fp32 1/n)
__device__ __forceinline__ float quant_code(float p, float n) { return rintf(__fmul_rn(p, n)); }

// d/dx Phi((x-loc)/scale) without the 1/scale factor, mirroring erf's backward argument
__device__ __forceinline__ float gauss_kernel_from_v(float v) { return expf(-__fmul_rn(v, v)); }

// ---- vector memory helpers --------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- reductions -------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of two values; result valid in every thread.  scratch: 2*32 T's of smem.
template <typename T>
__device__ __forceinline__ void block_sum2(T& a, T& b, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();                       // scratch may still be read from a previous call
  if (lane == 0) { scratch[wid] = a; scratch[32 + wid] = b; }
  __syncthreads();
  a = (lane < nw) ? scratch[lane] : T(0);
  b = (lane < nw) ? scratch[32 + lane] : T(0);
  a = warp_sum(a);
  b = warp_sum(b);
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace alignq
