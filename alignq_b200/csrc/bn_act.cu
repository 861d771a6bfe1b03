// Fused BatchNorm2d -> CDF activation quantizer -> (ReLU), forward and backward, for NHWC
// (channels_last) fp32 activations -- SURVEY.md section 8(f) item 1, the step either side of the
// quantizer in every model file:
//   out = F.relu(self.act_q0(self.bn0(out)))        cdf_alignment/resnet-20-cifar-10/model/resnet.py:72,121-123
//   out = self.relu(self.act_q0(self.bn1(x)))       cdf_alignment/dense-cifar-10/model/densenet.py:32-34
// Un-fused this is cuDNN BN (12 B/elem) + quantizer (8) + ReLU (8) forward and the same again
// backward; fused it is 12 B/elem forward (stats pass + apply pass) and 28 B/elem backward.
//
// x is viewed as [R = B*H*W, C] with C contiguous.  Every thread owns one float4 of channels
// (c4 = tid % (C/4); the block size is a multiple of C/4), so per-channel accumulators live in
// registers and every global access is a coalesced 128-bit load.  Block totals are added to fp64
// accumulators with one atomic per value; the LAST block to finish (threadfence + ticket, no
// spinning) finalises the statistics and re-zeroes accumulators and ticket -- no extra launch, no memset.
// The backward of tensors up to ~24 MB runs as ONE cooperative launch instead (bnq_bwd_fused_kernel below).
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"
#include "bn_stat.cuh"
#include "../../include/alignq_b200.h"

namespace cg = cooperative_groups;

namespace alignq {

constexpr int BN_MAX_THREADS = 256;
constexpr int BN_MAX_GRID = 4 * ALIGNQ_NUM_SMS;
constexpr int BN_SLOTS = ALIGNQ_BN_SLOTS;      // accumulator copies: same-address fp64 atomics serialise in L2 (~15 ns each)

struct BnQ {
  float n, inv_n, ar, gscale;
  int a_bit, variant, relu;
};

__device__ __forceinline__ float bnq_quant(float z, const BnQ& q) {
  float c = normal_cdf_std(z);
  if (q.variant != 0) c = __fmul_rn(sym_map(c), q.ar);
  float v;
  if (q.a_bit == 1) v = (c > 0.f) ? 1.f : ((c < 0.f) ? -1.f : c);
  else v = __fmul_rn(rintf(__fmul_rn(c, q.n)), q.inv_n);
  if (q.variant == 0) v = __fmul_rn(sym_map(v), q.ar);
  return v;
}
__device__ __forceinline__ float bnq_finish(float v, float res, const BnQ& q) {       // (+ shortcut) -> ReLU
  v += res;
  return (q.relu && v < 0.f) ? 0.f : v;
}

struct Lane4 { float v[4]; };
__device__ __forceinline__ Lane4 ld4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  return Lane4{{t.x, t.y, t.z, t.w}};
}

// upstream gradient of a tensor with two consumers: gy (+ gy2 when the second consumer's gradient arrives separately,
// so that autograd's accumulate kernel `ga + gb` never runs -- model/fused.py:_GradFork)
__device__ __forceinline__ Lane4 ld4g(const float* gy, const float* gy2, int64_t o) {
  Lane4 a = ld4(gy + o);
  if (gy2) {
    const Lane4 b = ld4(gy2 + o);
#pragma unroll
    for (int j = 0; j < 4; ++j) a.v[j] += b.v[j];
  }
  return a;
}

// last-block ticket: returns true in every thread of the block that finishes last
__device__ __forceinline__ bool last_block(unsigned* counter, unsigned* smem_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    *smem_flag = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  const bool last = *smem_flag != 0u;
  if (last) __threadfence();
  return last;
}

// ---- peer exchange of the per-channel fp64 sums over NVLink (data-parallel SyncBN, no NCCL call, no extra launch) ----
// Every rank owns one buffer of the same layout in PEER-MAPPED memory (torch symmetric memory: cuMem handles exchanged
// once at start-up); `bufs` is a device array of the world's base pointers.  One exchange = one sequence number `seq`.
// Low-latency protocol (the scheme of NCCL's LL protocol): every 8-byte word carries 4 bytes of payload and the 4-byte
// sequence number, and an aligned 8-byte store is a single transaction -- so the data IS the flag: no
// __threadfence_system, no separate flag store, no round trip.
//   producer (the last block of the statistics / backward-reduce kernel, thread c): splits its two fp64 sums of
//     channel c into four tagged words and stores them into slot (seq % 4), row `rank`, of EVERY rank's buffer;
//   consumer (the SAME thread, right afterwards): polls the four words of channel c in every rank's row of its OWN
//     buffer until their tag equals seq, adds the world's sums in rank order -- bit-identical totals on every rank --
//     and finishes the global statistics exactly as the single-device kernel does.  The apply kernel that follows is
//     the ordinary one.  (Earlier versions: every block of the apply kernel waiting and redoing the fp64 arithmetic was
//     slower than NCCL; data + fence.sys + flag cost ~7.5 us per exchange.)
// A slot is rewritten 4 exchanges later; a rank can be at most one exchange ahead of the slowest one (its consumer
// needs everybody's producer), so nothing is needed in the other direction.  The sequence counter lives in device
// memory and is advanced by the producer, so the kernels are CUDA-graph replayable.
constexpr int PEER_MAX = 8, PEER_SLOTS = 4, PEER_MAXC = 1024;
constexpr size_t PEER_BYTES = (size_t)PEER_SLOTS * PEER_MAX * 4 * PEER_MAXC * sizeof(unsigned long long);
struct PeerCtx {
  void* const* bufs;      // nullptr: single-device path
  unsigned* seq;
  int rank, world;
  double rows_global;
};
__device__ __forceinline__ unsigned long long* ll_row(void* base, unsigned seq, int r) {
  return reinterpret_cast<unsigned long long*>(base) + ((size_t)(seq % PEER_SLOTS) * PEER_MAX + r) * 4 * PEER_MAXC;
}
__device__ __forceinline__ void ll_put(const PeerCtx& pc, unsigned seq, int c, double S, double SS) {
  const unsigned long long a = (unsigned long long)__double_as_longlong(S), b = (unsigned long long)__double_as_longlong(SS);
  const unsigned long long tag = (unsigned long long)seq << 32;
  const unsigned long long w0 = tag | (a & 0xffffffffull), w1 = tag | (a >> 32), w2 = tag | (b & 0xffffffffull), w3 = tag | (b >> 32);
  for (int r = 0; r < pc.world; ++r) {
    unsigned long long* w = ll_row(pc.bufs[r], seq, pc.rank) + 4 * c;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(w), "l"(w0), "l"(w1) : "memory");
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(w + 2), "l"(w2), "l"(w3) : "memory");
  }
}
// totals of channel c over the world, rank order; spins (bounded) until every rank's words of exchange `seq` arrived
__device__ __forceinline__ void ll_get(const PeerCtx& pc, unsigned seq, int C, int c, double& S, double& SS) {
  void* mine = pc.bufs[pc.rank];
  S = 0.0; SS = 0.0;
  const long long t0 = clock64();
  for (int r = 0; r < pc.world; ++r) {
    const unsigned long long* w = ll_row(mine, seq, r) + 4 * c;
    unsigned long long v0, v1, v2, v3;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(w) : "memory");
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v2), "=l"(v3) : "l"(w + 2) : "memory");
      if ((unsigned)(v0 >> 32) == seq && (unsigned)(v1 >> 32) == seq && (unsigned)(v2 >> 32) == seq && (unsigned)(v3 >> 32) == seq) break;
      if (clock64() - t0 > 8000000000LL) __trap();           // a rank died or the ranks' layer order diverged
    }
    S += __longlong_as_double((long long)((v0 & 0xffffffffull) | (v1 << 32)));
    SS += __longlong_as_double((long long)((v2 & 0xffffffffull) | (v3 << 32)));
  }
}

// Reduce two float4 accumulators over the threads of the block that share c4 (shared memory, fp64),
// then add the block totals to the global fp64 accumulators acc[C][2] with one atomic per value.
// (Summation order across blocks is not fixed; in fp64 that is ~1e-16 relative, far below the fp32
// rounding of the finished statistics.)  The accumulators are zero on entry and the last block
// re-zeroes them, so no memset is ever launched.
__device__ __forceinline__ void block_accumulate(const float (&a)[4], const float (&b)[4], int C4, int k,
                                                 float* sh, double* acc /* [C][2] */) {
  const int t = threadIdx.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[t * 8 + j] = a[j]; sh[t * 8 + 4 + j] = b[j]; }
  __syncthreads();
  for (int idx = t; idx < 8 * C4; idx += blockDim.x) {      // one (channel quad, value) pair per thread
    const int c4 = idx >> 3, j = idx & 7;
    double v = 0.0;
#pragma unroll 8
    for (int r = 0; r < k; ++r) v += (double)sh[(r * C4 + c4) * 8 + j];
    atomicAdd(acc + ((size_t)(blockIdx.x % BN_SLOTS) * 4 * C4 + c4 * 4 + (j & 3)) * 2 + (j >> 2), v);
  }
}

// ---- forward pass 1: batch statistics -------------------------------------------------------------
__global__ void __launch_bounds__(BN_MAX_THREADS)
bnq_stats_kernel(const float* __restrict__ x, int64_t R, int C, float* __restrict__ running_mean,
                 float* __restrict__ running_var, float momentum, float bn_eps, float* __restrict__ save_mean,
                 float* __restrict__ save_invstd, double* __restrict__ ws, unsigned* __restrict__ counter,
                 long long* __restrict__ num_batches_tracked, double* __restrict__ sums_out, PeerCtx pc) {
  extern __shared__ float sh[];
  __shared__ unsigned flag;

  pdl_trigger();                                            // the apply kernel launches early and waits for this grid
  pdl_wait();                                               // (and this one beside the tail of the convolution before it)
  const int C4 = C >> 2, k = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, rsub = threadIdx.x / C4;
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * k;
  const Lane4 zero4 = {{0.f, 0.f, 0.f, 0.f}};
  for (int64_t r = (int64_t)blockIdx.x * k + rsub; r < R; r += 4 * stride) {   // 4 predicated 128-bit loads in flight,
    Lane4 v[4];                                                                // no serial tail
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (r + u * stride < R) ? ld4(x + (r + u * stride) * C + 4 * c4) : zero4;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[j] += v[u].v[j]; ss[j] = fmaf(v[u].v[j], v[u].v[j], ss[j]); }
  }
  block_accumulate(s, ss, C4, k, sh, ws);
  if (last_block(counter, &flag)) {
    const unsigned pseq = pc.bufs ? *reinterpret_cast<volatile unsigned*>(pc.seq) + 1u : 0u;
    auto finalize = [&](int c, double S, double SS, double Rd) {
      const double mean = S / Rd;
      double var = SS / Rd - mean * mean;                   // biased: what BN normalises with
      var = var < 0.0 ? 0.0 : var;
      save_mean[c] = (float)mean;
      save_invstd[c] = (float)(1.0 / sqrt(var + (double)bn_eps));
      if (running_mean) {
        const double unbiased = Rd > 1.0 ? var * Rd / (Rd - 1.0) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
      }
    };
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double S = 0.0, SS = 0.0;
#pragma unroll
      for (int sl = 0; sl < BN_SLOTS; ++sl) {                // fixed order over the accumulator copies
        double* a = ws + ((size_t)sl * C + c) * 2;
        S += __ldcg(a); SS += __ldcg(a + 1);
        a[0] = 0.0; a[1] = 0.0;                              // re-arm the accumulators
      }
      if (pc.bufs) {                                         // peer exchange: my sums into every rank's buffer, then the
        ll_put(pc, pseq, c, S, SS);                          // world's sums out of mine: global statistics, still
        ll_get(pc, pseq, C, c, S, SS);                       // inside the statistics kernel
        finalize(c, S, SS, pc.rows_global);
        continue;
      }
      if (sums_out) {                                        // data-parallel SyncBN through NCCL: the caller all-reduces (sum,
        sums_out[c] = S; sums_out[C + c] = SS;               // sum of squares), bnq_sync_finalize_kernel finishes
        continue;
      }
      finalize(c, S, SS, (double)R);
    }
    __syncthreads();                                        // every thread has read pseq
    if (threadIdx.x == 0) {
      if (pc.bufs) *pc.seq = pseq;
      *counter = 0u;                                        // re-arm for the next launch
      if (num_batches_tracked && !sums_out) *num_batches_tracked += 1;   // BatchNorm2d.num_batches_tracked
    }
  }
}

// SyncBN: statistics of the GLOBAL batch from the all-reduced fp64 (sum, sum of squares) and the global row count
__global__ void bnq_sync_finalize_kernel(const double* __restrict__ sums, double Rg, int C, float* __restrict__ running_mean,
                                         float* __restrict__ running_var, float momentum, float bn_eps,
                                         float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                         long long* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const double mean = sums[c] / Rg;
    double var = sums[C + c] / Rg - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    save_mean[c] = (float)mean;
    save_invstd[c] = (float)(1.0 / sqrt(var + (double)bn_eps));
    if (running_mean) {
      const double unbiased = Rg > 1.0 ? var * Rg / (Rg - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  }
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
}

// SyncBN backward: mean(g_z), mean(g_z xhat) over the GLOBAL batch from the all-reduced sums
__global__ void bnq_sync_coef_kernel(const double* __restrict__ sums, double Rg, int C, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { coef[2 * c] = (float)(sums[c] / Rg); coef[2 * c + 1] = (float)(sums[C + c] / Rg); }
}

// eval mode: statistics are the running ones
__global__ void bnq_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float bn_eps, int C, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { save_mean[c] = running_mean[c]; save_invstd[c] = 1.0f / sqrtf(running_var[c] + bn_eps); }
}

// ---- forward pass 2: normalise, quantise, ReLU ---------------------------------------------------------
__global__ void __launch_bounds__(BN_MAX_THREADS)
bnq_apply_kernel(const float* __restrict__ x, int64_t R, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                 BnQ q, const float* __restrict__ residual, float* __restrict__ y) {
  pdl_trigger();                                            // a convolution that follows may run its prologue beside us
  pdl_wait();                                               // (and we beside the tail of the kernel before us)
  const int C4 = C >> 2, k = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, rsub = threadIdx.x / C4;
  float m[4], is[4], g[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * c4 + j;
    m[j] = mean[c]; is[j] = invstd[c]; g[j] = gamma ? gamma[c] : 1.f; b[j] = beta ? beta[c] : 0.f;
  }
  for (int64_t r = (int64_t)blockIdx.x * k + rsub; r < R; r += (int64_t)gridDim.x * k) {
    const Lane4 v = ld4(x + r * C + 4 * c4);
    Lane4 rs = {{0.f, 0.f, 0.f, 0.f}};
    if (residual) rs = ld4(residual + r * C + 4 * c4);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = bnq_finish(bnq_quant(fmaf((v.v[j] - m[j]) * is[j], g[j], b[j]), q), rs.v[j], q);
    *reinterpret_cast<float4*>(y + r * C + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// g_z = gy * [y > 0 if relu] * gscale * exp(-(z/sqrt2)^2): the straight-through quantizer + ReLU backward
__device__ __forceinline__ float bnq_gz(float z, float gy, float yv, const BnQ& q) {
  return bnq_gz_raw(z, gy, yv, q.gscale, q.relu);               // shared with the convolution epilogue (bn_stat.cuh)
}

// ---- backward pass 1: d beta = sum g_z, d gamma = sum g_z * xhat ------------------------------------------
__global__ void __launch_bounds__(BN_MAX_THREADS)
bnq_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                      const float* __restrict__ gy2, int64_t R, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ mean, const float* __restrict__ invstd, BnQ q,
                      float* __restrict__ ggamma, float* __restrict__ gbeta, float* __restrict__ coef /* [C][2] */,
                      double* __restrict__ ws, unsigned* __restrict__ counter, double* __restrict__ sums_out, PeerCtx pc) {
  extern __shared__ float sh[];
  __shared__ unsigned flag;

  pdl_trigger();
  pdl_wait();
  const int C4 = C >> 2, k = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, rsub = threadIdx.x / C4;
  float m[4], is[4], g[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * c4 + j;
    m[j] = mean[c]; is[j] = invstd[c]; g[j] = gamma ? gamma[c] : 1.f; b[j] = beta ? beta[c] : 0.f;
  }
  float db[4] = {0, 0, 0, 0}, dg[4] = {0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * k;
  int64_t r = (int64_t)blockIdx.x * k + rsub;
  auto body = [&](const Lane4& xv, const Lane4& gv, const Lane4& yv) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (xv.v[j] - m[j]) * is[j];
      const float gz = bnq_gz(fmaf(xh, g[j], b[j]), gv.v[j], yv.v[j], q);
      db[j] += gz;
      dg[j] = fmaf(gz, xh, dg[j]);
    }
  };
  const Lane4 ones = {{1.f, 1.f, 1.f, 1.f}};
  for (; r < R; r += 4 * stride) {                          // 4 rows x 3 tensors = 12 predicated loads in flight
    Lane4 xs_[4], gs_[4], ys_[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ok[u] = r + u * stride < R;
      const int64_t o = (r + u * stride) * C + 4 * c4;
      xs_[u] = ok[u] ? ld4(x + o) : ones;
      gs_[u] = ok[u] ? ld4g(gy, gy2, o) : ones;
      ys_[u] = (ok[u] && q.relu) ? ld4(y + o) : ones;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (ok[u]) body(xs_[u], gs_[u], ys_[u]);
  }
  block_accumulate(db, dg, C4, k, sh, ws);
  if (last_block(counter, &flag)) {
    const unsigned pseq = pc.bufs ? *reinterpret_cast<volatile unsigned*>(pc.seq) + 1u : 0u;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double S = 0.0, SS = 0.0;
#pragma unroll
      for (int sl = 0; sl < BN_SLOTS; ++sl) {                // fixed order over the accumulator copies
        double* a = ws + ((size_t)sl * C + c) * 2;
        S += __ldcg(a); SS += __ldcg(a + 1);
        a[0] = 0.0; a[1] = 0.0;                              // re-arm the accumulators
      }
      if (gbeta) gbeta[c] = (float)S;                        // affine gradients: LOCAL sums in either mode (a data-parallel
      if (ggamma) ggamma[c] = (float)SS;                     // step averages them with the other parameter gradients)
      if (pc.bufs) {
        ll_put(pc, pseq, c, S, SS);
        ll_get(pc, pseq, C, c, S, SS);
        coef[2 * c] = (float)(S / pc.rows_global);
        coef[2 * c + 1] = (float)(SS / pc.rows_global);
        continue;
      }
      if (sums_out) { sums_out[c] = S; sums_out[C + c] = SS; continue; }
      coef[2 * c] = (float)(S / (double)R);
      coef[2 * c + 1] = (float)(SS / (double)R);
    }
    __syncthreads();                                        // every thread has read pseq
    if (threadIdx.x == 0) {
      if (pc.bufs) *pc.seq = pseq;
      *counter = 0u;
    }
  }
}

// ---- backward pass 2: gx = gamma * invstd * (g_z - mean(g_z) - xhat * mean(g_z xhat))  [training]
//                        gx = gamma * invstd * g_z                                           [eval] --------
__global__ void __launch_bounds__(BN_MAX_THREADS)
bnq_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                     const float* __restrict__ gy2, int64_t R, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ coef,
                     int training, BnQ q, float* __restrict__ gx, float* __restrict__ g_residual) {
  pdl_trigger();
  pdl_wait();
  const int C4 = C >> 2, k = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, rsub = threadIdx.x / C4;
  float m[4], is[4], g[4], b[4], k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * c4 + j;
    m[j] = mean[c]; is[j] = invstd[c]; g[j] = gamma ? gamma[c] : 1.f; b[j] = beta ? beta[c] : 0.f;
    k1[j] = training ? coef[2 * c] : 0.f; k2[j] = training ? coef[2 * c + 1] : 0.f;
  }
  for (int64_t r = (int64_t)blockIdx.x * k + rsub; r < R; r += (int64_t)gridDim.x * k) {
    const int64_t o = r * C + 4 * c4;
    const Lane4 xv = ld4(x + o), gv = ld4g(gy, gy2, o);
    Lane4 yv = {{1.f, 1.f, 1.f, 1.f}};
    if (q.relu) yv = ld4(y + o);
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (xv.v[j] - m[j]) * is[j];
      const float gz = bnq_gz(fmaf(xh, g[j], b[j]), gv.v[j], yv.v[j], q);
      out[j] = g[j] * is[j] * (gz - k1[j] - xh * k2[j]);
    }
    *reinterpret_cast<float4*>(gx + o) = make_float4(out[0], out[1], out[2], out[3]);
    if (g_residual) {                                       // the shortcut sees the ReLU-masked upstream gradient
      float gr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gr[j] = (q.relu && !(yv.v[j] > 0.f)) ? 0.f : gv.v[j];
      *reinterpret_cast<float4*>(g_residual + o) = make_float4(gr[0], gr[1], gr[2], gr[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Single-launch backward for small and mid-size tensors (up to ~24 MB: every layer of ResNet-20, most of DenseNet-40
// and MobileNet-v2), where the reduce + apply pair above costs 9-35 us and much of it is latency: launch,
// last-block ticket, finalize, launch, per-channel parameter loads.  One COOPERATIVE launch does reduce -> grid
// barrier -> apply; every block re-derives the per-channel coefficients from the fp64 accumulators after the
// barrier (C <= 256: a few KB from L2), so there is no ticket and no serial finalize.  Measured (graph replay):
// [128,16,32,32] 15.4 -> 12.5 us, [128,168,16,16] 35.1 -> 24.9 us, [256,24,32,32] 34.7 -> 27.8 us; above ~24 MB
// it is a wash, and for the FORWARD pair the same fusion was slower (the apply pass prefers twice the blocks), so
// only the backward uses it.  The accumulators are double-buffered by a device-side epoch (counter[1]): a launch
// uses set (epoch & 1), which the previous launch on the same workspace zeroed, and zeroes the other one; block 0
// bumps the epoch after the barrier.  Its two sets live behind the set the two-kernel paths use.
__device__ __forceinline__ void fused_zero_other(double* other, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) other[i] = 0.0;
}

__global__ void __launch_bounds__(BN_MAX_THREADS)
bnq_bwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gy,
                     const float* __restrict__ gy2, int64_t R, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ mean, const float* __restrict__ invstd, int training, BnQ q,
                     float* __restrict__ gx, float* __restrict__ g_residual, float* __restrict__ ggamma,
                     float* __restrict__ gbeta, double* __restrict__ ws, unsigned* __restrict__ epoch, PeerCtx pc,
                     float* __restrict__ coefg) {
  extern __shared__ float sh[];
  cg::grid_group grid = cg::this_grid();
  pdl_trigger();
  const int C4 = C >> 2, k = blockDim.x / C4;
  const int c4 = threadIdx.x % C4, rsub = threadIdx.x / C4;
  const unsigned ep = *reinterpret_cast<volatile unsigned*>(epoch);
  const int setn = BN_SLOTS * C * 2;
  double* acc = ws + (size_t)(1u + (ep & 1u)) * setn;             // sets 1 and 2; set 0 belongs to the two-kernel paths
  fused_zero_other(ws + (size_t)(1u + ((ep & 1u) ^ 1u)) * setn, setn);
  float m[4], is[4], g[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = 4 * c4 + j;
    m[j] = mean[c]; is[j] = invstd[c]; g[j] = gamma ? gamma[c] : 1.f; b[j] = beta ? beta[c] : 0.f;
  }
  // ---- phase 1: sum g_z, sum g_z * xhat ----
  float db[4] = {0, 0, 0, 0}, dg[4] = {0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * k;
  const Lane4 ones = {{1.f, 1.f, 1.f, 1.f}};
  for (int64_t r = (int64_t)blockIdx.x * k + rsub; r < R; r += 4 * stride) {
    Lane4 xs_[4], gs_[4], ys_[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ok[u] = r + u * stride < R;
      const int64_t o = (r + u * stride) * C + 4 * c4;
      xs_[u] = ok[u] ? ld4(x + o) : ones;
      gs_[u] = ok[u] ? ld4g(gy, gy2, o) : ones;
      ys_[u] = (ok[u] && q.relu) ? ld4(y + o) : ones;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (xs_[u].v[j] - m[j]) * is[j];
        const float gz = bnq_gz(fmaf(xh, g[j], b[j]), gs_[u].v[j], ys_[u].v[j], q);
        db[j] += gz;
        dg[j] = fmaf(gz, xh, dg[j]);
      }
    }
  }
  block_accumulate(db, dg, C4, k, sh, acc);
  __threadfence();
  grid.sync();
  float* sm_k1 = sh;
  float* sm_k2 = sh + C;
  if (pc.bufs) {
    // data-parallel global statistics: block 0 exchanges the sums with the other ranks over NVLink peer memory (see
    // PeerCtx) and leaves the global means in coefg; a second grid barrier hands them to everybody
    if (blockIdx.x == 0) {
      const unsigned pseq = *reinterpret_cast<volatile unsigned*>(pc.seq) + 1u;
      for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double S = 0.0, SS = 0.0;
#pragma unroll
        for (int sl = 0; sl < BN_SLOTS; ++sl) {
          const double* a = acc + ((size_t)sl * C + c) * 2;
          S += __ldcg(a); SS += __ldcg(a + 1);
        }
        if (gbeta) gbeta[c] = (float)S;                       // affine gradients: local sums
        if (ggamma) ggamma[c] = (float)SS;
        ll_put(pc, pseq, c, S, SS);
        ll_get(pc, pseq, C, c, S, SS);
        coefg[2 * c] = (float)(S / pc.rows_global);
        coefg[2 * c + 1] = (float)(SS / pc.rows_global);
      }
      __syncthreads();
      if (threadIdx.x == 0) { *epoch = ep + 1u; *pc.seq = pseq; }
      __threadfence();
    }
    grid.sync();
    for (int c = threadIdx.x; c < C; c += blockDim.x) { sm_k1[c] = __ldcg(coefg + 2 * c); sm_k2[c] = __ldcg(coefg + 2 * c + 1); }
  } else
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double S = 0.0, SS = 0.0;
#pragma unroll
    for (int sl = 0; sl < BN_SLOTS; ++sl) {
      const double* a = acc + ((size_t)sl * C + c) * 2;
      S += __ldcg(a); SS += __ldcg(a + 1);
    }
    sm_k1[c] = (float)(S / (double)R);
    sm_k2[c] = (float)(SS / (double)R);
    if (blockIdx.x == 0) {
      if (gbeta) gbeta[c] = (float)S;
      if (ggamma) ggamma[c] = (float)SS;
    }
  }
  if (!pc.bufs && blockIdx.x == 0 && threadIdx.x == 0) *epoch = ep + 1u;
  __syncthreads();
  float k1[4], k2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k1[j] = training ? sm_k1[4 * c4 + j] : 0.f;
    k2[j] = training ? sm_k2[4 * c4 + j] : 0.f;
  }
  // ---- phase 2: gx (and the shortcut's gradient) ----
  for (int64_t r = (int64_t)blockIdx.x * k + rsub; r < R; r += stride) {
    const int64_t o = r * C + 4 * c4;
    const Lane4 xv = ld4(x + o), gv = ld4g(gy, gy2, o);
    Lane4 yv = {{1.f, 1.f, 1.f, 1.f}};
    if (q.relu) yv = ld4(y + o);
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float xh = (xv.v[j] - m[j]) * is[j];
      const float gz = bnq_gz(fmaf(xh, g[j], b[j]), gv.v[j], yv.v[j], q);
      out[j] = g[j] * is[j] * (gz - k1[j] - xh * k2[j]);
    }
    *reinterpret_cast<float4*>(gx + o) = make_float4(out[0], out[1], out[2], out[3]);
    if (g_residual) {
      float gr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) gr[j] = (q.relu && !(yv.v[j] > 0.f)) ? 0.f : gv.v[j];
      *reinterpret_cast<float4*>(g_residual + o) = make_float4(gr[0], gr[1], gr[2], gr[3]);
    }
  }
}

struct BnLaunch { int threads, grid, k; size_t smem; };
static BnLaunch bn_launch(int64_t R, int C, int max_grid = BN_MAX_GRID) {
  BnLaunch L;
  const int C4 = C / 4;
  L.k = BN_MAX_THREADS / C4;
  L.threads = L.k * C4;
  int64_t g = (R + (int64_t)L.k * 4 - 1) / ((int64_t)L.k * 4);          // >= 4 rows per thread
  if (g > max_grid) g = max_grid;
  if (g < 1) g = 1;
  L.grid = (int)g;
  L.smem = (size_t)L.threads * 8 * sizeof(float);
  return L;
}

// The single-launch path: small tensors, narrow layers, and a grid that is co-resident (cooperative launch).
static bool bn_use_fused(int64_t R, int C, int64_t max_elems) { return C <= 256 && R * (int64_t)C <= max_elems; }
// ALIGNQ_BN_COOP_PER_SM = n > 0 switches the single-launch (cooperative) backward on, with at most n co-resident blocks
// per SM.  Default: off.  A cooperative grid starts only when ALL its blocks fit beside whatever else is running, so in
// the training step it waited for the weight-gradient kernels of the side streams to drain (every bnq_bwd_fused launch
// began 0.6 us before the preceding conv3x3_wgrad kernel ended: profiles/r02_timeline_resnet20_*.txt), the driver adds
// a memcpy node after every cooperative launch of a CUDA graph, and a cooperative launch cannot be a programmatic
// dependent.  Measured on the ResNet-20 step: 1.205 ms with it, 1.115 ms with the reduce + apply pair.
static int bn_coop_cap() {
  const char* e = getenv("ALIGNQ_BN_COOP_PER_SM");
  return e ? atoi(e) : 0;
}
template <typename K>
static int bn_coop_grid(K kernel, const BnLaunch& L) {
  int per_sm = 0;
  if (bn_coop_cap() == 0) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, L.threads, L.smem) != cudaSuccess || per_sm < 1) return 0;
  if (bn_coop_cap() > 0 && per_sm > bn_coop_cap()) per_sm = bn_coop_cap();
  const int cap = per_sm * ALIGNQ_NUM_SMS;
  return L.grid < cap ? L.grid : cap;
}

static BnQ make_bnq(int a_bit, float act_range, int variant, int relu) {
  BnQ q;
  q.a_bit = a_bit; q.variant = variant == 0 ? 0 : 1; q.relu = relu; q.ar = act_range;
  q.n = (float)((1ull << a_bit) - 1);
  q.inv_n = 1.0f / q.n;
  q.gscale = 2.0f * act_range * kInvSqrt2Pi;
  return q;
}

static int bn_check(int64_t R, int C, int a_bit, int variant, const void* a, const void* b, const void* c) {
  if (R < 1 || C < 4 || (C & 3) || C > 4 * BN_MAX_THREADS || a_bit < 1 || a_bit > 31 || variant < 0 || variant > 2) return ALIGNQ_EINVAL;
  if (!aligned16(a) || !aligned16(b) || (c && !aligned16(c))) return ALIGNQ_EALIGN;
  return ALIGNQ_OK;
}

}  // namespace alignq

using namespace alignq;

extern "C" size_t alignq_bn_act_ws_doubles(int C) {
  if (C <= 0) return 0;
  // three fp64 accumulator sets [3][slots][C][2] (ZERO before first use: one for the two-kernel paths, two that the
  // single-launch backward alternates between) + [C][2] float coefficients
  return (size_t)3 * BN_SLOTS * C * 2 + (size_t)C;
}

extern "C" int alignq_bn_act_fwd(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, float momentum, float bn_eps, int training,
                                 int a_bit, float act_range, int variant, int relu, const float* residual, float* y,
                                 float* save_mean, float* save_invstd, double* ws, uint32_t* counter,
                                 int64_t* num_batches_tracked, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, y, residual);
  if (rc) return rc;
  if (!x || !y || !save_mean || !save_invstd || !ws || !counter) return ALIGNQ_EINVAL;
  if (!training && (!running_mean || !running_var)) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const BnLaunch L = bn_launch(rows, C);
  if (training) {
    { cudaError_t pe_ = launch_pdl(bnq_stats_kernel, dim3(L.grid), dim3(L.threads), L.smem, s, x, rows, C, running_mean, running_var, momentum, bn_eps,
                                                       save_mean, save_invstd, ws, counter,
                                                       reinterpret_cast<long long*>(num_batches_tracked), nullptr, PeerCtx{}); if (pe_ != cudaSuccess) return (int)pe_; }
  } else {
    bnq_eval_stats_kernel<<<(C + 255) / 256, 256, 0, s>>>(running_mean, running_var, bn_eps, C, save_mean, save_invstd);
  }
  ALIGNQ_LAUNCH_CHECK();
  { cudaError_t pe_ = launch_pdl(bnq_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, rows, C, gamma, beta, save_mean, save_invstd,
                                                    make_bnq(a_bit, act_range, variant, relu), residual, y); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_apply(const float* x, int64_t rows, int C, const float* gamma, const float* beta,
                                   const float* save_mean, const float* save_invstd, int a_bit, float act_range, int variant,
                                   int relu, const float* residual, float* y, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, y, residual);
  if (rc) return rc;
  if (!x || !y || !save_mean || !save_invstd) return ALIGNQ_EINVAL;
  const BnLaunch L = bn_launch(rows, C);
  { cudaError_t pe_ = launch_pdl(bnq_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, reinterpret_cast<cudaStream_t>(stream),
      x, rows, C, gamma, beta, save_mean, save_invstd, make_bnq(a_bit, act_range, variant, relu), residual, y); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_bwd_sum(const float* x, const float* y, const float* gy, const float* gy2, int64_t rows, int C,
                                     const float* gamma, const float* beta, const float* save_mean,
                                     const float* save_invstd, int training, int a_bit, float act_range, int variant,
                                     int relu, float* gx, float* g_residual, float* ggamma, float* gbeta, double* ws,
                                     uint32_t* counter, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, gy, gx);
  if (rc) return rc;
  if (!x || !gy || !gx || !save_mean || !save_invstd || !ws || !counter || (relu && !y)) return ALIGNQ_EINVAL;
  if ((relu && !aligned16(y)) || (g_residual && !aligned16(g_residual)) || (gy2 && !aligned16(gy2))) return ALIGNQ_EALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const BnLaunch L = bn_launch(rows, C);
  BnQ q = make_bnq(a_bit, act_range, variant, relu);
  if (bn_use_fused(rows, C, (int64_t)6 << 20)) {
    const int grid = bn_coop_grid(bnq_bwd_fused_kernel, L);
    if (grid > 0) {
      uint32_t* epoch = counter + 1;
      PeerCtx nopeer{};
      float* nocoef = nullptr;
      void* args[] = {(void*)&x, (void*)&y, (void*)&gy, (void*)&gy2, (void*)&rows, (void*)&C, (void*)&gamma, (void*)&beta, (void*)&save_mean,
                      (void*)&save_invstd, (void*)&training, (void*)&q, (void*)&gx, (void*)&g_residual, (void*)&ggamma,
                      (void*)&gbeta, (void*)&ws, (void*)&epoch, (void*)&nopeer, (void*)&nocoef};
      if (cudaLaunchCooperativeKernel((void*)bnq_bwd_fused_kernel, dim3(grid), dim3(L.threads), args, L.smem, s) == cudaSuccess) {
        ALIGNQ_LAUNCH_CHECK();
        return ALIGNQ_OK;
      }
      (void)cudaGetLastError();
    }
  }
  float* coef = reinterpret_cast<float*>(ws + (size_t)3 * BN_SLOTS * C * 2);       // [C][2] floats after the accumulators
  { cudaError_t pe_ = launch_pdl(bnq_bwd_reduce_kernel, dim3(L.grid), dim3(L.threads), L.smem, s, x, y, gy, gy2, rows, C, gamma, beta, save_mean, save_invstd, q,
                                                          ggamma, gbeta, coef, ws, counter, nullptr, PeerCtx{}); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  { cudaError_t pe_ = launch_pdl(bnq_bwd_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, y, gy, gy2, rows, C, gamma, beta, save_mean, save_invstd, coef,
                                                        training, q, gx, g_residual); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_bwd(const float* x, const float* y, const float* gy, int64_t rows, int C,
                                 const float* gamma, const float* beta, const float* save_mean,
                                 const float* save_invstd, int training, int a_bit, float act_range, int variant,
                                 int relu, float* gx, float* g_residual, float* ggamma, float* gbeta, double* ws,
                                 uint32_t* counter, alignq_stream_t stream) {
  return alignq_bn_act_bwd_sum(x, y, gy, nullptr, rows, C, gamma, beta, save_mean, save_invstd, training, a_bit, act_range,
                               variant, relu, gx, g_residual, ggamma, gbeta, ws, counter, stream);
}

// The apply pass alone: the reduce pass already ran in the epilogue of the data-gradient convolution that produced gy
// (alignq_conv3x3_bwd_data_bnreduce left mean(g_z), mean(g_z xhat) in the layer's workspace).
extern "C" int alignq_bn_act_bwd_apply(const float* x, const float* y, const float* gy, int64_t rows, int C,
                                       const float* gamma, const float* beta, const float* save_mean,
                                       const float* save_invstd, int a_bit, float act_range, int variant, int relu,
                                       float* gx, float* g_residual, double* ws, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, gy, gx);
  if (rc) return rc;
  if (!x || !gy || !gx || !save_mean || !save_invstd || !ws || (relu && !y)) return ALIGNQ_EINVAL;
  if ((relu && !aligned16(y)) || (g_residual && !aligned16(g_residual))) return ALIGNQ_EALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const BnLaunch L = bn_launch(rows, C);
  const float* coef = reinterpret_cast<const float*>(ws + (size_t)3 * BN_SLOTS * C * 2);
  const float* nogy2 = nullptr;
  cudaError_t pe = launch_pdl(bnq_bwd_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, y, gy, nogy2, rows, C, gamma, beta,
                              save_mean, save_invstd, coef, 1, make_bnq(a_bit, act_range, variant, relu), gx, g_residual);
  if (pe != cudaSuccess) return (int)pe;
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

// ---- data-parallel SyncBN: the same kernels cut at the point where the ranks' fp64 sums are all-reduced ----------
extern "C" int alignq_bn_act_sync_stats(const float* x, int64_t rows, int C, double* sums, double* ws, uint32_t* counter,
                                        alignq_stream_t stream) {
  if (rows < 1 || C < 4 || (C & 3) || C > 4 * BN_MAX_THREADS || !x || !sums || !ws || !counter) return ALIGNQ_EINVAL;
  if (!aligned16(x)) return ALIGNQ_EALIGN;
  const BnLaunch L = bn_launch(rows, C);
  { cudaError_t pe_ = launch_pdl(bnq_stats_kernel, dim3(L.grid), dim3(L.threads), L.smem, reinterpret_cast<cudaStream_t>(stream),
      x, rows, C, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, ws, counter, nullptr, sums, PeerCtx{}); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_sync_apply(const float* x, int64_t rows, int64_t rows_global, int C, const double* sums,
                                        const float* gamma, const float* beta, float* running_mean, float* running_var,
                                        float momentum, float bn_eps, int a_bit, float act_range, int variant, int relu,
                                        const float* residual, float* y, float* save_mean, float* save_invstd,
                                        int64_t* num_batches_tracked, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, y, residual);
  if (rc) return rc;
  if (!x || !y || !sums || !save_mean || !save_invstd || rows_global < rows) return ALIGNQ_EINVAL;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  bnq_sync_finalize_kernel<<<(C + 255) / 256, 256, 0, s>>>(sums, (double)rows_global, C, running_mean, running_var, momentum,
                                                            bn_eps, save_mean, save_invstd,
                                                            reinterpret_cast<long long*>(num_batches_tracked));
  ALIGNQ_LAUNCH_CHECK();
  const BnLaunch L = bn_launch(rows, C);
  { cudaError_t pe_ = launch_pdl(bnq_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, rows, C, gamma, beta, save_mean, save_invstd,
                                                    make_bnq(a_bit, act_range, variant, relu), residual, y); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_sync_bwd_reduce(const float* x, const float* y, const float* gy, int64_t rows, int C,
                                             const float* gamma, const float* beta, const float* save_mean,
                                             const float* save_invstd, int a_bit, float act_range, int variant, int relu,
                                             double* sums, float* ggamma, float* gbeta, double* ws, uint32_t* counter,
                                             alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, gy, nullptr);
  if (rc) return rc;
  if (!x || !gy || !sums || !save_mean || !save_invstd || !ws || !counter || (relu && !y)) return ALIGNQ_EINVAL;
  if (relu && !aligned16(y)) return ALIGNQ_EALIGN;
  const BnLaunch L = bn_launch(rows, C);
  float* coef = reinterpret_cast<float*>(ws + (size_t)3 * BN_SLOTS * C * 2);
  { cudaError_t pe_ = launch_pdl(bnq_bwd_reduce_kernel, dim3(L.grid), dim3(L.threads), L.smem, reinterpret_cast<cudaStream_t>(stream),
      x, y, gy, nullptr, rows, C, gamma, beta, save_mean, save_invstd, make_bnq(a_bit, act_range, variant, relu), ggamma, gbeta,
      coef, ws, counter, sums, PeerCtx{}); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_sync_bwd_apply(const float* x, const float* y, const float* gy, int64_t rows, int64_t rows_global,
                                            int C, const float* gamma, const float* beta, const float* save_mean,
                                            const float* save_invstd, int a_bit, float act_range, int variant, int relu,
                                            const double* sums, float* gx, float* g_residual, double* ws,
                                            alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, gy, gx);
  if (rc) return rc;
  if (!x || !gy || !gx || !sums || !save_mean || !save_invstd || !ws || (relu && !y) || rows_global < rows) return ALIGNQ_EINVAL;
  if ((relu && !aligned16(y)) || (g_residual && !aligned16(g_residual))) return ALIGNQ_EALIGN;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  float* coef = reinterpret_cast<float*>(ws + (size_t)3 * BN_SLOTS * C * 2);
  bnq_sync_coef_kernel<<<(C + 255) / 256, 256, 0, s>>>(sums, (double)rows_global, C, coef);
  ALIGNQ_LAUNCH_CHECK();
  const BnLaunch L = bn_launch(rows, C);
  { cudaError_t pe_ = launch_pdl(bnq_bwd_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, y, gy, nullptr, rows, C, gamma, beta, save_mean, save_invstd, coef, 1,
                                                        make_bnq(a_bit, act_range, variant, relu), gx, g_residual); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

// ---- data-parallel SyncBN with the exchange INSIDE the kernels: peer-mapped buffers over NVLink, no collective call ----
extern "C" size_t alignq_bn_act_peer_bytes(void) { return PEER_BYTES; }

static int peer_check(const void* const* peer_bufs, uint32_t* peer_seq, int rank, int world, int C, int64_t rows, int64_t rows_global) {
  if (!peer_bufs || !peer_seq || world < 2 || world > PEER_MAX || rank < 0 || rank >= world || C > PEER_MAXC || rows_global < rows)
    return ALIGNQ_EINVAL;
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_fwd_peer(const float* x, int64_t rows, int64_t rows_global, int C, const float* gamma,
                                      const float* beta, float* running_mean, float* running_var, float momentum,
                                      float bn_eps, int a_bit, float act_range, int variant, int relu, const float* residual,
                                      float* y, float* save_mean, float* save_invstd, double* ws, uint32_t* counter,
                                      int64_t* num_batches_tracked, const void* const* peer_bufs, uint32_t* peer_seq,
                                      int rank, int world, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, y, residual);
  if (rc) return rc;
  if (!x || !y || !save_mean || !save_invstd || !ws || !counter) return ALIGNQ_EINVAL;
  rc = peer_check(peer_bufs, peer_seq, rank, world, C, rows, rows_global);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const PeerCtx pc{const_cast<void* const*>(peer_bufs), peer_seq, rank, world, (double)rows_global};
  const BnLaunch L = bn_launch(rows, C);
  { cudaError_t pe_ = launch_pdl(bnq_stats_kernel, dim3(L.grid), dim3(L.threads), L.smem, s, x, rows, C, running_mean, running_var, momentum, bn_eps, save_mean,
                                                     save_invstd, ws, counter,
                                                     reinterpret_cast<long long*>(num_batches_tracked), nullptr, pc); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  { cudaError_t pe_ = launch_pdl(bnq_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, rows, C, gamma, beta, save_mean, save_invstd,
                                                    make_bnq(a_bit, act_range, variant, relu), residual, y); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_bwd_peer_sum(const float* x, const float* y, const float* gy, const float* gy2, int64_t rows,
                                          int64_t rows_global, int C, const float* gamma, const float* beta,
                                          const float* save_mean, const float* save_invstd, int a_bit, float act_range,
                                          int variant, int relu, float* gx, float* g_residual, float* ggamma, float* gbeta,
                                          double* ws, uint32_t* counter, const void* const* peer_bufs, uint32_t* peer_seq,
                                          int rank, int world, alignq_stream_t stream) {
  int rc = bn_check(rows, C, a_bit, variant, x, gy, gx);
  if (rc) return rc;
  if (!x || !gy || !gx || !save_mean || !save_invstd || !ws || !counter || (relu && !y)) return ALIGNQ_EINVAL;
  if ((relu && !aligned16(y)) || (g_residual && !aligned16(g_residual)) || (gy2 && !aligned16(gy2))) return ALIGNQ_EALIGN;
  rc = peer_check(peer_bufs, peer_seq, rank, world, C, rows, rows_global);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const PeerCtx pc{const_cast<void* const*>(peer_bufs), peer_seq, rank, world, (double)rows_global};
  const BnLaunch L = bn_launch(rows, C);
  BnQ q = make_bnq(a_bit, act_range, variant, relu);
  float* coef = reinterpret_cast<float*>(ws + (size_t)3 * BN_SLOTS * C * 2);
  if (bn_use_fused(rows, C, (int64_t)6 << 20)) {             // one cooperative launch: reduce -> exchange -> apply
    const int grid = bn_coop_grid(bnq_bwd_fused_kernel, L);
    if (grid > 0) {
      uint32_t* epoch = counter + 1;
      int training = 1;
      PeerCtx pcv = pc;
      void* args[] = {(void*)&x, (void*)&y, (void*)&gy, (void*)&gy2, (void*)&rows, (void*)&C, (void*)&gamma, (void*)&beta, (void*)&save_mean,
                      (void*)&save_invstd, (void*)&training, (void*)&q, (void*)&gx, (void*)&g_residual, (void*)&ggamma,
                      (void*)&gbeta, (void*)&ws, (void*)&epoch, (void*)&pcv, (void*)&coef};
      if (cudaLaunchCooperativeKernel((void*)bnq_bwd_fused_kernel, dim3(grid), dim3(L.threads), args, L.smem, s) == cudaSuccess) {
        ALIGNQ_LAUNCH_CHECK();
        return ALIGNQ_OK;
      }
      (void)cudaGetLastError();
    }
  }
  { cudaError_t pe_ = launch_pdl(bnq_bwd_reduce_kernel, dim3(L.grid), dim3(L.threads), L.smem, s, x, y, gy, gy2, rows, C, gamma, beta, save_mean, save_invstd, q, ggamma,
                                                          gbeta, coef, ws, counter, nullptr, pc); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  { cudaError_t pe_ = launch_pdl(bnq_bwd_apply_kernel, dim3(L.grid * 2), dim3(L.threads), 0, s, x, y, gy, gy2, rows, C, gamma, beta, save_mean, save_invstd, coef, 1, q,
                                                        gx, g_residual); if (pe_ != cudaSuccess) return (int)pe_; }
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_bn_act_bwd_peer(const float* x, const float* y, const float* gy, int64_t rows, int64_t rows_global, int C,
                                      const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                                      int a_bit, float act_range, int variant, int relu, float* gx, float* g_residual,
                                      float* ggamma, float* gbeta, double* ws, uint32_t* counter, const void* const* peer_bufs,
                                      uint32_t* peer_seq, int rank, int world, alignq_stream_t stream) {
  return alignq_bn_act_bwd_peer_sum(x, y, gy, nullptr, rows, rows_global, C, gamma, beta, save_mean, save_invstd, a_bit, act_range,
                                    variant, relu, gx, g_residual, ggamma, gbeta, ws, counter, peer_bufs, peer_seq, rank, world,
                                    stream);
}
