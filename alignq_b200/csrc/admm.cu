// ADMM correlation-preservation term: split-K Gram reduction, fused loss + dL/dD, and the Z/U
// dual update without a host sync.
//
// Replaces ADMM.forward (utils/admm.py:24-33) and the alterD/gamma branches of ADMM_OPT.step
// (utils/optimizer.py:97-124; the reference evaluates `if torch.norm(V, 2) > mu/rho` on the host).
// All matrices are [B, B] with B <= dim (<= 256 KB): latency-bound O(B^2) work, but one SM's L2
// bandwidth (~120 GB/s) would make a single CTA take ~15 us.  Each ADMM module is therefore handled
// by one thread-block CLUSTER of 8 CTAs (8 SMs): every CTA reduces its slice, the partial sums are
// exchanged through distributed shared memory (cluster.map_shared_rank + cluster.sync), and every CTA
// then applies the second pass to its slice.  Modules are batched through gridDim.x (= 8 * nmod).
#include <cooperative_groups.h>

#include "common.cuh"
#include "gram_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {

namespace cg = cooperative_groups;

constexpr int AT = 512;           // threads per CTA
constexpr int ACL = 8;            // CTAs per cluster (one cluster per ADMM module)

template <typename T>
__device__ __forceinline__ T block_sum1(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  v = (lane < nw) ? scratch[lane] : T(0);
  return warp_sum(v);
}

// Cluster-wide sum of NV doubles: block reduce, deposit into rank 0's shared array through DSMEM,
// cluster.sync, everyone reads the totals back from rank 0.  part: __shared__ double[ACL * NV] per CTA.
template <int NV>
__device__ __forceinline__ void cluster_sum(double (&v)[NV], double* part, double* scratch) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = block_sum1(v[i], scratch);
  cluster.sync();                        // every CTA of the cluster is resident before remote smem is touched
  double* root = cluster.map_shared_rank(part, 0);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) root[rank * NV + i] = v[i];
  }
  cluster.sync();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double t = 0.0;
    for (int r = 0; r < ACL; ++r) t += root[r * NV + i];
    v[i] = t;
  }
  cluster.sync();                        // rank 0's array may be reused after everyone has read it
}

// G = sum_slab P / F   (fused: D = G_t - G_x as two separately rounded fp32 results, QB:118-122)
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const float* __restrict__ partials, int nslabs, int B, float invF, int fused,
                   float* __restrict__ G, float* __restrict__ D) {
  const size_t bb = (size_t)B * B;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= bb) return;
  float sx = 0.f, st = 0.f;
  for (int s = 0; s < nslabs; ++s) sx += partials[(size_t)s * bb + e];
  const float gx = __fmul_rn(sx, invF);
  if (G) G[e] = gx;
  if (fused) {
    const float* pt = partials + (size_t)nslabs * bb;
    for (int s = 0; s < nslabs; ++s) st += pt[(size_t)s * bb + e];
    D[e] = (fused == 2) ? __fmul_rn(st, invF) : __fsub_rn(__fmul_rn(st, invF), gx);     // 2: the t sum itself
  }
}

// D = sum_t / F - sum_x / F (two separately rounded Grams, QB:118-122) from the un-normalised column sums of the
// data-parallel feature-sharded path, after their all-reduce; module m: sums + m*2*B*B, invF[m], D + m*B*B.
__global__ void __launch_bounds__(256)
gram_sums_to_d_kernel(const float* __restrict__ sums, const float* __restrict__ invF, int B, float* __restrict__ D) {
  const size_t bb = (size_t)B * B;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= bb) return;
  const int m = blockIdx.y;
  const float f = invF[m];
  const float* sm = sums + (size_t)m * 2 * bb;
  D[(size_t)m * bb + e] = __fsub_rn(__fmul_rn(sm[bb + e], f), __fmul_rn(sm[e], f));
}

__global__ void __launch_bounds__(256)
wsym_kernel(const float* __restrict__ dLdD, int B, int Bp, float* __restrict__ W) {
  pdl_trigger();                      // the Gram backward launches early and waits for this grid (common.cuh)
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Bp * Bp) return;
  const int i = e / Bp, j = e - i * Bp;
  W[e] = (i < B && j < B) ? dLdD[(size_t)i * B + j] + dLdD[(size_t)j * B + i] : 0.f;
}

__global__ void __cluster_dims__(ACL, 1, 1) __launch_bounds__(AT)
admm_loss_kernel(const float* __restrict__ D, int B, const float* __restrict__ Z, const float* __restrict__ U, int dim,
                 float mu, float rho, const float* __restrict__ gloss, int gloss_per_module,
                 float* __restrict__ loss, float* __restrict__ dLdD, float* __restrict__ dLdZ, float* __restrict__ dLdU) {
  __shared__ double scratch[32];
  __shared__ double part[ACL * 3];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int m = blockIdx.x / ACL;
  const float* Dm = D + (size_t)m * B * B;
  const float* Zm = Z + (size_t)m * dim * dim;
  const float* Um = U + (size_t)m * dim * dim;
  const int bb = B * B;
  double s[3] = {0.0, 0.0, 0.0};                            // sum |Z|, sum R^2, sum U |R| over this CTA's slice
  for (int e = rank * AT + threadIdx.x; e < bb; e += ACL * AT) {
    const int i = e / B, j = e - i * B;
    const float z = Zm[(size_t)i * dim + j], u = Um[(size_t)i * dim + j];
    const float r = __fsub_rn(Dm[e], z);
    s[0] += (double)fabsf(z);
    s[1] += (double)__fmul_rn(r, r);
    s[2] += (double)__fmul_rn(u, fabsf(r));
  }
  cluster_sum<3>(s, part, scratch);
  const float inv_bb = 1.0f / (float)bb;
  const float mean_r2 = (float)(s[1] / bb);
  const float rms = sqrtf(mean_r2);                               // mean(...) ** 0.5
  if (rank == 0 && threadIdx.x == 0 && loss) {
    const float reg = mu * (float)(s[0] / bb);
    const float con = (rho / 2.0f) * rms;
    loss[m] = (reg + con) + (float)(s[2] / bb);
  }
  if (dLdD || dLdZ || dLdU) {
    // d/dD [rho/2 sqrt(mean R^2)] = rho/2 * R / (B^2 rms);  d/dD mean(U |R|) = U sign(R) / B^2
    // d/dZ = mu sign(Z) / B^2 - d/dD;  d/dU = |R| / B^2;  all scaled by the upstream gradient gl.
    const float gl = gloss ? __ldg(gloss + (gloss_per_module ? m : 0)) : 1.0f;
    const float k1 = (rho / 2.0f) / rms * inv_bb;                 // rms == 0 -> inf, and 0 * inf = NaN as autograd gives
    float* GD = dLdD ? dLdD + (size_t)m * bb : nullptr;
    float* GZ = dLdZ ? dLdZ + (size_t)m * dim * dim : nullptr;
    float* GU = dLdU ? dLdU + (size_t)m * dim * dim : nullptr;
    for (int e = rank * AT + threadIdx.x; e < dim * dim; e += ACL * AT) {
      const int i = e / dim, j = e - i * dim;
      if (i < B && j < B) {
        const float z = Zm[e], u = Um[e];
        const float r = __fsub_rn(Dm[(size_t)i * B + j], z);
        const float sg = (r > 0.f) ? 1.f : ((r < 0.f) ? -1.f : 0.f);
        const float gd = k1 * r + u * sg * inv_bb;
        if (GD) GD[(size_t)i * B + j] = gl * gd;
        if (GZ) GZ[e] = gl * (mu * ((z > 0.f) ? 1.f : ((z < 0.f) ? -1.f : 0.f)) * inv_bb - gd);
        if (GU) GU[e] = gl * fabsf(r) * inv_bb;
      } else {
        if (GZ) GZ[e] = 0.f;
        if (GU) GU[e] = 0.f;
      }
    }
  }
}

__global__ void __cluster_dims__(ACL, 1, 1) __launch_bounds__(AT)
admm_zu_kernel(float* __restrict__ Z, float* __restrict__ U, const float* __restrict__ D, int B, int dim,
               float mu_over_rho, float inv_rho, float rho) {
  __shared__ double scratch[32];
  __shared__ double part[ACL];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int m = blockIdx.x / ACL;
  float* Zm = Z + (size_t)m * dim * dim;
  float* Um = U + (size_t)m * dim * dim;
  const float* Dm = D + (size_t)m * B * B;
  const int dd = dim * dim;
  double ss[1] = {0.0};
  for (int e = rank * AT + threadIdx.x; e < dd; e += ACL * AT) {
    const int i = e / dim, j = e - i * dim;
    const float d = (i < B && j < B) ? Dm[(size_t)i * B + j] : 0.f;
    const float v = __fadd_rn(d, __fmul_rn(inv_rho, Um[e]));       // V = D_ + 1/rho * gamma
    ss[0] += (double)v * (double)v;
  }
  cluster_sum<1>(ss, part, scratch);
  const float nv = (float)sqrt(ss[0]);                            // torch.norm(V, 2): Frobenius
  const bool shrink = nv > mu_over_rho;
  const float coef = __fsub_rn(1.0f, __fmul_rn(__frcp_rn(nv), mu_over_rho));   // 1 - mu/rho / ||V||
  for (int e = rank * AT + threadIdx.x; e < dd; e += ACL * AT) {
    const int i = e / dim, j = e - i * dim;
    const float d = (i < B && j < B) ? Dm[(size_t)i * B + j] : 0.f;
    const float u = Um[e];
    const float v = __fadd_rn(d, __fmul_rn(inv_rho, u));
    const float z = shrink ? __fmul_rn(coef, v) : 0.f;
    Zm[e] = z;
    Um[e] = __fadd_rn(u, __fmul_rn(rho, __fsub_rn(d, z)));         // gamma + rho * (D_ - alterD)
  }
}

int launch_gram_reduce(const float* partials, int nslabs, int B, int64_t F, int fused, float* G, float* D, cudaStream_t s) {
  const int bb = B * B;
  gram_reduce_kernel<<<(bb + 255) / 256, 256, 0, s>>>(partials, nslabs, B, 1.0f / (float)F, fused, G, D);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

int launch_gram_reduce_raw(const float* partials, int nslabs, int B, float* Gx, float* Gt, cudaStream_t s) {
  const int bb = B * B;
  gram_reduce_kernel<<<(bb + 255) / 256, 256, 0, s>>>(partials, nslabs, B, 1.0f, 2, Gx, Gt);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

int launch_wsym(const float* dLdD, int B, float* Wsym, cudaStream_t s) {
  const int Bp = gram_bp(B);
  wsym_kernel<<<(Bp * Bp + 255) / 256, 256, 0, s>>>(dLdD, B, Bp, Wsym);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

}  // namespace alignq

using namespace alignq;

extern "C" int alignq_admm_loss(const float* D, int B, const float* Z, const float* U, int dim, int nmod, float mu,
                                float rho, const float* gloss, int gloss_per_module, float* loss, float* dLdD,
                                float* dLdZ, float* dLdU, alignq_stream_t stream) {
  if (B < 1 || dim < B || nmod < 0 || dim > 4096) return ALIGNQ_EINVAL;
  if (nmod == 0) return ALIGNQ_OK;
  if (!D || !Z || !U) return ALIGNQ_EINVAL;
  admm_loss_kernel<<<nmod * ACL, AT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(D, B, Z, U, dim, mu, rho, gloss,
                                                                           gloss_per_module, loss, dLdD, dLdZ, dLdU);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_gram_sums_to_d(const float* sums, const float* inv_f, int B, int nmod, float* D, alignq_stream_t stream) {
  if (B < 1 || nmod < 0) return ALIGNQ_EINVAL;
  if (nmod == 0) return ALIGNQ_OK;
  if (!sums || !inv_f || !D) return ALIGNQ_EINVAL;
  gram_sums_to_d_kernel<<<dim3((B * B + 255) / 256, nmod), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sums, inv_f, B, D);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}

extern "C" int alignq_admm_zu_update(float* Z, float* U, const float* D, int B, int dim, int nmod, float mu, float rho,
                                     alignq_stream_t stream) {
  if (B < 1 || dim < B || nmod < 0 || dim > 4096) return ALIGNQ_EINVAL;
  if (nmod == 0) return ALIGNQ_OK;
  if (!D || !Z || !U) return ALIGNQ_EINVAL;
  // python: mu / rho and 1 / rho are evaluated in double, then become fp32 scalars of the tensor ops
  const float mu_over_rho = (float)((double)mu / (double)rho);
  const float inv_rho = (float)(1.0 / (double)rho);
  admm_zu_kernel<<<nmod * ACL, AT, 0, reinterpret_cast<cudaStream_t>(stream)>>>(Z, U, D, B, dim, mu_over_rho, inv_rho, rho);
  ALIGNQ_LAUNCH_CHECK();
  return ALIGNQ_OK;
}
