// Helpers shared by the small-batch (B <= 32) thread-per-column tcgen05 kernels (gram_tc_small*.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace alignq {
namespace tcsmall {

__device__ __forceinline__ float ld_once(const float* p) {        // read-once stream: do not keep the line in L1
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// predicated streaming store: keeps the unrolled per-row code one basic block (a C++ `if` around the store makes
// the compiler sink the row's arithmetic into a branch, which stops the 32 independent rows from interleaving)
__device__ __forceinline__ void st_if(float* p, float v, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.L1::no_allocate.f32 [%0], %1;\n\t}"
               ::"l"(p), "f"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// 4-byte asynchronous global -> shared copy (src_bytes = 0 zero-fills); each thread later reads only what it copied
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 8 consecutive batch rows of one feature column -> one 16-byte core-matrix row of the bf16 H operand and, when
// SPLIT, of the L operand (L = bf16(v - H): together 16 mantissa bits); two values per cvt.rn.bf16x2.f32
template <bool SPLIT>
__device__ __forceinline__ void store_chunk(uint8_t* dst, int l_offset, const float (&c)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(c[2 * p], c[2 * p + 1]);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    if (SPLIT) {
      const float h0 = __uint_as_float(h[p] << 16), h1 = __uint_as_float(h[p] & 0xFFFF0000u);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(c[2 * p] - h0, c[2 * p + 1] - h1);
      l[p] = *reinterpret_cast<const uint32_t*>(&ll);
    }
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (SPLIT) *reinterpret_cast<uint4*>(dst + l_offset) = make_uint4(l[0], l[1], l[2], l[3]);
}

// The same for NS = 1, 2 or 3 bf16 terms per value: tiles at dst, dst + stride, dst + 2 stride hold H, M, L with
// v = H + M (+ L) to 16 (24) mantissa bits (M = bf16(v - H), L = bf16(v - H - M); both subtractions are exact in fp32).
template <int NS>
__device__ __forceinline__ void store_chunk_n(uint8_t* dst, int stride, const float (&c)[8]) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(c[2 * p], c[2 * p + 1]);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    if (NS >= 2) {
      const float r0 = c[2 * p] - __uint_as_float(h[p] << 16), r1 = c[2 * p + 1] - __uint_as_float(h[p] & 0xFFFF0000u);
      const __nv_bfloat162 mm = __floats2bfloat162_rn(r0, r1);
      m[p] = *reinterpret_cast<const uint32_t*>(&mm);
      if (NS == 3) {
        const __nv_bfloat162 ll = __floats2bfloat162_rn(r0 - __uint_as_float(m[p] << 16), r1 - __uint_as_float(m[p] & 0xFFFF0000u));
        l[p] = *reinterpret_cast<const uint32_t*>(&ll);
      }
    }
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (NS >= 2) *reinterpret_cast<uint4*>(dst + stride) = make_uint4(m[0], m[1], m[2], m[3]);
  if (NS == 3) *reinterpret_cast<uint4*>(dst + 2 * stride) = make_uint4(l[0], l[1], l[2], l[3]);
}

// the same split without the stores: t[0..NS-1] = the 16-byte H (, M (, L)) chunks of 8 values
template <int NS>
__device__ __forceinline__ void split_chunk_n(const float (&c)[8], uint4 (&t)[NS]) {
  uint32_t h[4], m[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(c[2 * p], c[2 * p + 1]);
    h[p] = *reinterpret_cast<const uint32_t*>(&hh);
    if (NS >= 2) {
      const float r0 = c[2 * p] - __uint_as_float(h[p] << 16), r1 = c[2 * p + 1] - __uint_as_float(h[p] & 0xFFFF0000u);
      const __nv_bfloat162 mm = __floats2bfloat162_rn(r0, r1);
      m[p] = *reinterpret_cast<const uint32_t*>(&mm);
      if (NS == 3) {
        const __nv_bfloat162 ll = __floats2bfloat162_rn(r0 - __uint_as_float(m[p] << 16), r1 - __uint_as_float(m[p] & 0xFFFF0000u));
        l[p] = *reinterpret_cast<const uint32_t*>(&ll);
      }
    }
  }
  t[0] = make_uint4(h[0], h[1], h[2], h[3]);
  if (NS >= 2) t[1] = make_uint4(m[0], m[1], m[2], m[3]);
  if (NS == 3) t[NS - 1] = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tcsmall
}  // namespace alignq
