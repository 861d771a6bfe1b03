// tcgen05 Gram modes -- placeholder until the tensor-core kernels land (they return EINVAL so a
// caller can never silently get a different numerics mode than it asked for).
#include "common.cuh"
#include "gram_common.cuh"
#include "../../include/alignq_b200.h"

namespace alignq {
int gram_tc_corr(const float*, int, int64_t, float, float*, void*, size_t, int, cudaStream_t) { return ALIGNQ_EINVAL; }
int gram_tc_fused_fwd(const float*, int, int64_t, ActQ, float, float*, float*, void*, size_t, int, cudaStream_t) {
  return ALIGNQ_EINVAL;
}
}  // namespace alignq
