// ABI bookkeeping: version and error strings.
#include "common.cuh"
#include "../../include/alignq_b200.h"

#include <atomic>

static std::atomic<unsigned long long> g_launches{0};
extern "C" void alignq_count_launch_(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" uint64_t alignq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int alignq_abi_version(void) { return ALIGNQ_ABI_VERSION; }

extern "C" const char* alignq_error_string(int code) {
  switch (code) {
    case ALIGNQ_OK: return "ok";
    case ALIGNQ_EINVAL: return "alignq: invalid argument (null pointer, bad bit-width/variant/shape, or identity case that is the caller's job)";
    case ALIGNQ_EALIGN: return "alignq: pointer alignment requirement not met";
    case ALIGNQ_ERANGE: return "alignq: size outside the supported range (codes overflow int16, batch too large for the kernel)";
    case ALIGNQ_ENOSPACE: return "alignq: workspace too small (see alignq_gram_ws_bytes)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "alignq: unknown error";
}
