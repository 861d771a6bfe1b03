// ABI bookkeeping: version and error strings.
#include "common.cuh"
#include "../../include/alignq_b200.h"

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
