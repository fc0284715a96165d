// ABI version and status strings.
#include "common.cuh"

HOISDF_API int hoisdf_abi_version(void) { return HOISDF_ABI_VERSION; }

HOISDF_API const char* hoisdf_status_string(int status) {
  switch (status) {
    case HOISDF_OK: return "ok";
    case HOISDF_E_NULL: return "required pointer is NULL";
    case HOISDF_E_SHAPE: return "size out of the supported range";
    case HOISDF_E_ALIGN: return "pointer or leading dimension not 16-byte aligned";
    case HOISDF_E_UNSUPPORTED: return "unsupported configuration";
    case HOISDF_E_WORKSPACE: return "workspace too small";
    case HOISDF_E_TOO_FEW_POINTS: return "a sample has fewer lattice points inside its bbox than num_points";
    default: break;
  }
  if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
  return "unknown hoisdf status";
}
