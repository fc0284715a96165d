// TEST INFRASTRUCTURE: csrc/attention.cu hands the long sequences to the tcgen05 kernel of attention_tc.cu when a
// workspace is given; that path cannot run on the CPU emulator and reports HOISDF_E_UNSUPPORTED there (the emulated
// tests call the SIMT kernels: workspace == NULL).
#include "cuda_emu.h"
namespace hoisdf {
int64_t attention_tc_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
int launch_attention_tc(const float*, int64_t, const float*, const float*, int64_t, float*, int64_t, int64_t, int64_t,
                        int64_t, int64_t, int64_t, void*, cudaStream_t, uint16_t*, uint16_t*, float, uint64_t, float*) {
  return HOISDF_E_UNSUPPORTED;
}
int64_t attention_bwd_tc_workspace_bytes(int64_t, int64_t, int64_t, int64_t) { return 0; }
int launch_attention_bwd_tc(const float*, int64_t, const float*, const float*, int64_t, const float*, const float*, int64_t,
                            const float*, float*, float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, float,
                            uint64_t, void*, cudaStream_t) {
  return HOISDF_E_UNSUPPORTED;
}
}  // namespace hoisdf
