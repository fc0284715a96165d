// TEST INFRASTRUCTURE: csrc/linear.cu dispatches to the tcgen05 3xTF32 kernels when w_lo is given; those cannot run on the
// CPU emulator and report HOISDF_E_UNSUPPORTED there (the emulated tests only use the fp32 FMA kernel, w_lo == NULL).
#include "cuda_emu.h"
namespace hoisdf {
int launch_linear_tf32x3(const hoisdf_linear_args*, cudaStream_t) { return HOISDF_E_UNSUPPORTED; }
int launch_linear_tf32x3_2sm(const hoisdf_linear_args*, cudaStream_t) { return HOISDF_E_UNSUPPORTED; }
}  // namespace hoisdf
