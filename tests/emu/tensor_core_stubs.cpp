// TEST INFRASTRUCTURE: the tcgen05 GEMM entry points cannot run on the CPU emulator; entry points of emulated files that
// call them (the SDF-decoder chains in csrc/sdf.cu) report HOISDF_E_UNSUPPORTED there.
extern "C" int hoisdf_linear_fwd(const void*, void*) { return -4; }
extern "C" int hoisdf_linear_h3_fwd(const void*, void*) { return -4; }
