// TEST INFRASTRUCTURE: the FP16x3 tcgen05 GEMM entry point cannot run on the CPU emulator (see stubs_sdf.cpp).
extern "C" int hoisdf_linear_h3_fwd(const void*, void*) { return -4; }
