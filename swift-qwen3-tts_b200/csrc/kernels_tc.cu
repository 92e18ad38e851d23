// tcgen05 / TMEM / TMA multi-tap GEMM (placeholder until the kernel lands: reports "unsupported"
// so the engine routes 16-bit GEMMs through the CUDA-core kernel).
#include "kernels.cuh"
namespace q3 {
bool tc_supported(const ConvGemmParams&, int) { return false; }
cudaError_t launch_conv_gemm_tc(const ConvGemmParams&, const BatchGeom&, int, int, cudaStream_t) {
  return cudaErrorNotSupported;
}
}  // namespace q3
