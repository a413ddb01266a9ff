// tcgen05 GEMM instantiations for the 128-column tile (all epilogue modes); see gemm_kernel.cuh.
#include "gemm_kernel.cuh"

namespace b200 {
int launch_gemm_bn128(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                     const GemmParams& p, int grid, cudaStream_t stream) {
  return launch_gemm_bn<128>(ta, tb, to, to2, p, grid, stream);
}
}  // namespace b200
