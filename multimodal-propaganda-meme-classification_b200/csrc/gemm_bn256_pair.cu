// CTA-pair (cta_group::2) instantiations of the tcgen05 GEMM: 256 x 256 output tile per cluster of two CTAs.
#include "gemm_kernel.cuh"

namespace b200 {
int launch_gemm_bn256_pair(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& to2,
                           const GemmParams& p, int grid, cudaStream_t stream) {
  return launch_gemm_pair<256>(ta, tb, to, to2, p, grid, stream);
}
}  // namespace b200
