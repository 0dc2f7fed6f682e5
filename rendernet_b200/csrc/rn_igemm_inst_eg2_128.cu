// igemm_kernel with two epilogue warp groups, 128-column N tile: CTA pairs (CG = 2) or a multicast cluster of 2 (CG = 1).
#include "rn_igemm_kernel.cuh"

namespace rn {
cudaError_t launch_eg2_128_pair(const IgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (p.ms == 2) return launch_ms<128, 2, 2, 2, 2>(p, grid, smem, stream);
  return launch_ms<128, 2, 2, 1, 2>(p, grid, smem, stream);
}
}  // namespace rn
