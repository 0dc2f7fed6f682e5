// igemm_kernel with two epilogue warp groups, 256-column N tile, CTA pairs (projection unit and other short-K layers).
#include "rn_igemm_kernel.cuh"

namespace rn {
cudaError_t launch_eg2_256(const IgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  return launch_ms<256, 2, 2, 1, 2>(p, grid, smem, stream);
}
}  // namespace rn
