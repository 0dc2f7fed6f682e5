// Instantiations of igemm_kernel for the narrow N tiles (64 / 32 / 16 columns: thin decoder layers), 1 or 2 M sub-tiles.
#include "rn_igemm_kernel.cuh"

namespace rn {
// two epilogue warp groups, one per M sub-tile (MS = 2): the thin 512^2 decoder layers are pure epilogue / data movement
cudaError_t launch_small_eg2(int BN, const IgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (BN == 64) return launch_ms<64, 1, 1, 2, 2>(p, grid, smem, stream);
  if (BN == 32) return launch_ms<32, 1, 1, 2, 2>(p, grid, smem, stream);
  return launch_ms<16, 1, 1, 2, 2>(p, grid, smem, stream);
}

cudaError_t launch_small(int BN, const IgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (BN == 64) return launch_bn<64, 1, 1>(p, grid, smem, stream);
  if (BN == 32) return launch_bn<32, 1, 1>(p, grid, smem, stream);
  return launch_bn<16, 1, 1>(p, grid, smem, stream);
}
}  // namespace rn
