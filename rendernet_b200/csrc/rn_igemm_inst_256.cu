// Instantiations of igemm_kernel for the 256-column N tile (trunk, res3, e_conv5/6, wide transposed convs).
#include "rn_igemm_kernel.cuh"

namespace rn {
cudaError_t launch_bn256(int CL, int CG, const IgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (CG == 2) return launch_ms<256, 2, 2, 1>(p, grid, smem, stream);
  if (CL == 4) return launch_ms<256, 4, 1, 1>(p, grid, smem, stream);
  if (CL == 2) return launch_ms<256, 2, 1, 1>(p, grid, smem, stream);
  return launch_ms<256, 1, 1, 1>(p, grid, smem, stream);
}
}  // namespace rn
