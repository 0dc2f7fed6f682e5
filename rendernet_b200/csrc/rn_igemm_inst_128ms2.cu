// Instantiations of igemm_kernel for the 128-column N tile with two M sub-tiles (accumulators) per CTA tile.
#include "rn_igemm_kernel.cuh"

namespace rn {
cudaError_t launch_bn128_ms2(int CL, int CG, const IgemmParams& p, int grid, size_t smem, cudaStream_t stream) {
  if (CG == 2) return launch_ms<128, 2, 2, 2>(p, grid, smem, stream);
  if (CL == 4) return launch_ms<128, 4, 1, 2>(p, grid, smem, stream);
  if (CL == 2) return launch_ms<128, 2, 1, 2>(p, grid, smem, stream);
  return launch_ms<128, 1, 1, 2>(p, grid, smem, stream);
}
}  // namespace rn
