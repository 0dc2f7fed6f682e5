// rn_phong.cuh -- one pixel of the demo's Phong composite (tools/Phong_shading.py:138-228 np_mask / np_phong_shading /
// np_phong_composite, RenderNet_demo.py:54-58), shared by the stand-alone kernel (rn_ops.cu) and the fused epilogue of the
// last up-conv (rn_igemm_kernel.cuh) so that both produce the same bits.
#pragma once
#include <cstdint>

namespace rn {

// (r,g,b) = normal map in [0,1]; l = light direction (not normalised), col = light colour.  Returns the shaded colour.
__device__ __forceinline__ void phong_pixel(float r, float g, float bl, float lx, float ly, float lz, const float* col,
                                            float ambient, float k_diffuse, int white, int with_mask, float* out) {
  // np_phong_shading (:162-200): n = (img-0.5)/|img-0.5| ; diffuse = k_d * max(n.l, 0) * light_col, clipped
  const float nx = r - 0.5f, ny = g - 0.5f, nz = bl - 0.5f;
  const float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);
  const float linv = 1.0f / sqrtf(lx * lx + ly * ly + lz * lz);
  lx *= linv; ly *= linv; lz *= linv;
  const float ndl = fmaxf((nx * lx + ny * ly + nz * lz) * inv, 0.f);
  float mask = 1.f;
  if (with_mask) {  // np_mask (:138-148) / np_mask_white (:150-160)
    const float nrm = white ? sqrtf((1.f - r) * (1.f - r) + (1.f - g) * (1.f - g) + (1.f - bl) * (1.f - bl))
                            : sqrtf(r * r + g * g + bl * bl);
    mask = 1.f / (1.f + expf(-(255.f * nrm - (white ? 80.f : 150.f))));
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float diff = fminf(fmaxf(k_diffuse * ndl * col[c], 0.f), 1.f);
    float v = with_mask ? mask * (ambient + diff) + (1.f - mask) : ambient + diff;
    out[c] = fminf(fmaxf(v, 0.f), 1.f);
  }
}
__device__ __forceinline__ uint8_t phong_u8(float v) { return static_cast<uint8_t>(fminf(fmaxf(255.f * v, 0.f), 255.f)); }

}  // namespace rn
