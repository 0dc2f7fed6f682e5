// rn_wgrad.cu -- weight gradient of a stride-1 SAME 2-D convolution on the 5th-gen tensor cores (backward pass, stage 2;
// the training step of RenderNet_Shader.py:159-167 differentiates every conv filter).
//
//   dW[tap][ci][co] = sum over pixels p of  x[p + off(tap)][ci] * g[p][co]              (x = layer input, g = dL/d(conv output))
//
// A GEMM per tap with M = ci, N = co and K = pixels.  Both operands are channel-last ([pixel][channel]), i.e. M/N-contiguous:
// tcgen05.mma reads them as MN-MAJOR operands straight from what TMA delivers -- a box {64 channels, 64 pixels} lands as 64
// rows of 128 bytes in the 128-byte swizzle, which is the canonical MN-major layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte
// units: SBO = 1024 B between 8-pixel groups, LBO = 8192 B between the 64-channel boxes of a tile.  No transposed copy of the
// activations is ever made.  The tap offset is a shifted TMA coordinate for x (zero fill outside the image = SAME padding).
//
// One CTA tile = (tap, 128 input channels, BN output channels), fp32 accumulator in TMEM; the K loop runs over pixel blocks of a
// slice of the batch (split-K over CTAs when there are fewer tiles than SMs; partial results are combined with fp32 atomics
// into a zero-initialised dW).  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.  Exact mode (fmt 2): three
// passes x_lo.g_hi, x_hi.g_lo, x_hi.g_hi into the same accumulator, corrections first.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/rendernet_b200.h"
#include "rn_igemm.cuh"
#include "rn_ptx.cuh"

namespace rn {
extern std::atomic<long long> g_launch_count;

struct alignas(64) WgradParams {
  CUtensorMap tmX, tmX2;        // activations {Cin, W, H, B}, box {64, PX, PY, 1}; tmX2 = LO plane (fmt 2)
  CUtensorMap tmG, tmG2;        // output gradient {Cout, W, H, B}, same box
  float* dw;                    // [ntaps][Cin][Cout] fp32, accumulated with atomics
  int Cin, Cout, W, H, B;
  int PX, PY;                   // pixel block = PX x PY = 64 pixels
  int ntaps;
  int8_t tap[16][2];            // (dx, dy) of each filter tap
  int m_tiles, n_tiles;         // Cin / 128, Cout / BN
  int ksplit;                   // CTAs sharing one output tile (each takes a slice of the batch)
  int stages;
  int split;                    // fmt 2
};

constexpr int kWgBM = 128;                 // input channels per tile (2 boxes of 64)
constexpr int kBoxBytes = 64 * 128;        // one TMA box: 64 pixels x 64 channels x 2 B

__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);            // start address
  d |= static_cast<uint64_t>(kBoxBytes >> 4) << 16;                   // LBO: next 64-channel box
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                        // SBO: next group of 8 pixels
  d |= static_cast<uint64_t>(1) << 46;                                // descriptor version (Blackwell)
  d |= 2ull << 61;                                                    // SWIZZLE_128B
  return d;
}

template <int BN>
__global__ void __launch_bounds__(192, 1) wgrad2d_kernel(const __grid_constant__ WgradParams p) {
  constexpr int NBOX_B = BN / 64;
  constexpr int kStageBytes = (2 + NBOX_B) * kBoxBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * kStageBytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* acc_bar = empty_bar + p.stages;          // accumulator complete
  uint64_t* free_bar = acc_bar + 1;                  // accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(free_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmG);
    if (p.split) { tma_prefetch_desc(&p.tmX2); tma_prefetch_desc(&p.tmG2); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_bar, 1);
    mbar_init(free_bar, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int bx = p.W / p.PX, by = p.H / p.PY;
  const int kb_per_img = bx * by;
  const int total_work = p.ntaps * p.m_tiles * p.n_tiles * p.ksplit;
  const int passes = p.split ? 3 : 1;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int wk = blockIdx.x; wk < total_work; wk += gridDim.x) {
        const int ks = wk % p.ksplit; int t = wk / p.ksplit;
        const int ni = t % p.n_tiles; t /= p.n_tiles;
        const int mi = t % p.m_tiles; const int tap = t / p.m_tiles;
        const int b0 = (p.B * ks) / p.ksplit, b1 = (p.B * (ks + 1)) / p.ksplit;
        const int dx = p.tap[tap][0], dy = p.tap[tap][1];
        for (int pass = 0; pass < passes; ++pass) {
          // pass order in exact mode: x_lo.g_hi, x_hi.g_lo, x_hi.g_hi
          const CUtensorMap* mx = (p.split && pass == 0) ? &p.tmX2 : &p.tmX;
          const CUtensorMap* mg = (p.split && pass == 1) ? &p.tmG2 : &p.tmG;
          for (int b = b0; b < b1; ++b)
            for (int kb = 0; kb < kb_per_img; ++kb) {
              const int x0 = (kb % bx) * p.PX, y0 = (kb / bx) * p.PY;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], kStageBytes);
              uint8_t* dst = smem + static_cast<size_t>(stage) * kStageBytes;
              tma_load_4d(dst, mx, &full_bar[stage], mi * kWgBM, x0 + dx, y0 + dy, b);
              tma_load_4d(dst + kBoxBytes, mx, &full_bar[stage], mi * kWgBM + 64, x0 + dx, y0 + dy, b);
#pragma unroll
              for (int j = 0; j < NBOX_B; ++j)
                tma_load_4d(dst + (2 + j) * kBoxBytes, mg, &full_bar[stage], ni * BN + 64 * j, x0, y0, b);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // M = 128, N = BN, both operands MN-major (bits 15, 16), fp16 in, fp32 accumulate
      const uint32_t idesc = make_idesc_f16(kWgBM, BN, 0) | (1u << 15) | (1u << 16);
      int stage = 0; uint32_t phase = 0, free_phase = 0;
      bool first_tile = true;
      for (int wk = blockIdx.x; wk < total_work; wk += gridDim.x) {
        const int ks = wk % p.ksplit;
        const int b0 = (p.B * ks) / p.ksplit, b1 = (p.B * (ks + 1)) / p.ksplit;
        const int nkb = (b1 - b0) * kb_per_img * passes;
        if (!first_tile) { mbar_wait(free_bar, free_phase); free_phase ^= 1; }   // epilogue drained the previous tile
        first_tile = false;
        tc_fence_after();
        uint32_t accum = 0;
        for (int it = 0; it < nkb; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + static_cast<size_t>(stage) * kStageBytes);
          const uint64_t da = make_smem_desc_mn(a0), db = make_smem_desc_mn(a0 + 2 * kBoxBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {           // 64 pixels = 4 x (K = 16); 16 pixel rows = 2048 B further on
            umma_f16<1>(tmem_base, da + static_cast<uint64_t>(k * (2048 >> 4)), db + static_cast<uint64_t>(k * (2048 >> 4)),
                        idesc, accum);
            accum = 1;
          }
          umma_commit<1>(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit<1>(acc_bar);
      }
    }
  } else {
    const int quad = warp & 3;                   // TMEM lane quadrant
    const int m = quad * 32 + lane;              // accumulator row = input channel within the tile
    uint32_t acc_phase = 0;
    for (int wk = blockIdx.x; wk < total_work; wk += gridDim.x) {
      int t = wk / p.ksplit;
      const int ni = t % p.n_tiles; t /= p.n_tiles;
      const int mi = t % p.m_tiles; const int tap = t / p.m_tiles;
      mbar_wait(acc_bar, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      float* orow = p.dw + (static_cast<size_t>(tap) * p.Cin + (mi * kWgBM + m)) * p.Cout + ni * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c0), r);
        tmem_ld_wait();
        if (p.ksplit == 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(orow + c0 + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                    __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(orow + c0 + i, __uint_as_float(r[i]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(free_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, BN < 32 ? 32 : BN);
}

template <int BN>
static cudaError_t launch_wgrad(const WgradParams& p, int grid, size_t smem, cudaStream_t st) {
  static std::atomic<bool> attr_set[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad2d_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return e;
    attr_set[dev].store(true, std::memory_order_release);
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  wgrad2d_kernel<BN><<<grid, 192, smem, st>>>(p);
  return cudaGetLastError();
}

// sum over pixels of a 16-bit channel-last tensor: db[c] = sum_p g[p][c] (bias gradient), fp32 atomics into a zeroed vector
__global__ void bias_grad_kernel(const uint16_t* __restrict__ g, float* __restrict__ db, long long npix, int C, int fmt,
                                 long long plane) {
  // one thread per (pixel chunk, channel): blockDim.x = channels handled per block row, grid-stride over pixel chunks
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (long long px = blockIdx.x; px < npix; px += gridDim.x) {
    const long long i = px * C + c;
    float v = __half2float(*reinterpret_cast<const __half*>(g + i));
    if (fmt == 2) v += __half2float(*reinterpret_cast<const __half*>(g + i + plane));
    acc += v;
  }
  atomicAdd(db + c, acc);
}

}  // namespace rn

using namespace rn;

extern "C" int rn_conv2d_weight_grad(const void* x, const void* g, float* dw, int B, int H, int W, int Cin, int Cout, int kh,
                                     int kw, int fmt, void* stream) {
  if (!x || !g || !dw || B < 1 || H < 1 || W < 1 || kh < 1 || kw < 1 || kh * kw > 16) return -1;
  if (fmt != 0 && fmt != 2) return -2;                                   // fp16 or fp16 hi/lo pairs
  if (Cin % 128 != 0 || Cout % 128 != 0) return -3;
  int PX = 64, PY = 1;
  while (PX > W) { PX >>= 1; PY <<= 1; }
  if (W % PX != 0 || H % PY != 0 || PX * PY != 64) return -4;
  const int BN = (Cout % 256 == 0) ? 256 : 128;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) return -8;
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.dw = dw; p.Cin = Cin; p.Cout = Cout; p.W = W; p.H = H; p.B = B; p.PX = PX; p.PY = PY; p.ntaps = kh * kw;
  const int pby = (kh - 1) / 2, pbx = (kw - 1) / 2;                      // TF SAME, stride 1: pad-before = (k-1)//2
  for (int ky = 0; ky < kh; ++ky)
    for (int kx = 0; kx < kw; ++kx) { p.tap[ky * kw + kx][0] = static_cast<int8_t>(kx - pbx); p.tap[ky * kw + kx][1] = static_cast<int8_t>(ky - pby); }
  p.m_tiles = Cin / kWgBM; p.n_tiles = Cout / BN; p.split = fmt == 2 ? 1 : 0;
  const int tiles = p.ntaps * p.m_tiles * p.n_tiles;
  const int sms = num_sms();
  int ksplit = 1;
  while (tiles * ksplit < sms && ksplit * 2 <= B) ksplit *= 2;          // fewer tiles than SMs: split the batch over CTAs
  p.ksplit = ksplit;
  const int stage_bytes = (2 + BN / 64) * kBoxBytes;
  p.stages = (232448 - 1024 - 256) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) return -10;
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 1024 + 256;
  const cuuint32_t ones[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(PX), static_cast<cuuint32_t>(PY), 1};
  auto encode = [&](CUtensorMap* m, const void* base, int C) -> CUresult {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(C) * 2 * W, static_cast<cuuint64_t>(C) * 2 * W * H};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  const long long xplane = static_cast<long long>(B) * H * W * Cin, gplane = static_cast<long long>(B) * H * W * Cout;
  CUresult r = encode(&p.tmX, x, Cin);
  if (r == CUDA_SUCCESS) r = encode(&p.tmG, g, Cout);
  if (r == CUDA_SUCCESS && p.split) r = encode(&p.tmX2, static_cast<const uint16_t*>(x) + xplane, Cin);
  if (r == CUDA_SUCCESS && p.split) r = encode(&p.tmG2, static_cast<const uint16_t*>(g) + gplane, Cout);
  if (r != CUDA_SUCCESS) return 1000 + static_cast<int>(r);
  if (ksplit > 1) {
    cudaError_t e = cudaMemsetAsync(dw, 0, static_cast<size_t>(p.ntaps) * Cin * Cout * sizeof(float), st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int grid = tiles * ksplit;
  if (grid > sms) grid = sms;
  const cudaError_t e = (BN == 256) ? launch_wgrad<256>(p, grid, smem, st) : launch_wgrad<128>(p, grid, smem, st);
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

extern "C" int rn_bias_grad_16(const void* g, float* db, long long npix, int C, int fmt, void* stream) {
  if (!g || !db || npix < 1 || C < 1 || (fmt != 0 && fmt != 2)) return -1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(db, 0, static_cast<size_t>(C) * sizeof(float), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int bx = C < 128 ? C : 128;
  dim3 grid(static_cast<unsigned>(npix < 1024 ? npix : 1024), static_cast<unsigned>((C + bx - 1) / bx));
  bias_grad_kernel<<<grid, bx, 0, st>>>(static_cast<const uint16_t*>(g), db, npix, C, fmt, npix * C);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return static_cast<int>(cudaGetLastError());
}
