// NHWC <-> (channel, pixel) layout changes of the pixel agents' 32-channel maps, through a 32 x 33 shared-memory tile so that
// both the 128-byte NHWC rows and the per-channel pixel runs are read / written as whole lines (the one-element-per-thread
// versions touched a different line per lane on one side: 45 us for 80 MB).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace rlrep {

// out[b * ld + c * P + p] = in[(b * P + p) * 32 + c]      grid (ceil(P / 32), B), block (32, 8)
static __global__ void __launch_bounds__(256) nhwc_to_cp_kernel(const float* __restrict__ in, int P, float* __restrict__ out,
                                                         long long ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y, p0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;
  const float* src = in + (size_t)b * P * 32;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r;
    if (p < P) tile[ty + 8 * r][tx] = src[(size_t)p * 32 + tx];
  }
  __syncthreads();
  float* dst = out + (size_t)b * ld;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = ty + 8 * r, p = p0 + tx;
    if (p < P) dst[(size_t)c * P + p] = tile[tx][c];
  }
}

// out[(b * P + p) * 32 + c] = in[b * ld + c * P + p] (* (mask[(b * P + p) * 32 + c] > 0) when mask != nullptr)
static __global__ void __launch_bounds__(256) cp_to_nhwc_kernel(const float* __restrict__ in, long long ld, int P,
                                                         const float* __restrict__ mask, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.y, p0 = blockIdx.x * 32, tx = threadIdx.x, ty = threadIdx.y;
  const float* src = in + (size_t)b * ld;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = ty + 8 * r, p = p0 + tx;
    if (p < P) tile[c][tx] = src[(size_t)c * P + p];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = p0 + ty + 8 * r;
    if (p < P) {
      const size_t o = ((size_t)b * P + p) * 32 + tx;
      float v = tile[tx][ty + 8 * r];
      if (mask != nullptr) v = mask[o] > 0.f ? v : 0.f;
      out[o] = v;
    }
  }
}

inline void launch_nhwc_to_cp(const float* in, int B, int P, float* out, long long ld, cudaStream_t s) {
  nhwc_to_cp_kernel<<<dim3((P + 31) / 32, B), dim3(32, 8), 0, s>>>(in, P, out, ld);
}
inline void launch_cp_to_nhwc(const float* in, long long ld, int B, int P, const float* mask, float* out, cudaStream_t s) {
  cp_to_nhwc_kernel<<<dim3((P + 31) / 32, B), dim3(32, 8), 0, s>>>(in, ld, P, mask, out);
}

}  // namespace rlrep
