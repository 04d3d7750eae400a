// Implicit "full correlation" of a 32-channel NHWC map with a 3x3 kernel on the tcgen05 GEMM (see conv_implicit.cu):
// the shared core of the encoder's data gradient (conv.cu) and the decoder's stride-1 transposed convolutions
// (deconv.cu).
#pragma once
#include "agent.cuh"

namespace rlrep {

// Scratch for one full correlation at a time: the zero-padded input grid, the output grid and the repacked weights.
struct FullCorrScratch {
  float *padded = nullptr, *out_grid = nullptr, *w_flip = nullptr;
  void want(DeviceArena& a, int batch, int max_hi) {
    const size_t rows = (size_t)batch * (max_hi + 4) * (max_hi + 4);
    a.want(&padded, rows * 32);
    a.want(&out_grid, rows * 32);
    a.want(&w_flip, 32 * 288);
  }
};

enum FullCorrWeights : int {
  FC_CONV_DGRAD = 0,   // W is a conv weight [32 (co), (ky, kx, ci)]: out channel n = ci, contraction over co
  FC_DECONV_FWD = 1,   // W is a transposed-conv weight [(ky, kx, co), ci]: out channel n = co, contraction over ci
};

// out[b, oy, ox, n] = act(bias[n] + sum_{ky, kx, c} in[b, oy - ky, ox - kx, c] * Wt[n, (ky, kx), c]) * (mask > 0),
// in [B, Hi, Hi, 32] (zero outside), out / mask [B, Hi + 2, Hi + 2, 32]; bias and mask may be null.
void full_correlation_3x3(GemmRunner& g, cudaStream_t s, int B, int Hi, const float* in, const float* W, int weights,
                          const float* bias, int act, const float* mask, float* out, FullCorrScratch& scratch);

// Valid 3x3 convolution of a dense NHWC map without a column matrix: out[b, oy, ox, n] = act(bias[n] + sum_{ky, kx, c}
// in[b, oy + ky, ox + kx, c] * W[n, (ky, kx), c]), in [B, Hi, Hi, 32], out [B, Hi - 2, Hi - 2, 32].  The GEMM runs on the
// input's own grid (rows whose window crosses the border are computed and dropped by the compaction).
void valid_conv_3x3(GemmRunner& g, cudaStream_t s, int B, int Hi, const float* in, const float* W, int ldw,
                    const float* bias, int act, float* out, FullCorrScratch& scratch);

}  // namespace rlrep
