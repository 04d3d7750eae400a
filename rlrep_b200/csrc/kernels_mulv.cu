// Non-GEMM kernels of the muLV-Rep DrQ-v2 pixel update (reference: agent/mulvdrq/drqv2.py:313-461, vae.py:13-124):
// LayerNorm heads with / without tanh, the reparameterised sample, the KL + ELBO backward into both Gaussian heads, the
// noise-averaged critic's input expansion (fresh noise per forward, scaled by c_noise), Huber critic loss and the
// reward loss.  Gaussian heads live as two [B, ld] matrices (mean after tanh(LayerNorm), raw log-std after LayerNorm;
// consumers apply the [-20, 2] clamp), padding columns [D, ld) are kept at zero.  See kernels.cuh for the contracts.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "epilogue.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace rlrep {

namespace {

constexpr float kLsMin = -20.f, kLsMax = 2.f;  // vae.py:9-10
__device__ __forceinline__ float clamp_ls(float raw) { return fminf(fmaxf(raw, kLsMin), kLsMax); }
__device__ __forceinline__ float clamp_grad(float raw) { return (raw >= kLsMin && raw <= kLsMax) ? 1.f : 0.f; }

// One warp per row: y = act(LayerNorm_n(x) * gamma + beta), act = tanh or identity; columns [n, zero_to) of y zeroed.
__global__ void ln_act_fwd_kernel(const float* __restrict__ x, int ld_x, int B, int n, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int use_tanh, float* __restrict__ y, int ld_y,
                                  int zero_to, float* __restrict__ xhat, int ld_h, float* __restrict__ rstd) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (size_t)row * ld_x;
  float s = 0.f;
  for (int j = lane; j < n; j += 32) s += xr[j];
  const float mean = warp_sum(s) / (float)n;
  float v = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float d = xr[j] - mean;
    v = fmaf(d, d, v);
  }
  const float rs = rsqrtf(warp_sum(v) / (float)n + 1e-5f);
  for (int j = lane; j < zero_to || j < n; j += 32) {
    float out = 0.f;
    if (j < n) {
      const float h = (xr[j] - mean) * rs;
      if (xhat) xhat[(size_t)row * ld_h + j] = h;
      out = fmaf(h, __ldg(gamma + j), __ldg(beta + j));
      if (use_tanh) out = tanhf(out);
    }
    y[(size_t)row * ld_y + j] = out;
  }
  if (rstd && lane == 0) rstd[row] = rs;
}

// dz = dy * (1 - y^2) (tanh) or dy;  g_beta = dz, g_gamma = dz * xhat;  dx = rstd * (dxhat - mean(dxhat) - xhat *
// mean(dxhat * xhat)), dxhat = dz * gamma.  Columns [n, zero_to) of dx are zeroed.
__global__ void ln_act_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ y, int ld_y,
                                  const float* __restrict__ xhat, int ld_h, const float* __restrict__ rstd, int B, int n,
                                  const float* __restrict__ gamma, int use_tanh, float* __restrict__ dx, int ld_dx,
                                  int zero_to, float* __restrict__ g_beta, float* __restrict__ g_gamma, int ld_g) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane; j < n; j += 32) {
    float dz = dy[(size_t)row * ld_dy + j];
    if (use_tanh) {
      const float yy = y[(size_t)row * ld_y + j];
      dz *= 1.f - yy * yy;
    }
    const float h = xhat[(size_t)row * ld_h + j];
    g_beta[(size_t)row * ld_g + j] = dz;
    g_gamma[(size_t)row * ld_g + j] = dz * h;
    const float dh = dz * __ldg(gamma + j);
    s1 += dh;
    s2 = fmaf(dh, h, s2);
  }
  s1 = warp_sum(s1) / (float)n;
  s2 = warp_sum(s2) / (float)n;
  const float rs = rstd[row];
  for (int j = lane; j < zero_to || j < n; j += 32) {
    float out = 0.f;
    if (j < n) {
      const float h = xhat[(size_t)row * ld_h + j];
      const float dh = g_beta[(size_t)row * ld_g + j] * __ldg(gamma + j);
      out = rs * (dh - s1 - h * s2);
    }
    dx[(size_t)row * ld_dx + j] = out;
  }
}

__global__ void gauss_sample_kernel(const float* __restrict__ m, const float* __restrict__ raw, int ld, int B, int D,
                                    const float* __restrict__ eps, float* __restrict__ z, int ld_z) {
  const int total = B * ld_z;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / ld_z, d = i - b * ld_z;
    float v = 0.f;
    if (d < D) v = m[(size_t)b * ld + d] + eps[(size_t)b * D + d] * expf(clamp_ls(raw[(size_t)b * ld + d]));
    z[i] = v;
  }
}

// KL(N(m1, s1) || N(m2, s2)).mean() over B*D elements (drqv2.py:371-376) times `w`, plus the backward of the sample
// z = m1 + s1 eps (dz) into head 1 and the critic's gradient (dcm, dcraw; may be null) into head 2.
__global__ void __launch_bounds__(256) gauss_kl_bwd_kernel(const float* __restrict__ m1, const float* __restrict__ raw1,
                                                           const float* __restrict__ m2, const float* __restrict__ raw2,
                                                           int ld, int B, int D, const float* __restrict__ eps,
                                                           const float* __restrict__ dz, float w,
                                                           const float* __restrict__ dcm, const float* __restrict__ dcraw,
                                                           float* __restrict__ dm1, float* __restrict__ draw1,
                                                           float* __restrict__ dm2, float* __restrict__ draw2,
                                                           float* __restrict__ partial) {
  __shared__ float scratch[33];
  const int total = B * ld;
  const float g = w / ((float)B * (float)D);
  float kl_sum = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int b = i / ld, d = i - b * ld;
    float o1 = 0.f, o2 = 0.f, o3 = 0.f, o4 = 0.f;
    if (d < D) {
      const float a1 = m1[i], r1 = raw1[i], a2 = m2[i], r2 = raw2[i];
      const float ls1 = clamp_ls(r1), ls2 = clamp_ls(r2);
      const float var1 = expf(2.f * ls1), var2 = expf(2.f * ls2);
      const float diff = a1 - a2;
      const float q = (var1 + diff * diff) / var2;
      kl_sum += ls2 - ls1 + 0.5f * q - 0.5f;
      const float dzi = dz[i];
      const float dm = g * diff / var2;
      o1 = dzi + dm;
      o2 = (dzi * eps[(size_t)b * D + d] * expf(ls1) + g * (var1 / var2 - 1.f)) * clamp_grad(r1);
      o3 = -dm + (dcm ? dcm[i] : 0.f);
      o4 = g * (1.f - q) * clamp_grad(r2) + (dcraw ? dcraw[i] : 0.f);
    }
    dm1[i] = o1;
    draw1[i] = o2;
    dm2[i] = o3;
    draw2[i] = o4;
  }
  kl_sum = block_sum<256>(kl_sum, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = kl_sum;
}

// x[b*NN + j, d] = m[b, d] + (exp(clamp(raw[b, d])) * noise[j, d]) * c_noise   (Critic.forward, drqv2.py:177-183)
__global__ void gauss_noise_expand_kernel(const float* __restrict__ m, const float* __restrict__ raw, int ld, int B, int D,
                                          const float* __restrict__ noise, int NN, float c_noise, float* __restrict__ x,
                                          int ld_x) {
  const size_t total = (size_t)B * NN * ld_x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % ld_x);
    const size_t row = i / ld_x;
    const int j = (int)(row % NN), b = (int)(row / NN);
    float v = 0.f;
    if (d < D) {
      const float sd = expf(clamp_ls(raw[(size_t)b * ld + d]));
      v = m[(size_t)b * ld + d] + __fmul_rn(__fmul_rn(sd, __ldg(noise + (size_t)j * D + d)), c_noise);
    }
    x[i] = v;
  }
}

// dm[b, d] = sum_j dx;  draw[b, d] = (sum_j dx * noise_j) * c_noise * std * clamp'.  The raw gradient of the clamp lands
// on the log-std head's LayerNorm output.  Padding columns zeroed.
__global__ void gauss_noise_expand_bwd_kernel(const float* __restrict__ dx, int ld_x, const float* __restrict__ raw, int ld,
                                              int B, int D, const float* __restrict__ noise, int NN, float c_noise,
                                              float* __restrict__ dm, float* __restrict__ draw) {
  const int total = B * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / ld, d = i - b * ld;
    float o1 = 0.f, o2 = 0.f;
    if (d < D) {
      float sm = 0.f, sn = 0.f;
      for (int j = 0; j < NN; ++j) {
        const float g = dx[((size_t)b * NN + j) * ld_x + d];
        sm += g;
        sn = fmaf(g, __ldg(noise + (size_t)j * D + d), sn);
      }
      const float r = raw[i];
      o1 = sm;
      o2 = sn * c_noise * expf(clamp_ls(r)) * clamp_grad(r);
    }
    dm[i] = o1;
    draw[i] = o2;
  }
}

__global__ void group_mean_bwd_dact_kernel(const float* __restrict__ dmean, const float* __restrict__ hid, int ld, int B,
                                           int NN, int C, int dact, float* __restrict__ dhid,
                                           float* __restrict__ colsum_partial) {
  const int total = B * C;
  const float inv = 1.f / (float)NN;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / C, c = i - b * C;
    const float g = dmean[i] * inv;
    float acc = 0.f;
    for (int j = 0; j < NN; ++j) {
      const size_t o = ((size_t)b * NN + j) * ld + c;
      const float v = g * apply_dact(hid[o], dact);
      dhid[o] = v;
      acc += v;
    }
    colsum_partial[i] = acc;
  }
}

// y = r + discount * min(tq1, tq2);  loss = smooth_l1(q1, y) + smooth_l1(q2, y) (beta = 1, mean over B each);
// dq = clamp(q - y, -1, 1) / B;  metrics = {critic_loss, mean(q1), mean(q2), mean(y)}
__global__ void __launch_bounds__(256) huber_critic_loss_kernel(const float* __restrict__ reward,
                                                                const float* __restrict__ discount,
                                                                const float* __restrict__ tq1,
                                                                const float* __restrict__ tq2, const float* __restrict__ q1,
                                                                const float* __restrict__ q2, int B, float* __restrict__ dq1,
                                                                float* __restrict__ dq2, float* __restrict__ metrics) {
  __shared__ float scratch[33];
  float l = 0.f, s1 = 0.f, s2 = 0.f, st = 0.f;
  const float inv = 1.f / (float)B;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float y = reward[i] + discount[i] * fminf(tq1[i], tq2[i]);
    const float e1 = q1[i] - y, e2 = q2[i] - y;
    const float a1 = fabsf(e1), a2 = fabsf(e2);
    l += a1 < 1.f ? 0.5f * e1 * e1 : a1 - 0.5f;
    l += a2 < 1.f ? 0.5f * e2 * e2 : a2 - 0.5f;
    s1 += q1[i];
    s2 += q2[i];
    st += y;
    dq1[i] = fminf(fmaxf(e1, -1.f), 1.f) * inv;
    dq2[i] = fminf(fmaxf(e2, -1.f), 1.f) * inv;
  }
  l = block_sum<256>(l, scratch);
  s1 = block_sum<256>(s1, scratch);
  s2 = block_sum<256>(s2, scratch);
  st = block_sum<256>(st, scratch);
  if (threadIdx.x == 0) {
    metrics[0] = l * inv;
    metrics[1] = s1 * inv;
    metrics[2] = s2 * inv;
    metrics[3] = st * inv;
  }
}

// r_loss = mse(r_hat, r) (mean over B);  dr = w * 2 (r_hat - r) / B;  out[0] = r_loss
__global__ void __launch_bounds__(256) reward_mse_kernel(const float* __restrict__ r_hat, const float* __restrict__ reward,
                                                         int B, float w, float* __restrict__ dr, float* __restrict__ out) {
  __shared__ float scratch[33];
  float l = 0.f;
  const float inv = 1.f / (float)B;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float e = r_hat[i] - reward[i];
    l = fmaf(e, e, l);
    dr[i] = w * 2.f * e * inv;
  }
  l = block_sum<256>(l, scratch);
  if (threadIdx.x == 0) out[0] = l * inv;
}

int blocks_for(size_t work, int threads, int per_sm = 8) {
  const size_t want = (work + threads - 1) / threads;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)kNumSMs * per_sm));
}

}  // namespace

void launch_ln_act_fwd(const float* x, int ld_x, int B, int n, const float* gamma, const float* beta, bool use_tanh,
                       float* y, int ld_y, int zero_to, float* xhat, int ld_h, float* rstd, cudaStream_t s) {
  ln_act_fwd_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(x, ld_x, B, n, gamma, beta, use_tanh ? 1 : 0, y, ld_y, zero_to,
                                                         xhat, ld_h, rstd);
  RLREP_LAUNCHED("ln_act_fwd", s);
}
void launch_ln_act_bwd(const float* dy, int ld_dy, const float* y, int ld_y, const float* xhat, int ld_h, const float* rstd,
                       int B, int n, const float* gamma, bool use_tanh, float* dx, int ld_dx, int zero_to, float* g_beta,
                       float* g_gamma, int ld_g, cudaStream_t s) {
  ln_act_bwd_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(dy, ld_dy, y, ld_y, xhat, ld_h, rstd, B, n, gamma,
                                                         use_tanh ? 1 : 0, dx, ld_dx, zero_to, g_beta, g_gamma, ld_g);
  RLREP_LAUNCHED("ln_act_bwd", s);
}
void launch_gauss_sample(const float* m, const float* raw, int ld, int B, int D, const float* eps, float* z, int ld_z,
                         cudaStream_t s) {
  gauss_sample_kernel<<<blocks_for((size_t)B * ld_z, 256), 256, 0, s>>>(m, raw, ld, B, D, eps, z, ld_z);
  RLREP_LAUNCHED("gauss_sample", s);
}
void launch_gauss_kl_bwd(const float* m1, const float* raw1, const float* m2, const float* raw2, int ld, int B, int D,
                         const float* eps, const float* dz, float w, const float* dcm, const float* dcraw, float* dm1,
                         float* draw1, float* dm2, float* draw2, float* partial, int n_blocks, cudaStream_t s) {
  gauss_kl_bwd_kernel<<<n_blocks, 256, 0, s>>>(m1, raw1, m2, raw2, ld, B, D, eps, dz, w, dcm, dcraw, dm1, draw1, dm2, draw2,
                                               partial);
  RLREP_LAUNCHED("gauss_kl_bwd", s);
}
void launch_gauss_noise_expand(const float* m, const float* raw, int ld, int B, int D, const float* noise, int NN,
                               float c_noise, float* x, int ld_x, cudaStream_t s) {
  gauss_noise_expand_kernel<<<blocks_for((size_t)B * NN * ld_x, 256, 16), 256, 0, s>>>(m, raw, ld, B, D, noise, NN, c_noise,
                                                                                     x, ld_x);
  RLREP_LAUNCHED("gauss_noise_expand", s);
}
void launch_gauss_noise_expand_bwd(const float* dx, int ld_x, const float* raw, int ld, int B, int D, const float* noise,
                                   int NN, float c_noise, float* dm, float* draw, cudaStream_t s) {
  gauss_noise_expand_bwd_kernel<<<blocks_for((size_t)B * ld, 256), 256, 0, s>>>(dx, ld_x, raw, ld, B, D, noise, NN, c_noise,
                                                                               dm, draw);
  RLREP_LAUNCHED("gauss_noise_expand_bwd", s);
}
void launch_group_mean_bwd_dact(const float* dmean, const float* hid, int ld, int B, int NN, int C, int dact, float* dhid,
                                float* colsum_partial, cudaStream_t s) {
  group_mean_bwd_dact_kernel<<<blocks_for((size_t)B * C, 256), 256, 0, s>>>(dmean, hid, ld, B, NN, C, dact, dhid,
                                                                           colsum_partial);
  RLREP_LAUNCHED("group_mean_bwd", s);
}
void launch_huber_critic_loss(const float* reward, const float* discount, const float* tq1, const float* tq2,
                              const float* q1, const float* q2, int B, float* dq1, float* dq2, float* metrics,
                              cudaStream_t s) {
  huber_critic_loss_kernel<<<1, 256, 0, s>>>(reward, discount, tq1, tq2, q1, q2, B, dq1, dq2, metrics);
  RLREP_LAUNCHED("huber_critic_loss", s);
}
void launch_reward_mse(const float* r_hat, const float* reward, int B, float w, float* dr, float* out, cudaStream_t s) {
  reward_mse_kernel<<<1, 256, 0, s>>>(r_hat, reward, B, w, dr, out);
  RLREP_LAUNCHED("reward_mse", s);
}

}  // namespace rlrep
