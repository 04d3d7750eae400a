// Non-GEMM kernels of the DrQ-v2 pixel update (reference: agent/diffsrdrq/network_arch/drqv2.py, drqv2.py:93-148):
// LayerNorm + tanh trunk (forward / backward), TruncatedNormal sampling with the straight-through clamp, the critic
// (MSE on stacked twin Q) and actor losses.  See kernels.cuh for the contracts.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace rlrep {

namespace {

// One warp per row.  y = tanh(LayerNorm(x) * gamma + beta), eps = 1e-5, biased variance (nn.LayerNorm).
__global__ void ln_tanh_fwd_kernel(const float* __restrict__ x, int ld_x, int B, int n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ y, int ld_y,
                                   float* __restrict__ xhat, int ld_h, float* __restrict__ rstd) {
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
  for (int j = lane; j < n; j += 32) {
    const float h = (xr[j] - mean) * rs;
    if (xhat) xhat[(size_t)row * ld_h + j] = h;
    y[(size_t)row * ld_y + j] = tanhf(fmaf(h, __ldg(gamma + j), __ldg(beta + j)));
  }
  if (rstd && lane == 0) rstd[row] = rs;
}

// dz = dy * (1 - y^2);  g_beta = dz, g_gamma = dz * xhat (column sums are the LayerNorm parameter gradients);
// dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)),  dxhat = dz * gamma.  Columns [n, ld_dx) of dx are zeroed.
__global__ void ln_tanh_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ y, int ld_y,
                                   const float* __restrict__ xhat, int ld_h, const float* __restrict__ rstd, int B, int n,
                                   const float* __restrict__ gamma, float* __restrict__ dx, int ld_dx,
                                   float* __restrict__ g_beta, float* __restrict__ g_gamma, int ld_g) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float yy = y[(size_t)row * ld_y + j];
    const float dz = dy[(size_t)row * ld_dy + j] * (1.f - yy * yy);
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
  for (int j = lane; j < ld_dx; j += 32) {
    float out = 0.f;
    if (j < n) {
      const float h = xhat[(size_t)row * ld_h + j];
      const float dh = g_beta[(size_t)row * ld_g + j] * __ldg(gamma + j);
      out = rs * (dh - s1 - h * s2);
    }
    dx[(size_t)row * ld_dx + j] = out;
  }
}

// mu = tanh(raw);  a = clamp(mu + clamp(eps * std, -clip, clip), -1 + 1e-6, 1 - 1e-6)   (TruncatedNormal.sample)
__global__ void trunc_normal_sample_kernel(const float* __restrict__ raw, int ld_raw, int B, int A,
                                           const float* __restrict__ eps, float std, float clip, float* __restrict__ mu,
                                           float* __restrict__ action, int ld_a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * A) return;
  const int b = i / A, j = i - b * A;
  const float m = tanhf(raw[(size_t)b * ld_raw + j]);
  float e = __fmul_rn(eps[i], std);
  e = fminf(fmaxf(e, -clip), clip);
  const float x = __fadd_rn(m, e);
  mu[i] = m;
  action[(size_t)b * ld_a + j] = fminf(fmaxf(x, -0.999999f), 0.999999f);
}
// straight-through clamp: d action / d mu = 1;  d raw = d action * (1 - mu^2); padding columns [A, ld_out) zeroed
__global__ void trunc_normal_bwd_kernel(const float* __restrict__ d_action, int ld_da, const float* __restrict__ mu, int B,
                                        int A, float* __restrict__ draw, int ld_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * ld_out) return;
  const int b = i / ld_out, j = i - b * ld_out;
  float v = 0.f;
  if (j < A) {
    const float m = mu[b * A + j];
    v = d_action[(size_t)b * ld_da + j] * (1.f - m * m);
  }
  draw[i] = v;
}

// F.mse_loss(q_pred [2,B,1], target repeated): mean over 2B elements.  metrics = {critic_loss, mean(q_pred),
// mean(q_target), mean(reward)}
__global__ void __launch_bounds__(256) drq_critic_loss_kernel(const float* __restrict__ reward,
                                                              const float* __restrict__ discount,
                                                              const float* __restrict__ tq1, const float* __restrict__ tq2,
                                                              const float* __restrict__ q1, const float* __restrict__ q2,
                                                              int B, float* __restrict__ dq1, float* __restrict__ dq2,
                                                              float* __restrict__ metrics) {
  __shared__ float scratch[33];
  float l = 0.f, sq = 0.f, st = 0.f, sr = 0.f;
  const float inv = 1.f / (2.f * (float)B);
  for (int i = threadIdx.x; i < B; i += 256) {
    const float y = reward[i] + discount[i] * fminf(tq1[i], tq2[i]);
    const float e1 = q1[i] - y, e2 = q2[i] - y;
    l = fmaf(e1, e1, l);
    l = fmaf(e2, e2, l);
    sq += q1[i] + q2[i];
    st += y;
    sr += reward[i];
    dq1[i] = 2.f * e1 * inv;
    dq2[i] = 2.f * e2 * inv;
  }
  l = block_sum<256>(l, scratch);
  sq = block_sum<256>(sq, scratch);
  st = block_sum<256>(st, scratch);
  sr = block_sum<256>(sr, scratch);
  if (threadIdx.x == 0) {
    metrics[0] = l * inv;
    metrics[1] = sq * inv;
    metrics[2] = st / (float)B;
    metrics[3] = sr / (float)B;
  }
}

// actor_loss = -mean(min(q1, q2)); the gradient goes to the argmin (first index on ties, like torch.min(dim))
__global__ void __launch_bounds__(256) drq_actor_loss_kernel(const float* __restrict__ q1, const float* __restrict__ q2,
                                                             int B, float* __restrict__ dq1, float* __restrict__ dq2,
                                                             float* __restrict__ metrics) {
  __shared__ float scratch[33];
  float s = 0.f;
  const float g = -1.f / (float)B;
  for (int i = threadIdx.x; i < B; i += 256) {
    const float a = q1[i], b = q2[i];
    s += fminf(a, b);
    dq1[i] = a <= b ? g : 0.f;
    dq2[i] = a <= b ? 0.f : g;
  }
  s = block_sum<256>(s, scratch);
  if (threadIdx.x == 0) metrics[0] = -s / (float)B;
}

}  // namespace

void launch_ln_tanh_fwd(const float* x, int ld_x, int B, int n, const float* gamma, const float* beta, float* y, int ld_y,
                        float* xhat, int ld_h, float* rstd, cudaStream_t s) {
  ln_tanh_fwd_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(x, ld_x, B, n, gamma, beta, y, ld_y, xhat, ld_h, rstd);
  RLREP_LAUNCHED("ln_tanh_fwd", s);
}
void launch_ln_tanh_bwd(const float* dy, int ld_dy, const float* y, int ld_y, const float* xhat, int ld_h, const float* rstd,
                        int B, int n, const float* gamma, float* dx, int ld_dx, float* g_beta, float* g_gamma, int ld_g,
                        cudaStream_t s) {
  ln_tanh_bwd_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(dy, ld_dy, y, ld_y, xhat, ld_h, rstd, B, n, gamma, dx, ld_dx,
                                                          g_beta, g_gamma, ld_g);
  RLREP_LAUNCHED("ln_tanh_bwd", s);
}
void launch_trunc_normal_sample(const float* raw, int ld_raw, int B, int A, const float* eps, float std, float clip,
                                float* mu, float* action, int ld_a, cudaStream_t s) {
  trunc_normal_sample_kernel<<<ceil_div(B * A, 128), 128, 0, s>>>(raw, ld_raw, B, A, eps, std, clip, mu, action, ld_a);
  RLREP_LAUNCHED("trunc_normal_sample", s);
}
void launch_trunc_normal_bwd(const float* d_action, int ld_da, const float* mu, int B, int A, float* draw, int ld_out,
                             cudaStream_t s) {
  trunc_normal_bwd_kernel<<<ceil_div(B * ld_out, 128), 128, 0, s>>>(d_action, ld_da, mu, B, A, draw, ld_out);
  RLREP_LAUNCHED("trunc_normal_bwd", s);
}
void launch_drq_critic_loss(const float* reward, const float* discount, const float* tq1, const float* tq2, const float* q1,
                            const float* q2, int B, float* dq1, float* dq2, float* metrics, cudaStream_t s) {
  drq_critic_loss_kernel<<<1, 256, 0, s>>>(reward, discount, tq1, tq2, q1, q2, B, dq1, dq2, metrics);
  RLREP_LAUNCHED("drq_critic_loss", s);
}
void launch_drq_actor_loss(const float* q1, const float* q2, int B, float* dq1, float* dq2, float* metrics, cudaStream_t s) {
  drq_actor_loss_kernel<<<1, 256, 0, s>>>(q1, q2, B, dq1, dq2, metrics);
  RLREP_LAUNCHED("drq_actor_loss", s);
}

}  // namespace rlrep
