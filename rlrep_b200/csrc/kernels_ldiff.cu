// DRAFT (branch draft/ldiffsr-agent; not verified on hardware).
// Non-GEMM kernels of the latent Diff-SR DrQ-v2 pixel update (reference: agent/diffsrdrq/latent_diff_sr.py:234-390,
// network_arch/vae_1d.py:24-60,118-131, score_idql.py:9-70): LayerNorm with {none, tanh, swish}, Mish, dropout-mask
// application, the diagonal-Gaussian posterior (sample, KL, backward), the DDPM perturbation of the next latent.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace rlrep {

namespace {

int blocks_for(size_t work, int threads, int per_sm = 8) {
  const size_t want = (work + threads - 1) / threads;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)kNumSMs * per_sm));
}

__device__ __forceinline__ float sigmoidf_(float u) { return 1.f / (1.f + expf(-u)); }
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch threshold 20

// One warp per row: y = act(LayerNorm_n(x) * gamma + beta); act 0 none, 1 tanh, 2 swish; columns [n, zero_to) zeroed.
__global__ void ln_act2_fwd_kernel(const float* __restrict__ x, int ld_x, int B, int n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int act, float* __restrict__ y, int ld_y, int zero_to,
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
  for (int j = lane; j < zero_to || j < n; j += 32) {
    float out = 0.f;
    if (j < n) {
      const float h = (xr[j] - mean) * rs;
      if (xhat) xhat[(size_t)row * ld_h + j] = h;
      out = fmaf(h, __ldg(gamma + j), __ldg(beta + j));
      if (act == 1) out = tanhf(out);
      else if (act == 2) out = out * sigmoidf_(out);
    }
    y[(size_t)row * ld_y + j] = out;
  }
  if (rstd && lane == 0) rstd[row] = rs;
}

__global__ void ln_act2_bwd_kernel(const float* __restrict__ dy, int ld_dy, const float* __restrict__ y, int ld_y,
                                   const float* __restrict__ xhat, int ld_h, const float* __restrict__ rstd, int B, int n,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                   float* __restrict__ dx, int ld_dx, int zero_to, float* __restrict__ g_beta,
                                   float* __restrict__ g_gamma, int ld_g) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  float s1 = 0.f, s2 = 0.f;
  for (int j = lane; j < n; j += 32) {
    float dz = dy[(size_t)row * ld_dy + j];
    const float h = xhat[(size_t)row * ld_h + j];
    if (act == 1) {
      const float yy = y[(size_t)row * ld_y + j];
      dz *= 1.f - yy * yy;
    } else if (act == 2) {  // d swish(u) = s (1 + u (1 - s)), u recomputed from xhat
      const float u = fmaf(h, __ldg(gamma + j), __ldg(beta + j));
      const float sg = sigmoidf_(u);
      dz *= sg * (1.f + u * (1.f - sg));
    }
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

// y = mish(x) = x tanh(softplus(x));  dx = dy (t + x (1 - t^2) sigmoid(x)), t = tanh(softplus(x))
__global__ void mish_fwd_kernel(const float* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v * tanhf(softplusf_(v));
  }
}
__global__ void mish_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, size_t n, float* __restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float t = tanhf(softplusf_(v));
    dx[i] = dy[i] * (t + v * (1.f - t * t) * sigmoidf_(v));
  }
}

// out = (accumulate ? out : 0) + x * (mask ? mask * inv_keep : 1)   (F.dropout with a host-drawn Bernoulli mask)
__global__ void mask_scale_kernel(const float* __restrict__ x, const float* __restrict__ mask, float inv_keep, size_t n,
                                  int accumulate, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = mask ? x[i] * (mask[i] * inv_keep) : x[i];
    out[i] = accumulate ? out[i] + v : v;
  }
}

// h [N, 2L] = (mean | raw logvar): logvar = clamp(raw, -30, 20); z = mean + exp(0.5 logvar) eps; mean_out / z contiguous [N, L];
// partial[block] = sum over elements of 0.5 (mean^2 + var - 1 - logvar)
__global__ void __launch_bounds__(256) posterior_fwd_kernel(const float* __restrict__ h, int N, int L, const float* __restrict__ eps,
                                                            float* __restrict__ mean_out, float* __restrict__ z,
                                                            float* __restrict__ partial) {
  __shared__ float scratch[33];
  const int total = N * L;
  float kl = 0.f;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int r = i / L, j = i - r * L;
    const float m = h[(size_t)r * 2 * L + j];
    const float lv = fminf(fmaxf(h[(size_t)r * 2 * L + L + j], -30.f), 20.f);
    const float sd = expf(0.5f * lv), var = expf(lv);
    mean_out[i] = m;
    z[i] = m + sd * eps[i];
    kl += 0.5f * (m * m + var - 1.f - lv);
  }
  kl = block_sum<256>(kl, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = kl;
}
// dh = (dz + dmean_extra + w_kl mean | (dz eps 0.5 std + w_kl 0.5 (var - 1)) * clamp'),  w_kl = kl weight / N
__global__ void posterior_bwd_kernel(const float* __restrict__ h, int N, int L, const float* __restrict__ eps,
                                     const float* __restrict__ dz, const float* __restrict__ dmean_extra, int n_extra,
                                     float w_kl, float* __restrict__ dh) {
  const int total = N * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / L, j = i - r * L;
    const float m = h[(size_t)r * 2 * L + j], raw = h[(size_t)r * 2 * L + L + j];
    const float lv = fminf(fmaxf(raw, -30.f), 20.f);
    const float sd = expf(0.5f * lv), var = expf(lv);
    const float g = dz[i];
    const float extra = (dmean_extra != nullptr && i < n_extra) ? dmean_extra[i] : 0.f;
    dh[(size_t)r * 2 * L + j] = g + extra + w_kl * m;
    const float pass = (raw >= -30.f && raw <= 20.f) ? 1.f : 0.f;
    dh[(size_t)r * 2 * L + L + j] = (g * eps[i] * 0.5f * sd + w_kl * 0.5f * (var - 1.f)) * pass;
  }
}

// zin[b] = [ sqrt(ab_b) x_b + sqrt(1 - ab_b) noise_b | t_emb_b | 0 ... ] (row pitch ld);  target = -noise;
// coef_b = sqrt(1 - ab_b) / feat   (latent_diff_sr.py:279-285 with score_idql.py:191's division folded into coef)
__global__ void ldiff_perturb_kernel(const float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ ab,
                                     const float* __restrict__ temb, int B, int L, int T, float inv_feat,
                                     float* __restrict__ zin, int ld, float* __restrict__ target, float* __restrict__ coef) {
  const int total = B * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / ld, j = i - b * ld;
    const float a = ab[b];
    float v = 0.f;
    if (j < L) {
      const float nz = noise[(size_t)b * L + j];
      v = sqrtf(a) * x[(size_t)b * L + j] + sqrtf(1.f - a) * nz;
      target[(size_t)b * L + j] = -nz;
    } else if (j < L + T) {
      v = temb[(size_t)b * T + (j - L)];
    }
    zin[i] = v;
    if (j == 0) coef[b] = sqrtf(1.f - a) * inv_feat;
  }
}
// dx[b, j] (+)= sqrt(ab_b) * dzin[b, j]
__global__ void ldiff_perturb_bwd_kernel(const float* __restrict__ dzin, int ld, const float* __restrict__ ab, int B, int L,
                                         float* __restrict__ dx) {
  const int total = B * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / L, j = i - b * L;
    dx[i] += sqrtf(ab[b]) * dzin[(size_t)b * ld + j];
  }
}
// y[i] += x[i]
__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += x[i];
}
// p *= s (AdamW's decoupled weight decay, applied before the Adam step: torch/optim/adamw.py)
__global__ void scale_inplace_kernel(float4* __restrict__ p, size_t n4, float s) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = p[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    p[i] = v;
  }
}

}  // namespace

void launch_ln_act2_fwd(const float* x, int ld_x, int B, int n, const float* gamma, const float* beta, int act, float* y,
                        int ld_y, int zero_to, float* xhat, int ld_h, float* rstd, cudaStream_t s) {
  ln_act2_fwd_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(x, ld_x, B, n, gamma, beta, act, y, ld_y, zero_to, xhat, ld_h, rstd);
  RLREP_LAUNCHED("ln_act2_fwd", s);
}
void launch_ln_act2_bwd(const float* dy, int ld_dy, const float* y, int ld_y, const float* xhat, int ld_h, const float* rstd,
                        int B, int n, const float* gamma, const float* beta, int act, float* dx, int ld_dx, int zero_to,
                        float* g_beta, float* g_gamma, int ld_g, cudaStream_t s) {
  ln_act2_bwd_kernel<<<ceil_div(B * 32, 128), 128, 0, s>>>(dy, ld_dy, y, ld_y, xhat, ld_h, rstd, B, n, gamma, beta, act, dx,
                                                          ld_dx, zero_to, g_beta, g_gamma, ld_g);
  RLREP_LAUNCHED("ln_act2_bwd", s);
}
void launch_mish_fwd(const float* x, size_t n, float* y, cudaStream_t s) {
  mish_fwd_kernel<<<blocks_for(n, 256, 16), 256, 0, s>>>(x, n, y);
  RLREP_LAUNCHED_W("mish_fwd", s, 8.0 * n, 0.0);
}
void launch_mish_bwd(const float* dy, const float* x, size_t n, float* dx, cudaStream_t s) {
  mish_bwd_kernel<<<blocks_for(n, 256, 16), 256, 0, s>>>(dy, x, n, dx);
  RLREP_LAUNCHED_W("mish_bwd", s, 12.0 * n, 0.0);
}
void launch_mask_scale(const float* x, const float* mask, float inv_keep, size_t n, int accumulate, float* out, cudaStream_t s) {
  mask_scale_kernel<<<blocks_for(n, 256, 16), 256, 0, s>>>(x, mask, inv_keep, n, accumulate, out);
  RLREP_LAUNCHED("mask_scale", s);
}
void launch_posterior_fwd(const float* h, int N, int L, const float* eps, float* mean_out, float* z, float* partial,
                          int n_blocks, cudaStream_t s) {
  posterior_fwd_kernel<<<n_blocks, 256, 0, s>>>(h, N, L, eps, mean_out, z, partial);
  RLREP_LAUNCHED("posterior_fwd", s);
}
void launch_posterior_bwd(const float* h, int N, int L, const float* eps, const float* dz, const float* dmean_extra,
                          int n_extra, float w_kl, float* dh, cudaStream_t s) {
  posterior_bwd_kernel<<<blocks_for((size_t)N * L, 256), 256, 0, s>>>(h, N, L, eps, dz, dmean_extra, n_extra, w_kl, dh);
  RLREP_LAUNCHED("posterior_bwd", s);
}
void launch_ldiff_perturb(const float* x, const float* noise, const float* ab, const float* temb, int B, int L, int T,
                          float inv_feat, float* zin, int ld, float* target, float* coef, cudaStream_t s) {
  ldiff_perturb_kernel<<<blocks_for((size_t)B * ld, 256), 256, 0, s>>>(x, noise, ab, temb, B, L, T, inv_feat, zin, ld, target,
                                                                      coef);
  RLREP_LAUNCHED("ldiff_perturb", s);
}
void launch_ldiff_perturb_bwd(const float* dzin, int ld, const float* ab, int B, int L, float* dx, cudaStream_t s) {
  ldiff_perturb_bwd_kernel<<<blocks_for((size_t)B * L, 256), 256, 0, s>>>(dzin, ld, ab, B, L, dx);
  RLREP_LAUNCHED("ldiff_perturb_bwd", s);
}
void launch_add_inplace(float* y, const float* x, size_t n, cudaStream_t s) {
  add_inplace_kernel<<<blocks_for(n, 256, 16), 256, 0, s>>>(y, x, n);
  RLREP_LAUNCHED("add_inplace", s);
}
void launch_scale_inplace(float* p, size_t n, float scale, cudaStream_t s) {
  scale_inplace_kernel<<<blocks_for(n / 4, 256, 16), 256, 0, s>>>(reinterpret_cast<float4*>(p), n / 4, scale);
  RLREP_LAUNCHED_W("scale_inplace", s, 8.0 * n, 0.0);
}

}  // namespace rlrep
