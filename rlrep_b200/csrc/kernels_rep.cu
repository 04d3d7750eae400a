// Representation-loss kernels of SPEDER-SAC (agent/spedersac/spedersac_agent.py:181-219) and Diff-SR-SAC
// (agent/diffsrsac/diffsrsac_agent.py:271-318).  See kernels.cuh for the contract of each launcher.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "kernels.cuh"
#include "reduce.cuh"

namespace rlrep {

namespace {

// ------------------------------------------------------------------------------------------- SPEDER
// One warp per row i < B: diag_i = <phi_i, mu_i>, rpred_i = <phi_i, theta> + b, c_i = <mu~_i, u>.
// One CTA per row (a warp per row left 32 SMs with 64 dependent scalar iterations each: 21 us at B = 256, D = 2048): float4
// loads, three block sums in a fixed order.
__global__ void __launch_bounds__(256) speder_rows_kernel(const float* __restrict__ zphi, const float* __restrict__ zmu,
                                                          int D, int B, const float* __restrict__ theta_w,
                                                          const float* __restrict__ theta_b, const float* __restrict__ u,
                                                          float* __restrict__ diag, float* __restrict__ rpred,
                                                          float* __restrict__ c) {
  __shared__ float scratch[33];
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* p = zphi + (size_t)row * D;
  const float* m = zmu + (size_t)row * D;
  const float* mr = zmu + (size_t)(B + row) * D;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(zphi) | reinterpret_cast<uintptr_t>(zmu) |
                                     reinterpret_cast<uintptr_t>(theta_w) | reinterpret_cast<uintptr_t>(u)) & 15) == 0;
  if (vec) {
    const float4 *p4 = reinterpret_cast<const float4*>(p), *m4 = reinterpret_cast<const float4*>(m),
                 *r4 = reinterpret_cast<const float4*>(mr), *t4 = reinterpret_cast<const float4*>(theta_w),
                 *u4 = reinterpret_cast<const float4*>(u);
    for (int j = tid; j < D / 4; j += 256) {
      const float4 pv = p4[j], mv = m4[j], rv = r4[j], tv = __ldg(t4 + j), uv = u4[j];
      a0 = fmaf(pv.x, mv.x, fmaf(pv.y, mv.y, fmaf(pv.z, mv.z, fmaf(pv.w, mv.w, a0))));
      a1 = fmaf(pv.x, tv.x, fmaf(pv.y, tv.y, fmaf(pv.z, tv.z, fmaf(pv.w, tv.w, a1))));
      a2 = fmaf(rv.x, uv.x, fmaf(rv.y, uv.y, fmaf(rv.z, uv.z, fmaf(rv.w, uv.w, a2))));
    }
  } else {
    for (int j = tid; j < D; j += 256) {
      const float pj = p[j];
      a0 = fmaf(pj, m[j], a0);
      a1 = fmaf(pj, __ldg(theta_w + j), a1);
      a2 = fmaf(mr[j], u[j], a2);
    }
  }
  a0 = block_sum<256>(a0, scratch);
  a1 = block_sum<256>(a1, scratch);
  a2 = block_sum<256>(a2, scratch);
  if (tid == 0) {
    diag[row] = a0;
    rpred[row] = a1 + __ldg(theta_b);
    c[row] = a2;
  }
}

__global__ void __launch_bounds__(256) speder_finalize_kernel(const float* __restrict__ diag, const float* __restrict__ c,
                                                              const float* __restrict__ rpred,
                                                              const float* __restrict__ reward, int ld_r, int B,
                                                              float* __restrict__ drp, float* __restrict__ metrics) {
  __shared__ float scratch[33];
  const float inv_b = 1.f / (float)B;
  float sd = 0.f, sc = 0.f, se = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) {
    sd += diag[i];
    sc = fmaf(c[i], c[i], sc);
    const float d = rpred[i] - reward[(size_t)i * ld_r];
    se = fmaf(d, d, se);
    drp[i] = d * inv_b;
  }
  sd = block_sum<256>(sd, scratch);
  sc = block_sum<256>(sc, scratch);
  se = block_sum<256>(se, scratch);
  if (threadIdx.x == 0) {
    const float model = -2.f * sd * inv_b + sc * inv_b * inv_b;
    const float r = 0.5f * (se * inv_b);
    metrics[0] = model + r;
    metrics[1] = model;
    metrics[2] = r;
  }
}

__global__ void speder_grad_kernel(const float* __restrict__ zphi, const float* __restrict__ zmu, int D, int B,
                                   const float* __restrict__ drp, const float* __restrict__ theta_w,
                                   const float* __restrict__ c, const float* __restrict__ u,
                                   const float* __restrict__ w, float* __restrict__ dzphi, float* __restrict__ dzmu) {
  const float k1 = -2.f / (float)B, k2 = 2.f / ((float)B * (float)B);
  const size_t half = (size_t)B * D, total = 2 * half;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), j = (int)(i - (size_t)r * D);
    if (i < half) {
      dzphi[i] = fmaf(drp[r], __ldg(theta_w + j), k1 * zmu[i]);
      dzmu[i] = k1 * zphi[i];
    } else {
      dzphi[i] = k2 * w[j];
      dzmu[i] = k2 * c[r - B] * u[j];
    }
  }
}

// ------------------------------------------------------------------------------------------- Diff-SR
__global__ void diffsr_perturb_kernel(const float* __restrict__ next_state, int ld_rec, const float* __restrict__ noise,
                                      const long long* __restrict__ level, const float* __restrict__ alphabars,
                                      int n_levels, float sigma, int B, int S, float* __restrict__ xin, int ld_x,
                                      float* __restrict__ target, float* __restrict__ coef) {
  const int total = B * ld_x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / ld_x, j = i - b * ld_x;
    long long lv = level[b];
    lv = lv < 0 ? 0 : (lv >= n_levels ? n_levels - 1 : lv);
    const float ab = alphabars[lv];
    float out = 0.f;
    if (j < S) {
      // torch ops are separately rounded: sqrt(ab) * s' + sqrt(1 - ab) * n;  target = -(s~' - sqrt(ab) * s')
      const float sa = sqrtf(ab), sb = sqrtf(__fsub_rn(1.0f, ab));
      const float scaled = __fmul_rn(sa, next_state[(size_t)b * ld_rec + j]);
      out = __fadd_rn(scaled, __fmul_rn(sb, noise[(size_t)b * S + j]));
      target[(size_t)b * S + j] = -__fsub_rn(out, scaled);
    } else if (j == S) {
      out = ab;
      coef[b] = __fmul_rn(__fsub_rn(1.0f, ab), sigma);
    }
    xin[i] = out;
  }
}

// One CTA per batch row.  Threads are laid out as (s, group): thread (s, g) sums phi[d] * flat[d*S + s] over
// d = g, g + G, ...; consecutive s are consecutive addresses (coalesced); the groups are combined in a fixed order.
__global__ void __launch_bounds__(256) diffsr_score_kernel(const float* __restrict__ phi, const float* __restrict__ flat,
                                                           int D, int S, const float* __restrict__ target,
                                                           const float* __restrict__ coef, int B,
                                                           float* __restrict__ dscore, float* __restrict__ loss_rows) {
  extern __shared__ float sm[];  // [G][S_tile] partial scores, then 33 floats of reduction scratch
  const int b = blockIdx.x;
  const float* f = flat + (size_t)b * D * S;
  const float* p = phi + (size_t)b * D;
  const int s_tile = S < 256 ? S : 256;
  const int G = 256 / s_tile;
  const int g = threadIdx.x / s_tile, sl = threadIdx.x - g * s_tile;
  float* scratch = sm + G * s_tile;
  const float cf = coef[b], two_over_b = 2.f / (float)B;
  float row_loss = 0.f;
  for (int s0 = 0; s0 < S; s0 += s_tile) {
    const int s = s0 + sl;
    float acc = 0.f;
    if (g < G && s < S)
      for (int d = g; d < D; d += G) acc = fmaf(p[d], f[(size_t)d * S + s], acc);
    if (g < G) sm[g * s_tile + sl] = acc;
    __syncthreads();
    if (g == 0 && s < S) {
      float score = 0.f;
      for (int k = 0; k < G; ++k) score += sm[k * s_tile + sl];
      const float diff = target[(size_t)b * S + s] - cf * score;
      row_loss = fmaf(diff, diff, row_loss);
      dscore[(size_t)b * S + s] = -cf * two_over_b * diff;
    }
    __syncthreads();
  }
  row_loss = block_sum<256>(row_loss, scratch);
  if (threadIdx.x == 0) loss_rows[b] = row_loss;
}

// dflat[b, d*S + s] = phi[b, d] * dscore[b, s];  dphi[b, d] = sum_s dscore[b, s] * flat[b, d*S + s].
// One CTA per batch row: dflat over the row's D*S contiguous elements (coalesced, every lane busy), then dphi[b, d] by
// thread d in a fixed order over s.  The warp-per-(b, d) version kept 17 of 32 lanes busy for one iteration on 65k warps
// (11.5 us for 9 MB at HalfCheetah's S = 17).
__global__ void __launch_bounds__(256) diffsr_score_bwd_kernel(const float* __restrict__ phi, const float* __restrict__ flat,
                                                               int D, int S, int B, const float* __restrict__ dscore,
                                                               float* __restrict__ dflat, float* __restrict__ dphi) {
  extern __shared__ float sm[];  // [S] dscore row, [D] phi row
  float* ds = sm;
  float* ph = sm + S;
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < S; i += 256) ds[i] = dscore[(size_t)b * S + i];
  for (int i = tid; i < D; i += 256) ph[i] = phi[(size_t)b * D + i];
  __syncthreads();
  const float* f = flat + (size_t)b * D * S;
  float* df = dflat + (size_t)b * D * S;
  const int total = D * S;
  int d = tid / S, si = tid - d * S;
  const int dd = 256 / S, dsi = 256 - dd * S;  // advancing an element index by 256 moves (d, s) by (dd, dsi) with a carry
  for (int e = tid; e < total; e += 256) {
    df[e] = ph[d] * ds[si];
    d += dd;
    si += dsi;
    if (si >= S) { si -= S; ++d; }
  }
  for (int dr = tid; dr < D; dr += 256) {
    const float* fr = f + (size_t)dr * S;
    float acc = 0.f;
    for (int k = 0; k < S; ++k) acc = fmaf(ds[k], fr[k], acc);
    dphi[(size_t)b * D + dr] = acc;
  }
}

__global__ void __launch_bounds__(256) sum_scaled_kernel(const float* __restrict__ x, int n, float scale,
                                                         float* __restrict__ out) {
  __shared__ float scratch[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += x[i];
  s = block_sum<256>(s, scratch);
  if (threadIdx.x == 0) out[0] = s * scale;
}

}  // namespace

void launch_speder_rows(const float* zphi, const float* zmu, int D, int B, const float* theta_w, const float* theta_b,
                        const float* u, float* diag, float* rpred, float* c, cudaStream_t s) {
  speder_rows_kernel<<<B, 256, 0, s>>>(zphi, zmu, D, B, theta_w, theta_b, u, diag, rpred, c);
  RLREP_LAUNCHED("speder_rows", s);
}
void launch_speder_finalize(const float* diag, const float* c, const float* rpred, const float* reward, int ld_r, int B,
                            float* drp, float* metrics, cudaStream_t s) {
  speder_finalize_kernel<<<1, 256, 0, s>>>(diag, c, rpred, reward, ld_r, B, drp, metrics);
  RLREP_LAUNCHED("speder_finalize", s);
}
void launch_speder_grad(const float* zphi, const float* zmu, int D, int B, const float* drp, const float* theta_w,
                        const float* c, const float* u, const float* w, float* dzphi, float* dzmu, cudaStream_t s) {
  const size_t total = (size_t)2 * B * D;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)kNumSMs * 8);
  speder_grad_kernel<<<blocks, 256, 0, s>>>(zphi, zmu, D, B, drp, theta_w, c, u, w, dzphi, dzmu);
  RLREP_LAUNCHED("speder_grad", s);
}

void launch_diffsr_perturb(const float* next_state, int ld_rec, const float* noise, const long long* level,
                           const float* alphabars, int n_levels, float sigma, int B, int S, float* xin, int ld_x,
                           float* target, float* coef, cudaStream_t s) {
  RLREP_CHECK(ld_x > S, "perturbed input needs room for the alpha-bar column");
  diffsr_perturb_kernel<<<std::min(ceil_div(B * ld_x, 256), kNumSMs * 8), 256, 0, s>>>(
      next_state, ld_rec, noise, level, alphabars, n_levels, sigma, B, S, xin, ld_x, target, coef);
  RLREP_LAUNCHED("diffsr_perturb", s);
}
void launch_diffsr_score(const float* phi, const float* flat, int D, int S, const float* target, const float* coef, int B,
                         float* dscore, float* loss_rows, cudaStream_t s) {
  const int s_tile = S < 256 ? S : 256;
  const int G = 256 / s_tile;
  const size_t smem = ((size_t)G * s_tile + 33) * sizeof(float);
  diffsr_score_kernel<<<B, 256, smem, s>>>(phi, flat, D, S, target, coef, B, dscore, loss_rows);
  RLREP_LAUNCHED("diffsr_score", s);
}
void launch_diffsr_score_bwd(const float* phi, const float* flat, int D, int S, int B, const float* dscore, float* dflat,
                             float* dphi, cudaStream_t s) {
  const size_t smem = (size_t)(S + D) * sizeof(float);
  RLREP_CHECK(smem <= 48 * 1024, "score / feature widths too large for the score-gradient kernel");
  diffsr_score_bwd_kernel<<<B, 256, smem, s>>>(phi, flat, D, S, B, dscore, dflat, dphi);
  RLREP_LAUNCHED("diffsr_score_bwd", s);
}
void launch_sum_scaled(const float* x, int n, float scale, float* out, cudaStream_t s) {
  sum_scaled_kernel<<<1, 256, 0, s>>>(x, n, scale, out);
  RLREP_LAUNCHED("sum_scaled", s);
}

}  // namespace rlrep
