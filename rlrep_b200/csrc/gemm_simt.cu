// CUDA-core FP32 GEMM (FFMA, exact fp32 products and accumulation) with the shared epilogue.
// Used for the layers tensor cores cannot help: K = 17/23-wide input layers (optionally reading the
// concatenated input from two buffers), N = 1 heads, the 12-wide actor head, and for the strict-fp32
// mode of the library (bit-for-bit IEEE fp32 multiply-add, like torch with allow_tf32 = False).
//
// 64x64 output tile per 256-thread CTA, BK = 16, 4x4 register micro-tile, operands staged through
// shared memory with the k index outermost so inner-loop reads are conflict-free float4 broadcasts.
#include "common.cuh"
#include "gemm.cuh"

namespace rlrep {

namespace {

constexpr int TBM = 64, TBN = 64, TBK = 16;

struct SimtParams {
  int M, N, K, K1;
  const float* A;
  long long sam, sak;  // A(m,k) = A[m*sam + k*sak]
  const float* A2;
  long long sam2;      // second K segment (K-major only): A2[m*sam2 + (k-K1)]
  const float* B;
  long long sbn, sbk;  // B(n,k) = B[n*sbn + k*sbk]
  float* C;
  int ldc;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p, const Epilogue epi) {
  __shared__ __align__(16) float As[TBK][TBM + 4];
  __shared__ __align__(16) float Bs[TBK][TBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4 outputs
  float acc[4][4] = {};

  const bool a_kmajor = (p.sak == 1);
  const bool b_kmajor = (p.sbk == 1);

  for (int k0 = 0; k0 < p.K; k0 += TBK) {
    // ---- stage A tile [TBM x TBK]
#pragma unroll
    for (int i = 0; i < (TBM * TBK) / 256; ++i) {
      const int idx = tid + i * 256;
      int mm, kk;
      if (a_kmajor) { kk = idx % TBK; mm = idx / TBK; } else { mm = idx % TBM; kk = idx / TBM; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < p.M && k < p.K) {
        if (p.A2 != nullptr && k >= p.K1) v = __ldg(p.A2 + (long long)m * p.sam2 + (k - p.K1));
        else v = __ldg(p.A + (long long)m * p.sam + (long long)k * p.sak);
      }
      As[kk][mm] = v;
    }
    // ---- stage B tile [TBN x TBK]
#pragma unroll
    for (int i = 0; i < (TBN * TBK) / 256; ++i) {
      const int idx = tid + i * 256;
      int nn, kk;
      if (b_kmajor) { kk = idx % TBK; nn = idx / TBK; } else { nn = idx % TBN; kk = idx / TBN; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < p.N && k < p.K) v = __ldg(p.B + (long long)n * p.sbn + (long long)k * p.sbk);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float* cp = p.C + (size_t)m * p.ldc + n;
      *cp = epilogue_apply(epi, acc[i][j], m, n, cp);
    }
  }
}

}  // namespace

void launch_simt(const GemmArgs& a, cudaStream_t stream) {
  RLREP_CHECK(a.M > 0 && a.N > 0 && a.K > 0, "empty GEMM");
  RLREP_CHECK(a.A2 == nullptr || !a.a_mn, "two-segment A requires a K-major A");
  SimtParams p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.K1 = a.A2 ? a.K1 : a.K;
  p.A = a.A;
  p.sam = a.a_mn ? 1 : a.lda;
  p.sak = a.a_mn ? a.lda : 1;
  p.A2 = a.A2;
  p.sam2 = a.lda2;
  p.B = a.B;
  p.sbn = a.b_mn ? 1 : a.ldb;
  p.sbk = a.b_mn ? a.ldb : 1;
  p.C = a.C;
  p.ldc = a.ldc;
  dim3 grid(ceil_div(a.N, TBN), ceil_div(a.M, TBM));
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(p, a.epi);
  RLREP_LAUNCHED("gemm_simt", stream);
}

}  // namespace rlrep
