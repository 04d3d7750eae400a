// tcgen05 TF32 GEMM for sm_100a: TMA-fed shared-memory ring -> tcgen05.mma (kind::tf32) with the FP32
// accumulator in TMEM -> tcgen05.ld epilogue with fused bias / activation / activation-derivative.
//
// One CTA computes one 128 x BN output tile over a K-slice (split-K along gridDim.z).
//   warp 0     : TMA producer (one elected lane)
//   warp 1     : MMA issuer (one elected lane)
//   warps 2..5 : epilogue (warp w owns TMEM lanes 32*(w&3) ..); warp 2 also owns TMEM alloc/dealloc
//
// Operands stay FP32 in HBM; the tensor maps are typed TFLOAT32 so the TMA unit rounds to TF32 on the way
// into shared memory and the weights are never re-materialised in a second precision.
//
// Both K-major and MN-major operands are supported (see gemm.cuh), so forward, dgrad and wgrad all run
// here without transposed copies of weights or activations:
//   K-major  tile: rows x 32 fp32 (128 B) per k-block, one TMA box, SWIZZLE_128B, SBO = 1024 B;
//                  the 4 UMMA_K=8 slices of a k-block advance the descriptor start address by 32 B.
//   MN-major tile: 32 k-rows x 32 fp32 boxes (4 KB each), one per 32 rows of M/N.  tcgen05 accepts exactly one
//                  layout for MN-major tf32: SWIZZLE_128B_BASE32B (32-byte chunks XOR k-row % 4, 4-row atoms),
//                  which TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = 4096 B between boxes,
//                  SBO = 512 B between 4-k-row groups; the 4 UMMA_K=8 slices advance the start by 1024 B.
#include <cstdint>
#include <mutex>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace rlrep {

namespace {

constexpr int BM = 128;           // UMMA M (cta_group::1)
constexpr int BK = 32;            // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;         // tf32
constexpr int kThreads = 192;     // 6 warps
constexpr int kSmemBudget = 200 * 1024;

__host__ __device__ constexpr int stage_bytes(int bn) { return (BM + bn) * BK * 4; }
__host__ __device__ constexpr int num_stages(int bn) {
  return kSmemBudget / stage_bytes(bn) > 8 ? 8 : kSmemBudget / stage_bytes(bn);
}
__host__ __device__ constexpr int smem_bytes(int bn) { return num_stages(bn) * stage_bytes(bn) + 1024 + 256; }

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 float* __restrict__ C, int ldc, int M, int N, int K, int kb_per_split,
                 float* __restrict__ ws, const Epilogue epi) {
  constexpr int STAGES = num_stages(BN);
  constexpr int A_BYTES = BM * BK * 4;
  constexpr int B_BYTES = BN * BK * 4;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN must be a power of two in [32,256]");

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms must sit on 1024-byte boundaries.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int nkb_total = (K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        const int k0 = (kb_begin + it) * BK;
        uint8_t* a_dst = sA + s * A_BYTES;
        uint8_t* b_dst = sB + s * B_BYTES;
        if (!A_MN) {
          ptx::tma_load_2d(a_dst, &tmA, &full_bar[s], k0, m0);
        } else {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) ptx::tma_load_2d(a_dst + j * 4096, &tmA, &full_bar[s], m0 + 32 * j, k0);
        }
        if (!B_MN) {
          ptx::tma_load_2d(b_dst, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) ptx::tma_load_2d(b_dst + j * 4096, &tmB, &full_bar[s], n0 + 32 * j, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_tf32(BM, BN, A_MN, B_MN);
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ptx::mbar_wait(&full_bar[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(sA + s * A_BYTES);
        const uint32_t b_addr = ptx::smem_u32(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t adesc = A_MN ? ptx::make_smem_desc(a_addr + k * 1024, 4096, 512, 1)
                                      : ptx::make_smem_desc(a_addr + k * UMMA_K * 4, 16, 1024, 2);
          const uint64_t bdesc = B_MN ? ptx::make_smem_desc(b_addr + k * 1024, 4096, 512, 1)
                                      : ptx::make_smem_desc(b_addr + k * UMMA_K * 4, 16, 1024, 2);
          ptx::mma_tf32_ss(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        ptx::mma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
      }
      ptx::mma_commit(accum_bar);  // accumulator complete
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = m0 + 32 * q + lane;
    ptx::mbar_wait(accum_bar, 0);
    ptx::tc_fence_after_sync();
    const bool split = gridDim.z > 1;
    float* out = split ? ws + (size_t)blockIdx.z * M * N : C;
    const int ldo = split ? N : ldc;
    const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(32 * q) << 16) + c * 32, r);
      ptx::tmem_ld_wait();
      if (row < M) {
        const int nb = n0 + c * 32;
        float* orow = out + (size_t)row * ldo;
        if (split) {
          if (vec_ok && nb + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(orow + nb + j) =
                  make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                              __uint_as_float(r[j + 3]));
          } else {
            for (int j = 0; j < 32; ++j)
              if (nb + j < N) orow[nb + j] = __uint_as_float(r[j]);
          }
        } else {
          if (vec_ok && nb + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v;
              v.x = epilogue_apply(epi, __uint_as_float(r[j]), row, nb + j, orow + nb + j);
              v.y = epilogue_apply(epi, __uint_as_float(r[j + 1]), row, nb + j + 1, orow + nb + j + 1);
              v.z = epilogue_apply(epi, __uint_as_float(r[j + 2]), row, nb + j + 2, orow + nb + j + 2);
              v.w = epilogue_apply(epi, __uint_as_float(r[j + 3]), row, nb + j + 3, orow + nb + j + 3);
              *reinterpret_cast<float4*>(orow + nb + j) = v;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (nb + j < N) orow[nb + j] = epilogue_apply(epi, __uint_as_float(r[j]), row, nb + j, orow + nb + j);
          }
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// Deterministic split-K reduction (fixed summation order) + epilogue.
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, float* __restrict__ C,
                                     int ldc, const Epilogue epi) {
  const size_t total = (size_t)M * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += ws[(size_t)z * total + i];
    const int m = static_cast<int>(i / N);
    const int n = static_cast<int>(i - (size_t)m * N);
    float* cp = C + (size_t)m * ldc + n;
    *cp = epilogue_apply(epi, acc, m, n, cp);
  }
}

// ------------------------------------------------------------------ host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  RLREP_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

// Tensor map over a row-major fp32 matrix [rows, cols] with leading dimension ld; box = box_rows x 32 columns.
CUtensorMap make_map(const float* base, int rows, int cols, int ld, int box_rows, bool mn_major) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), gdim, gstride, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLREP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  return m;
}

template <int BN, bool A_MN, bool B_MN>
void launch_variant(const TcGemmPlan& p, cudaStream_t stream) {
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    RLREP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes(BN)));
    attr_set = true;
  }
  const GemmArgs& a = p.args;
  dim3 grid(ceil_div(a.N, BN), ceil_div(a.M, BM), p.split_k);
  kern<<<grid, kThreads, smem_bytes(BN), stream>>>(p.tmA, p.tmB, a.C, a.ldc, a.M, a.N, a.K, p.kb_per_split, p.ws,
                                                   a.epi);
  RLREP_LAUNCHED("gemm_tf32", stream);
}

template <int BN>
void launch_bn(const TcGemmPlan& p, cudaStream_t stream) {
  const bool a = p.args.a_mn, b = p.args.b_mn;
  if (!a && !b) launch_variant<BN, false, false>(p, stream);
  else if (!a && b) launch_variant<BN, false, true>(p, stream);
  else if (a && !b) launch_variant<BN, true, false>(p, stream);
  else launch_variant<BN, true, true>(p, stream);
}

}  // namespace

bool tc_eligible(const GemmArgs& a) {
  auto ok = [](const float* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0 && ld > 0; };
  return a.A2 == nullptr && ok(a.A, a.lda) && ok(a.B, a.ldb) && a.M > 0 && a.N > 0 && a.K > 0;
}

TcGemmPlan make_tc_plan(const GemmArgs& a, int bn, int split_k, float* ws, size_t ws_floats) {
  RLREP_CHECK(tc_eligible(a), "operands violate TMA alignment (16-byte base, ld % 4 == 0)");
  TcGemmPlan p;
  p.args = a;
  if (bn == 0) bn = a.N >= 128 ? 128 : (a.N > 32 ? 64 : 32);
  RLREP_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 256, "bn must be 32/64/128/256");
  p.bn = bn;
  const int nkb = ceil_div(a.K, BK);
  const int base_ctas = ceil_div(a.M, BM) * ceil_div(a.N, bn);
  if (split_k == 0) {
    split_k = 1;
    if (ws != nullptr && base_ctas < kNumSMs) {
      split_k = kNumSMs / base_ctas;
      if (split_k > nkb / 2) split_k = nkb / 2;  // keep >= 2 k-blocks per split
      if (split_k > 16) split_k = 16;
      if (split_k < 1) split_k = 1;
    }
  }
  if (split_k > nkb) split_k = nkb;
  if (ws == nullptr) split_k = 1;
  while (split_k > 1 && (size_t)split_k * a.M * a.N > ws_floats) --split_k;
  p.kb_per_split = ceil_div(nkb, split_k);
  p.split_k = ceil_div(nkb, p.kb_per_split);  // no empty splits
  p.ws = ws;
  // K-major operand: matrix [rows = M|N, cols = K]; MN-major: matrix [rows = K, cols = M|N].
  p.tmA = a.a_mn ? make_map(a.A, a.K, a.M, a.lda, 32, true) : make_map(a.A, a.M, a.K, a.lda, BM, false);
  p.tmB = a.b_mn ? make_map(a.B, a.K, a.N, a.ldb, 32, true) : make_map(a.B, a.N, a.K, a.ldb, bn, false);
  return p;
}

void launch_tc(const TcGemmPlan& p, cudaStream_t stream) {
  switch (p.bn) {
    case 32: launch_bn<32>(p, stream); break;
    case 64: launch_bn<64>(p, stream); break;
    case 128: launch_bn<128>(p, stream); break;
    case 256: launch_bn<256>(p, stream); break;
    default: throw Error("bad bn");
  }
  if (p.split_k > 1) {
    const GemmArgs& a = p.args;
    const size_t total = (size_t)a.M * a.N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p.ws, p.split_k, a.M, a.N, a.C, a.ldc, a.epi);
    RLREP_LAUNCHED("splitk_reduce", stream);
  }
}

}  // namespace rlrep
