// extern "C" surface of librlrep_b200.so (declared in include/rlrep_b200.h).
#include <string>

#include "common.cuh"
#include "gemm.cuh"
#include "rlrep_b200.h"

namespace rlrep {
namespace {
thread_local std::string g_last_error;
}
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }

Epilogue to_epilogue(const rlrep_epilogue* e) {
  Epilogue o;
  if (e == nullptr) return o;
  o.bias = e->bias_dev;
  o.r1_u = e->r1_u_dev;
  o.r1_v = e->r1_v_dev;
  o.aux = e->aux_dev;
  o.pre_out = e->pre_out_dev;
  o.ld_aux = e->ld_aux;
  o.ld_pre = e->ld_pre;
  o.act = e->act;
  o.dact = e->dact;
  o.accumulate = e->accumulate;
  o.scale = e->scale;
  return o;
}
}  // namespace rlrep

using namespace rlrep;

extern "C" {

int rlrep_abi_version(void) { return RLREP_ABI_VERSION; }
const char* rlrep_last_error(void) { return get_last_error(); }

int rlrep_gemm(void* stream, int path, int M, int N, int K, const float* A, int lda, int a_mn, const float* A2,
               int lda2, int K1, const float* B, int ldb, int b_mn, float* C, int ldc, const rlrep_epilogue* epi,
               int bn, int split_k, float* ws, size_t ws_floats) {
  RLREP_API_BEGIN
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_mn = a_mn != 0;
  g.A2 = A2; g.lda2 = lda2; g.K1 = K1;
  g.B = B; g.ldb = ldb; g.b_mn = b_mn != 0;
  g.C = C; g.ldc = ldc;
  g.epi = to_epilogue(epi);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (path == 0) {
    TcGemmPlan p = make_tc_plan(g, bn, split_k, ws, ws_floats);
    launch_tc(p, st);
  } else {
    launch_simt(g, st);
  }
  RLREP_API_END
}

// Times `iters` back-to-back launches of one planned GEMM with CUDA events on `stream` (tensor maps encoded
// once, as the agent handles do).  Tuning/benchmark aid; ms_out = average milliseconds per GEMM.
int rlrep_gemm_bench(void* stream, int path, int M, int N, int K, const float* A, int lda, int a_mn, const float* B,
                     int ldb, int b_mn, float* C, int ldc, const rlrep_epilogue* epi, int bn, int split_k, float* ws,
                     size_t ws_floats, int iters, float* ms_out, int* bn_out, int* split_out) {
  RLREP_API_BEGIN
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.A = A; g.lda = lda; g.a_mn = a_mn != 0;
  g.B = B; g.ldb = ldb; g.b_mn = b_mn != 0;
  g.C = C; g.ldc = ldc;
  g.epi = to_epilogue(epi);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  RLREP_CUDA(cudaEventCreate(&e0));
  RLREP_CUDA(cudaEventCreate(&e1));
  if (path == 0) {
    TcGemmPlan p = make_tc_plan(g, bn, split_k, ws, ws_floats);
    if (bn_out) *bn_out = p.bn;
    if (split_out) *split_out = p.split_k;
    for (int i = 0; i < 3; ++i) launch_tc(p, st);
    RLREP_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) launch_tc(p, st);
    RLREP_CUDA(cudaEventRecord(e1, st));
  } else {
    for (int i = 0; i < 3; ++i) launch_simt(g, st);
    RLREP_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) launch_simt(g, st);
    RLREP_CUDA(cudaEventRecord(e1, st));
  }
  RLREP_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  RLREP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  RLREP_API_END
}

}  // extern "C"
