/*
 * rlrep_b200 -- C ABI of the B200-native representation-learning update step.
 *
 * The reference (haotiansun14/rl-rep) has no FFI: its boundary is the duck-typed Python surface
 * `Agent(**kwargs)`, `agent.train(buffer, batch_size) -> dict`, `agent.select_action(state)`,
 * `ReplayBuffer.add/sample` (reference: main.py:71-144, utils/buffer.py:13-48, agent/sac/sac_agent.py:19-188).
 * This header is what a binding for that surface calls; `rlrep_b200/` (Python) is such a binding and
 * INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns int: 0 = ok, non-zero = error; rlrep_last_error() gives the message
 *     (thread-local);
 *   - `stream` arguments are a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - pointers named *_dev are device pointers, *_host are host pointers; the caller owns host buffers,
 *     the library owns every device allocation it makes and frees it in the matching *_destroy;
 *   - one handle = one agent (or one ring) = one stream; handles are independent (thread-compatible,
 *     not thread-safe per handle), so a population runs N handles per GPU;
 *   - there is no CPU fallback anywhere: without a CUDA device every compute call returns an error.
 */
#ifndef RLREP_B200_H_
#define RLREP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RLREP_EXPORT __attribute__((visibility("default")))
#else
#define RLREP_EXPORT
#endif

#define RLREP_ABI_VERSION 1

RLREP_EXPORT int rlrep_abi_version(void);
RLREP_EXPORT const char* rlrep_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Kernel-level entry points (used by the parity tests; the agent handles below call the same code).
 * ---------------------------------------------------------------------------------------------- */

/* Activation / activation-derivative selectors of the fused GEMM epilogue (see csrc/epilogue.cuh). */
enum { RLREP_ACT_NONE = 0, RLREP_ACT_ELU = 1, RLREP_ACT_RELU = 2, RLREP_ACT_TANH = 3, RLREP_ACT_SIN = 4 };
enum {
  RLREP_DACT_NONE = 0,
  RLREP_DACT_ELU_OUT = 1,
  RLREP_DACT_RELU_OUT = 2,
  RLREP_DACT_TANH_OUT = 3,
  RLREP_DACT_COS_PRE = 4
};

typedef struct rlrep_epilogue {
  const float* bias_dev;  /* [N] or NULL                                   */
  const float* r1_u_dev;  /* rank-1 term u[m] * v[n], both NULL to disable */
  const float* r1_v_dev;
  const float* aux_dev;   /* [M, ld_aux] operand of the activation derivative */
  float* pre_out_dev;     /* optional [M, ld_pre]: pre-activation values      */
  int ld_aux;
  int ld_pre;
  int act;
  int dact;
  int accumulate;         /* C += result */
  float scale;            /* accumulator scale, normally 1 */
} rlrep_epilogue;

/*
 * C[M,N] = epilogue(sum_k A(m,k) B(n,k)).  a_mn / b_mn = 0: operand is K-major (A[m*lda+k]); 1: MN-major
 * (A[k*lda+m]).  This single contraction covers nn.Linear forward (reference utils/util.py:85-96,
 * agent/ctrlsac/ctrlsac_agent.py:72-77), its dgrad and its wgrad, and the contrastive logits
 * phi(s,a) . mu(s')^T (ctrlsac_agent.py:229).
 *   path = 0: tcgen05 TF32 tensor-core kernel (TMA operands: 16-byte aligned bases, ld % 4 == 0)
 *   path = 1: CUDA-core exact-FP32 kernel (any alignment; optional second K segment A2 for cat() inputs)
 * bn / split_k = 0 lets the library choose; ws_dev (split-K partials, ws_floats floats) may be NULL.
 */
RLREP_EXPORT int rlrep_gemm(void* stream, int path, int M, int N, int K, const float* A_dev, int lda, int a_mn,
                            const float* A2_dev, int lda2, int K1, const float* B_dev, int ldb, int b_mn,
                            float* C_dev, int ldc, const rlrep_epilogue* epi, int bn, int split_k, float* ws_dev,
                            size_t ws_floats);

/* Tuning aid: plans the GEMM once (tensor maps encoded once, like the agent handles do), launches it `iters`
 * times back to back and reports the CUDA-event average in milliseconds plus the tile/split actually used. */
RLREP_EXPORT int rlrep_gemm_bench(void* stream, int path, int M, int N, int K, const float* A_dev, int lda, int a_mn,
                                  const float* B_dev, int ldb, int b_mn, float* C_dev, int ldc,
                                  const rlrep_epilogue* epi, int bn, int split_k, float* ws_dev, size_t ws_floats,
                                  int iters, float* ms_out, int* bn_out, int* split_out);

#ifdef __cplusplus
}
#endif
#endif /* RLREP_B200_H_ */
