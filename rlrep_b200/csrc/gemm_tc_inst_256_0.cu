// tcgen05 TF32 GEMM variants with BN = 256, K-major A (both B majors); see gemm_tc_kernel.cuh.
#define RLREP_TC_DEVICE_CODE
#include "gemm_tc_kernel.cuh"

namespace rlrep {
namespace tc {
RLREP_TC_DEFINE(256, 0)
}  // namespace tc
}  // namespace rlrep
