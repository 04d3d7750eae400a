// tcgen05 TF32 GEMM variants with BN = 32, MN-major A (both B majors); see gemm_tc_kernel.cuh.
#define RLREP_TC_DEVICE_CODE
#include "gemm_tc_kernel.cuh"

namespace rlrep {
namespace tc {
RLREP_TC_DEFINE(32, 1)
}  // namespace tc
}  // namespace rlrep
