// Tensor-core GEMM modes (tcgen05).  Placeholder until the tcgen05 path lands: the modes report
// CAPDEC_ERR_UNSUPPORTED instead of silently falling back to the fp32 kernel.
#include "handle.cuh"

namespace capdec {

int gemm_tc(int precision, const GemmArgs&, int, cudaStream_t) {
  set_error("precision mode %d is not built into this libcapdec", precision);
  return CAPDEC_ERR_UNSUPPORTED;
}

int gemm_tc_prepare(capdec_handle*, cudaStream_t) { return CAPDEC_OK; }

}  // namespace capdec
