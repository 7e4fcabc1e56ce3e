// tcgen05 / TMA GEMM for X * W^T (kind::tf32).  Placeholder until the tensor-core path lands.
#include "kernels.cuh"

namespace bigcn {
bool xw_tc_available() { return false; }
int xw_tc(const float*, int64_t, int64_t, const float*, int, float*, int64_t, int, cudaStream_t) {
  set_error("xw: the tcgen05 GEMM modes are not built in this version; use BIGCN_GEMM_FP32");
  return 1;
}
}  // namespace bigcn
