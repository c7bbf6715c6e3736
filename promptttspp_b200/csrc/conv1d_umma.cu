// tcgen05 (UMMA) implicit-GEMM Conv1d path -- placeholder until the split-fp16 kernel lands.
#include "common.h"

namespace pttspp {

bool conv1d_umma_supported(const pttspp_conv1d_desc&) { return false; }

void conv1d_umma_cl(const pttspp_conv1d_desc&, cudaStream_t) {
  throw Error("conv1d: the tcgen05 path is not available in this build");
}

}  // namespace pttspp
