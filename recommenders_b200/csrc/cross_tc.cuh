// cross_tc.cuh -- entry points of the tensor-core Cross backward GEMMs (cross_tc_bwd.cu), called from cross.cu.
#pragma once
#include "common.cuh"

namespace tfrs {
namespace tc {

size_t cross_tc_bwd_gemm_workspace(long long B, int D);
// dx = gp . W^T + diag * gp + dout (if dx != NULL);  dW = x^T . gp (if dW != NULL).  gp is dense [B,D]; x / dout / dx
// have row stride ld.
int cross_tc_bwd_gemms(const float* x, const float* W, const float* gp, const float* dout, long long B, int D, long long ld,
                       float diag, float* dx, float* dW, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace tc
}  // namespace tfrs
