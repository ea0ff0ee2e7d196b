// common.cuh -- shared helpers for libtfrs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/tfrs_b200.h"

namespace tfrs {

// thread-local error text behind tfrs_last_error()
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define TFRS_CHECK_ARG(cond, ...)                          \
  do {                                                     \
    if (!(cond)) {                                         \
      ::tfrs::set_error(__VA_ARGS__);                      \
      return TFRS_ERR_INVALID_ARG;                         \
    }                                                      \
  } while (0)

#define TFRS_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      ::tfrs::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));   \
      return TFRS_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define TFRS_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    ::tfrs::count_launch();                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) {                                                                    \
      ::tfrs::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return TFRS_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// (score desc, index asc): the total order of tf.math.top_k that the whole path relies on.
__host__ __device__ __forceinline__ bool better(float sa, long long ia, float sb, long long ib) {
  return (sa > sb) || (sa == sb && ia < ib);
}

int sm_count();  // SMs of the CURRENT device (cached per device)

// Hook of the sharded scan (topk_tc.cu <-> comm.cu): all device pointers; thr[q] = L_q - margin_q in the shard's screening
// units (scores scaled by 2^(*exp_corpus + qexp[q])), cut[q] = 2 eps_q.  The hook may raise thr[].
typedef int (*ThrHook)(void* ctx, float* thr, const float* margin, const float* cut, const int* qexp, const int* exp_corpus,
                       long long Q, cudaStream_t st);
int tc_topk_sharded_local(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d, int k,
                          int64_t index_offset, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, void* stream,
                          ThrHook hook, void* hook_ctx);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site, device): function attributes belong to
// a device/context, so a process-wide "done" flag would leave the second GPU of a process at the 48 KB default.
struct DeviceOnce {
  unsigned char done_[64] = {};
  int dev_ = 0;
  bool need() { if (cudaGetDevice(&dev_) != cudaSuccess) dev_ = 0; dev_ &= 63; return __atomic_load_n(&done_[dev_], __ATOMIC_ACQUIRE) == 0; }
  void done() { __atomic_store_n(&done_[dev_], (unsigned char)1, __ATOMIC_RELEASE); }
};
#define TFRS_DYN_SMEM(kernel, bytes)                                                                              \
  do {                                                                                                            \
    static ::tfrs::DeviceOnce once__;                                                                             \
    if (once__.need()) {                                                                                          \
      TFRS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));         \
      /* without the carve-out hint the driver sizes the L1/shared split for ONE block of this kernel, which caps   \
         multi-block-per-SM kernels (the warp-per-query finalize) at one resident block */                          \
      TFRS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,                        \
                                     (int)cudaSharedmemCarveoutMaxShared));                                         \
      once__.done();                                                                                              \
    }                                                                                                             \
  } while (0)

}  // namespace tfrs
