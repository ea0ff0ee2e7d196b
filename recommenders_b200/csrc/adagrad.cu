// adagrad.cu -- K4: deterministic sparse Adagrad on the rows touched by a batch.
// Replaces optimizer.apply_gradients(IndexedSlices) (models/base.py:77-78, Adagrad per README.md:84).
//   1. keys = (id << 24 | position) sorted ascending (bitonic, shared-memory tiles + global strides)
//   2. one warp per run of equal ids: duplicate gradient rows are summed in order of occurrence
//      (lanes = columns, no float atomics), then  acc += g*g ; var -= lr*g/sqrt(acc+eps)  (or the
//      legacy sqrt(acc)+eps form).  Every step is a single IEEE fp32 op (no FMA contraction) so the
//      result is bit-identical to oracle/tfrs_oracle.c::orc_sparse_adagrad.
// HBM-bound: algorithmic bytes = unique_rows * d * 4 * 4 (table r/w + accum r/w) + n*d*4 (grads).
#include "common.cuh"

namespace tfrs {

constexpr int AG_TILE = 8192;       // keys per CTA tile (64 KB of shared memory)
constexpr int AG_THREADS = 1024;
constexpr unsigned long long AG_INVALID = ~0ull;

template <typename IdT>
__global__ void ag_build_keys(const IdT* __restrict__ ids, long long n, long long rows, long long P,
                              unsigned long long* __restrict__ keys) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  unsigned long long k = AG_INVALID;
  if (i < n) {
    long long r = (long long)ids[i];
    if (r >= 0 && r < rows) k = ((unsigned long long)r << 24) | (unsigned long long)i;
  }
  keys[i] = k;
}

__device__ __forceinline__ void ag_cmpswap(unsigned long long& a, unsigned long long& b, bool asc) {
  if ((a > b) == asc) { unsigned long long t = a; a = b; b = t; }
}

// Sorts (first_size == 2) or merges (first_size == size) one AG_TILE chunk in shared memory.
// Processes bitonic stages size = first_size..last_size with strides min(size/2, TILE/2)..1.
__global__ void __launch_bounds__(AG_THREADS)
ag_bitonic_local(unsigned long long* __restrict__ keys, long long P, long long first_size, long long last_size) {
  extern __shared__ unsigned long long sk[];
  const long long base = (long long)blockIdx.x * AG_TILE;
  const int tile = (int)min((long long)AG_TILE, P - base);
  for (int t = threadIdx.x; t < tile; t += AG_THREADS) sk[t] = keys[base + t];
  __syncthreads();
  for (long long size = first_size; size <= last_size; size <<= 1) {
    int s0 = (int)min(size >> 1, (long long)(tile >> 1));
    for (int stride = s0; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (tile >> 1); t += AG_THREADS) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool asc = (((base + lo) & size) == 0);
        unsigned long long a = sk[lo], b = sk[hi];
        ag_cmpswap(a, b, asc);
        sk[lo] = a; sk[hi] = b;
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < tile; t += AG_THREADS) keys[base + t] = sk[t];
}

// n <= AG_SINGLE: keys are built and fully sorted by ONE CTA in shared memory (one launch instead of build + tile sorts
// + global merge stages); the batch sizes of the training step (cfg3: 16384 ids) take this path.
constexpr int AG_SINGLE = 16384;   // 128 KB of 64-bit keys
template <typename IdT>
__global__ void __launch_bounds__(AG_THREADS)
ag_sort_single(const IdT* __restrict__ ids, long long n, long long rows, int P, unsigned long long* __restrict__ keys) {
  extern __shared__ unsigned long long sk[];
  for (int i = threadIdx.x; i < P; i += AG_THREADS) {
    unsigned long long k = AG_INVALID;
    if (i < n) {
      const long long r = (long long)ids[i];
      if (r >= 0 && r < rows) k = ((unsigned long long)r << 24) | (unsigned long long)i;
    }
    sk[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += AG_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool asc = ((lo & size) == 0);
        unsigned long long a = sk[lo], b = sk[hi];
        ag_cmpswap(a, b, asc);
        sk[lo] = a; sk[hi] = b;
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < P; i += AG_THREADS) keys[i] = sk[i];
}

__global__ void ag_bitonic_global(unsigned long long* __restrict__ keys, long long P, long long size, long long stride) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (P >> 1)) return;
  long long lo = 2 * t - (t & (stride - 1));
  long long hi = lo + stride;
  bool asc = ((lo & size) == 0);
  unsigned long long a = keys[lo], b = keys[hi];
  ag_cmpswap(a, b, asc);
  keys[lo] = a; keys[hi] = b;
}

__global__ void __launch_bounds__(256)
ag_apply(const unsigned long long* __restrict__ keys, long long n, const float* __restrict__ grad, int d,
         float* __restrict__ table, float* __restrict__ accum, float lr, float eps, int eps_inside) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per sorted slot
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  if (key == AG_INVALID) return;
  const unsigned long long id = key >> 24;
  if (i > 0 && (keys[i - 1] >> 24) == id) return;  // not the head of its run
  long long end = i + 1;
  while (end < n && (keys[end] >> 24) == id) ++end;
  float* trow = table + (long long)id * d;
  float* arow = accum + (long long)id * d;
  for (int c = lane; c < d; c += 32) {
    float g = grad[(long long)(key & 0xFFFFFFull) * d + c];
    for (long long j = i + 1; j < end; ++j) g = __fadd_rn(g, grad[(long long)(keys[j] & 0xFFFFFFull) * d + c]);
    float a = __fadd_rn(arow[c], __fmul_rn(g, g));
    arow[c] = a;
    float den = eps_inside ? __fsqrt_rn(__fadd_rn(a, eps)) : __fadd_rn(__fsqrt_rn(a), eps);
    trow[c] = __fsub_rn(trow[c], __fdiv_rn(__fmul_rn(lr, g), den));
  }
}

static long long ag_pow2(long long n) { long long p = 1; while (p < n) p <<= 1; return p; }

}  // namespace tfrs
using namespace tfrs;

extern "C" size_t tfrs_sparse_adagrad_workspace_bytes(int64_t n, int d) {
  (void)d;
  return (size_t)ag_pow2(n > 2 ? n : 2) * 8 + 256;
}

extern "C" int tfrs_sparse_adagrad_f32(float* table, float* accum, int64_t rows, int d, const void* ids,
                                       int ids_dtype, int64_t n, const float* grad_rows, float lr, float eps,
                                       int eps_inside_sqrt, void* ws, size_t ws_bytes, void* stream) {
  TFRS_CHECK_ARG(table && accum && rows > 0 && d > 0, "sparse_adagrad: bad table");
  TFRS_CHECK_ARG(ids_dtype == TFRS_I32 || ids_dtype == TFRS_I64, "sparse_adagrad: ids_dtype must be I32 or I64");
  TFRS_CHECK_ARG(n >= 0 && n < (1ll << 24), "sparse_adagrad: n=%lld must be < 2^24", (long long)n);
  TFRS_CHECK_ARG(rows < (1ll << 40), "sparse_adagrad: rows must be < 2^40");
  if (n == 0) return TFRS_OK;
  TFRS_CHECK_ARG(ids && grad_rows, "sparse_adagrad: NULL ids/grad");
  const long long P = ag_pow2(n > 2 ? n : 2);
  if (!ws || ws_bytes < (size_t)P * 8) { set_error("sparse_adagrad: workspace too small"); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* keys = (unsigned long long*)ws;
  if (P <= AG_SINGLE) {
    if (ids_dtype == TFRS_I32) {
      TFRS_DYN_SMEM(ag_sort_single<int32_t>, AG_SINGLE * 8);
      ag_sort_single<int32_t><<<1, AG_THREADS, (size_t)P * 8, st>>>((const int32_t*)ids, n, rows, (int)P, keys);
    } else {
      TFRS_DYN_SMEM(ag_sort_single<int64_t>, AG_SINGLE * 8);
      ag_sort_single<int64_t><<<1, AG_THREADS, (size_t)P * 8, st>>>((const int64_t*)ids, n, rows, (int)P, keys);
    }
    TFRS_LAUNCH_CHECK();
    ag_apply<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(keys, n, grad_rows, d, table, accum, lr, eps, eps_inside_sqrt);
    TFRS_LAUNCH_CHECK();
    return TFRS_OK;
  }
  unsigned kb = (unsigned)ceil_div(P, 256);
  if (ids_dtype == TFRS_I32) ag_build_keys<int32_t><<<kb, 256, 0, st>>>((const int32_t*)ids, n, rows, P, keys);
  else ag_build_keys<int64_t><<<kb, 256, 0, st>>>((const int64_t*)ids, n, rows, P, keys);
  TFRS_LAUNCH_CHECK();
  TFRS_DYN_SMEM(ag_bitonic_local, AG_TILE * 8);
  const unsigned tiles = (unsigned)ceil_div(P, AG_TILE);
  const long long local_max = P < AG_TILE ? P : AG_TILE;
  ag_bitonic_local<<<tiles, AG_THREADS, AG_TILE * 8, st>>>(keys, P, 2, local_max);
  TFRS_LAUNCH_CHECK();
  for (long long size = 2ll * AG_TILE; size <= P; size <<= 1) {
    for (long long stride = size >> 1; stride >= AG_TILE; stride >>= 1) {
      ag_bitonic_global<<<(unsigned)ceil_div(P >> 1, 256), 256, 0, st>>>(keys, P, size, stride);
      TFRS_LAUNCH_CHECK();
    }
    ag_bitonic_local<<<tiles, AG_THREADS, AG_TILE * 8, st>>>(keys, P, size, size);
    TFRS_LAUNCH_CHECK();
  }
  ag_apply<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(keys, n, grad_rows, d, table, accum, lr, eps, eps_inside_sqrt);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}
