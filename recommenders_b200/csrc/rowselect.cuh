// rowselect.cuh -- per-query exact top-K selection over an arbitrary candidate provider.
//
// One CTA per query row.  Candidates stream through a threshold filter (the current K-th best)
// into a shared-memory buffer; when the buffer fills it is bitonic-sorted on the total order
// (score desc, index asc) and cut back to K.  The result is deterministic (indices are unique, so
// the total order has no ties) and equals tf.math.top_k's contract.
//
// Provider concept:
//   __device__ void   begin(int row, void* smem_extra)       -- all threads; followed by __syncthreads
//   __device__ long long count(int row)
//   __device__ void   get(int row, long long t, float& s, long long& i)
#pragma once
#include "common.cuh"
#include <limits.h>

namespace tfrs {

constexpr int RS_THREADS = 256;

__device__ __forceinline__ void bitonic_sort_desc(float* bs, long long* bi, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool desc = ((lo & size) == 0);
        float sl = bs[lo], sh = bs[hi];
        long long il = bi[lo], ih = bi[hi];
        bool hi_better = better(sh, ih, sl, il);
        bool lo_better = better(sl, il, sh, ih);
        if (desc ? hi_better : lo_better) { bs[lo] = sh; bs[hi] = sl; bi[lo] = ih; bi[hi] = il; }
      }
      __syncthreads();
    }
  }
}

static inline int pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int rowselect_cap(int k) { int c = pow2_ceil(2 * k); return c < 1024 ? 1024 : c; }
static inline size_t rowselect_smem(int cap, size_t extra) { return (size_t)cap * 12 + extra; }

// out_s / out_i: [rows, out_ld]; the first min(k, count(row)) entries of each row are written.
template <class Provider>
__global__ void __launch_bounds__(RS_THREADS)
row_topk_kernel(Provider prov, int k, int cap, float* __restrict__ out_s, long long* __restrict__ out_i,
                int out_ld) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  long long* bi = reinterpret_cast<long long*>(rs_smem);
  float* bs = reinterpret_cast<float*>(rs_smem + (size_t)cap * 8);
  void* extra = rs_smem + (size_t)cap * 12;
  __shared__ int cnt_sh;
  __shared__ float thr_s_sh;
  __shared__ long long thr_i_sh;
  __shared__ int full_sh;

  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) { cnt_sh = 0; full_sh = 0; thr_s_sh = 0.f; thr_i_sh = 0; }
  prov.begin(row, extra);
  __syncthreads();

  const long long total = prov.count(row);
  const int min_room = cap / 4;
  long long pos = 0;
  while (pos < total) {
    __syncthreads();  // (a) cnt/threshold written by the previous iteration are visible
    const int cnt0 = cnt_sh;
    const bool full = full_sh != 0;
    const float ts = thr_s_sh;
    const long long ti = thr_i_sh;
    __syncthreads();  // (b) everybody has read them before anybody appends
    const int slab = (int)min((long long)(cap - cnt0), total - pos);
    for (int t = tid; t < slab; t += RS_THREADS) {
      float s; long long i;
      prov.get(row, pos + t, s, i);
      if (!full || better(s, i, ts, ti)) {
        int p = atomicAdd(&cnt_sh, 1);
        bs[p] = s; bi[p] = i;
      }
    }
    __syncthreads();  // (c)
    pos += slab;
    const int cnt = cnt_sh;
    if (pos < total && cap - cnt >= min_room) continue;  // keep filling; a stale threshold is only looser
    // ---- compaction: sort, keep the best k
    int P = 2; while (P < cnt) P <<= 1;
    for (int t = cnt + tid; t < P; t += RS_THREADS) { bs[t] = -INFINITY; bi[t] = LLONG_MAX; }
    __syncthreads();
    if (cnt > 1) bitonic_sort_desc(bs, bi, P);
    if (tid == 0) {
      int keep = cnt < k ? cnt : k;
      cnt_sh = keep;
      if (keep == k) { full_sh = 1; thr_s_sh = bs[k - 1]; thr_i_sh = bi[k - 1]; }
    }
  }
  __syncthreads();
  const int n_out = cnt_sh;
  for (int t = tid; t < n_out; t += RS_THREADS) {
    out_s[(long long)row * out_ld + t] = bs[t];
    out_i[(long long)row * out_ld + t] = bi[t];
  }
}

}  // namespace tfrs
