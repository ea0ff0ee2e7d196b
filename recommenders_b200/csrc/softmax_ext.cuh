// softmax_ext.cuh -- operand preparation for the remaining tfrs.tasks.Retrieval loss options inside the tensor-core
// softmax kernels: remove_accidental_hits (tasks/retrieval.py:194-200, layers/loss.py:114-147) needs the
// candidate ids next to the accumulators, score_mask (retrieval.py:202-203) a keep-bit per (query, candidate).
#pragma once
#include <stdint.h>
#include "common.cuh"

namespace tfrs {
namespace tc {

// int64 ids -> two int32 planes (the epilogue compares the low words and looks at the high word only on a match)
static __global__ void __launch_bounds__(256)
sx_ids_split_kernel(const long long* __restrict__ ids, long long C, long long npad, int* __restrict__ lo, int* __restrict__ hi) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= npad) return;
  const long long v = i < C ? ids[i] : 0;
  lo[i] = (int)(unsigned int)(v & 0xffffffffll);
  hi[i] = (int)(v >> 32);
}

// bits[row][w] bit t = (mask[row*C + 32w + t] != 0) for row < B, column < C; everything else 0.  mask: one byte per entry.
static __global__ void __launch_bounds__(256)
sx_mask_pack_kernel(const unsigned char* __restrict__ mask, long long B, long long C, long long rows_pad, int words,
                    uint32_t* __restrict__ bits) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= rows_pad * words) return;
  const long long row = e / words; const int w = (int)(e - row * words);
  uint32_t v = 0;
  if (row < B) {
    const unsigned char* src = mask + row * C + (long long)w * 32;
    const long long left = C - (long long)w * 32;
#pragma unroll 8
    for (int t = 0; t < 32; ++t)
      if (t < left && src[t] != 0) v |= 1u << t;
  }
  bits[e] = v;
}

// the transposed bit matrix for the dc launch of the backward pass (rows = candidates, bit columns = queries):
// bitsT[c][w] bit t = (mask[(32w + t)*C + c] != 0).  Threads of a warp take consecutive candidates: coalesced byte reads.
static __global__ void __launch_bounds__(256)
sx_mask_pack_t_kernel(const unsigned char* __restrict__ mask, long long B, long long C, long long cand_pad, int words,
                      uint32_t* __restrict__ bits_t) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  const int w = blockIdx.y;
  if (c >= cand_pad) return;
  uint32_t v = 0;
  if (c < C) {
#pragma unroll 8
    for (int t = 0; t < 32; ++t) {
      const long long qrow = (long long)w * 32 + t;
      if (qrow < B && mask[qrow * C + c] != 0) v |= 1u << t;
    }
  }
  bits_t[c * words + w] = v;
}

}  // namespace tc
}  // namespace tfrs
