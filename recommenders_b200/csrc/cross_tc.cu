// cross_tc.cu -- K5 on the tensor cores: DCN-v2 cross layer forward (layers/feature_interaction/dcn.py:176-186)
//   out = x0 * (x . W + bias + diag_scale * x) + x        W [D,D] in Keras [in,out] layout
// as ONE tcgen05 GEMM with the whole cross formula in the epilogue.
//
// fp32 parity on fp16 tensor cores: each operand is rescaled by an exact power of two and split into
//   v = hi + lo,  hi = fp16(v), lo = fp16(v - hi)            (|v - hi - lo| <= 2^-22 |v|)
// and the product is accumulated in fp32 (TMEM) as  hi_x*hi_w + lo_x*hi_w + hi_x*lo_w  (the dropped
// lo*lo term is 2^-22 relative), i.e. 3 MMAs per K step -- ~2^-21 relative error, inside the 1e-5 bar.
//
// Layout: x (per call) and W^T (once per weight version) are turned into UMMA SWIZZLE_128B K-major tile
// images, 128 rows x 64 K-elements per 16 KB block, hi block then lo block per K slab (32 KB per slab).
// Kernel: persistent CTAs (1/SM, 640 threads) over (256-row block, 128-column tile) pairs; per K slab a
// bulk-TMA stage brings 2x(hi,lo) A blocks + (hi,lo) of W^T (96 KB); 24 MMAs (2 A blocks x 3 products x 4
// K16 steps) accumulate into a 128-column TMEM buffer per A block (2 buffers -> next tile's MMAs overlap the
// epilogue).  16 epilogue warps read 64 columns of one row each, fetch x0 / x / bias, apply the formula and
// store fp32.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_split.cuh"

namespace tfrs {
namespace tc {

constexpr int CX_THREADS = 640;
constexpr int CX_STAGES = 2;
constexpr int CX_STAGE_BYTES = 6 * 16384;  // A: 2 blocks x (hi, lo); B: (hi, lo)
struct CrossParams {
  const unsigned char* ximg;  // [n_mtiles128][kb][hi|lo][16 KB]
  const unsigned char* wimg;  // [n_ntiles128][kb][hi|lo][16 KB]   (W^T: rows = output column)
  const CxStats* xst; const CxStats* wst;
  const float* x0; const float* x; const float* bias; float diag;
  float* out; float* prod;
  long long B; int D; long long ld;
  int kb, n_mb, n_nt;
  unsigned int* out_amax;     // nullable: max |out| (float bits) accumulated by the epilogue -- the next layer's rescale statistic
};

__global__ void __launch_bounds__(CX_THREADS, 1)
cross_tc_kernel(const CrossParams p) {
  extern __shared__ __align__(1024) unsigned char cx_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cx_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CX_STAGES * CX_STAGE_BYTES);
  uint64_t* full = bars;                  // [CX_STAGES]
  uint64_t* empty = bars + CX_STAGES;     // [CX_STAGES]
  uint64_t* t_full = empty + CX_STAGES;   // [2]
  uint64_t* t_empty = t_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (long long)p.n_mb * p.n_nt;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < CX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const long long mb = t / p.n_nt; const int nt = (int)(t % p.n_nt);
        for (int ks = 0; ks < p.kb; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], CX_STAGE_BYTES);
          unsigned char* s = smem + stage * CX_STAGE_BYTES;
          bulk_g2s(s, p.ximg + ((mb * 2 + 0) * p.kb + ks) * 32768, 32768, &full[stage]);
          bulk_g2s(s + 32768, p.ximg + ((mb * 2 + 1) * p.kb + ks) * 32768, 32768, &full[stage]);
          bulk_g2s(s + 65536, p.wimg + ((long long)nt * p.kb + ks) * 32768, 32768, &full[stage]);
          if (++stage == CX_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        mbar_wait(&t_empty[buf], tphase ^ 1);
        for (int ks = 0; ks < p.kb; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + stage * CX_STAGE_BYTES);
          const uint64_t b_hi = make_smem_desc(sb + 65536), b_lo = make_smem_desc(sb + 65536 + 16384);
#pragma unroll
          for (int ab = 0; ab < 2; ++ab) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((ab * 2 + buf) * 128);
            const uint64_t a_hi = make_smem_desc(sb + ab * 32768), a_lo = make_smem_desc(sb + ab * 32768 + 16384);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t o = (uint64_t)(k4 * 2);
              umma_f16(d_tmem, a_hi + o, b_hi + o, IDESC_F16_M128_N128, (uint32_t)((ks | k4) != 0));
              umma_f16(d_tmem, a_lo + o, b_hi + o, IDESC_F16_M128_N128, 1u);
              umma_f16(d_tmem, a_hi + o, b_lo + o, IDESC_F16_M128_N128, 1u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == CX_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&t_full[buf]);
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int half = ew >> 3, ab = (ew >> 2) & 1, quad = ew & 3;
    const float unscale = ldexpf(1.0f, -(p.xst->exp + p.wst->exp));
    float amax_out = 0.f;       // max |out| over this thread's elements (only used when p.out_amax is given)
    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t tphase = (it >> 1) & 1;
      const long long mb = t / p.n_nt; const int nt = (int)(t % p.n_nt);
      const int n0 = nt * 128 + half * 64;
      mbar_wait(&t_full[buf], tphase);
      tc_fence_after();
      // The accumulators arrive one ROW per lane.  Global memory wants the other orientation, so each 32x32 block is
      // transposed in registers (5 butterfly stages of shfl.xor): afterwards lane l holds COLUMN l of the 32 rows and every
      // x / x0 / out access of the warp is one contiguous 128-byte segment of a row.  The 64 columns are taken from TMEM in
      // two halves of 32 registers, which leaves room for 16 rows x 2 operands = 32 independent loads in flight per thread:
      // the epilogue is a latency-bound stream (4 x [B,D] of HBM traffic), and with 4-row batches it, not the MMAs, paced
      // the kernel (same time at K = 256 as at K = 845).
      const long long row_base = mb * 256 + ab * 128 + quad * 32;
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ab * 2 + buf) * 128 + half * 64 + blk * 32), r);
        tmem_ld_wait32(r);
        if (blk == 1) {   // both halves are in registers: the MMA warp may reuse this TMEM buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          const bool upper = (lane & s) != 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if ((i & s) == 0) {
              const uint32_t lo_v = r[i], hi_v = r[i | s];
              const uint32_t recv = __shfl_xor_sync(0xffffffffu, upper ? lo_v : hi_v, s);
              r[i] = upper ? recv : lo_v;
              r[i | s] = upper ? hi_v : recv;
            }
          }
        }
        // now r[j] = accumulator of row (row_base + j), column (n0 + blk*32 + lane)
        const int col = n0 + blk * 32 + lane;
        const long long base = row_base * p.ld + col;
        const int ldi = (int)p.ld;
        if (row_base + 32 <= p.B && n0 + blk * 32 + 32 <= p.D) {
          // interior block (all but the ragged last column block / row block): no per-element predicates or branches and
          // one 32-bit row offset shared by the four arrays -- the epilogue is INSTRUCTION-bound (ncu: issue slots, not
          // DRAM or the LSU, limit it; the checked path below costs ~70 instructions per element)
          const float* __restrict__ xp = p.x + base; const float* __restrict__ x0p = p.x0 + base;
          float* __restrict__ op = p.out + base; float* __restrict__ pp = p.prod ? p.prod + base : nullptr;
          const float bcol = p.bias ? __ldg(p.bias + col) : 0.f;
          const float diag = p.diag;
#pragma unroll
          for (int j0 = 0; j0 < 32; j0 += 16) {
            float xv[16], x0v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) { const int o = (j0 + u) * ldi; xv[u] = __ldg(xp + o); x0v[u] = __ldg(x0p + o); }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int o = (j0 + u) * ldi;
              float pv = fmaf(__uint_as_float(r[j0 + u]), unscale, bcol);
              pv = fmaf(diag, xv[u], pv);
              if (pp) pp[o] = pv;
              const float ov = fmaf(x0v[u], pv, xv[u]);
              op[o] = ov;
              amax_out = fmaxf(amax_out, fabsf(ov));
            }
          }
        } else if (col < p.D) {
          const float bcol = p.bias ? __ldg(p.bias + col) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const long long rr = row_base + j;
            if (rr < p.B) {
              const long long o = rr * p.ld + col;
              const float xv = __ldg(p.x + o), x0v = __ldg(p.x0 + o);
              float pv = fmaf(__uint_as_float(r[j]), unscale, bcol);
              pv = fmaf(p.diag, xv, pv);
              if (p.prod) p.prod[o] = pv;
              const float ov = fmaf(x0v, pv, xv);
              p.out[o] = ov;
              amax_out = fmaxf(amax_out, fabsf(ov));
            }
          }
        }
      }
    }
    if (p.out_amax) {   // same statistic, same bits, as a cx_amax_kernel pass over `out` (max is order-independent)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax_out = fmaxf(amax_out, __shfl_xor_sync(0xffffffffu, amax_out, o));
      if (lane == 0 && amax_out > 0.f) atomicMax(p.out_amax, __float_as_uint(amax_out));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace tfrs
using namespace tfrs;
using namespace tfrs::tc;

// ---- W image (built once per weight version, by the caller) --------------------------------------------
extern "C" size_t tfrs_cross_tc_weight_bytes(int D) {
  if (D <= 0) return 0;
  return 1024 + cx_img_bytes(D, D);
}
extern "C" int tfrs_cross_tc_weight_build(const float* W, int D, void* wbuf, size_t bytes, void* stream) {
  TFRS_CHECK_ARG(W && wbuf && D > 0, "cross_tc_weight_build: bad arguments");
  TFRS_CHECK_ARG(bytes >= tfrs_cross_tc_weight_bytes(D), "cross_tc_weight_build: buffer too small");
  TFRS_CHECK_ARG((reinterpret_cast<uintptr_t>(wbuf) & 15) == 0, "cross_tc_weight_build: buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CxStats* ws = (CxStats*)wbuf;
  TFRS_CUDA(cudaMemsetAsync(wbuf, 0, 1024, st));
  cx_amax_kernel<<<64, 256, 0, st>>>(W, D, D, D, ws);
  TFRS_LAUNCH_CHECK();
  cx_exp_kernel<<<1, 1, 0, st>>>(ws);
  TFRS_LAUNCH_CHECK();
  const int kb = (int)ceil_div(D, 64);
  const long long nt = ceil_div(D, 128);
  const long long chunks = nt * 128 * kb * 8;
  // rows of the image = output columns n; element (n, k) = W[k, n]  -> transposed read
  cx_split_image_kernel<true><<<(unsigned)ceil_div(chunks, 256), 256, 0, st>>>(W, D, D, D, kb, nt, ws, (unsigned char*)wbuf + 1024);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

extern "C" size_t tfrs_cross_tc_workspace_bytes(int64_t B, int D) {
  if (B <= 0 || D <= 0) return 0;
  return 1024 + cx_img_bytes(ceil_div(B, 256) * 256, D);
}

extern "C" int tfrs_cross_tc_fwd_ex_f32(const float* x0, const float* x, const void* wbuf, const float* bias, int64_t B, int D,
                                        int64_t ld, float diag_scale, float* out, float* prod, const unsigned int* x_amax_bits,
                                        unsigned int* out_amax_bits, void* ws, size_t ws_bytes, void* stream) {
  TFRS_CHECK_ARG(x0 && x && wbuf && out, "cross_tc_fwd: NULL pointer");
  TFRS_CHECK_ARG(B > 0 && D > 0 && ld >= D, "cross_tc_fwd: bad shape");
  TFRS_CHECK_ARG(diag_scale >= 0.f, "`diag_scale` should be non-negative. Got `diag_scale` = %g", diag_scale);
  if (!ws || ws_bytes < tfrs_cross_tc_workspace_bytes(B, D)) { set_error("cross_tc_fwd: workspace too small"); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  TFRS_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "cross_tc_fwd: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CxStats* xs = (CxStats*)ws;
  unsigned char* ximg = (unsigned char*)ws + 1024;
  const int kb = (int)ceil_div(D, 64);
  const int n_mb = (int)ceil_div(B, 256);
  const int n_nt = (int)ceil_div(D, 128);
  TFRS_CUDA(cudaMemsetAsync(ws, 0, 1024, st));
  if (x_amax_bits) {   // max |x| is already known (the previous layer's epilogue produced it): no pass over x
    TFRS_CUDA(cudaMemcpyAsync(&xs->amax_bits, x_amax_bits, sizeof(unsigned int), cudaMemcpyDeviceToDevice, st));
  } else {
    cx_amax_kernel<<<(unsigned)(148 * 8), 256, 0, st>>>(x, B, D, ld, xs);
    TFRS_LAUNCH_CHECK();
  }
  cx_exp_kernel<<<1, 1, 0, st>>>(xs);
  TFRS_LAUNCH_CHECK();
  {
    const long long chunks = (long long)n_mb * 2 * 128 * kb * 8;
    unsigned blocks = (unsigned)(ceil_div(chunks, 256) < (1 << 20) ? ceil_div(chunks, 256) : (1 << 20));
    cx_split_image_kernel<false><<<blocks, 256, 0, st>>>(x, B, D, ld, kb, (long long)n_mb * 2, xs, ximg);
    TFRS_LAUNCH_CHECK();
  }
  if (out_amax_bits) TFRS_CUDA(cudaMemsetAsync(out_amax_bits, 0, sizeof(unsigned int), st));
  CrossParams p{};
  p.ximg = ximg; p.wimg = (const unsigned char*)wbuf + 1024; p.xst = xs; p.wst = (const CxStats*)wbuf;
  p.x0 = x0; p.x = x; p.bias = bias; p.diag = diag_scale; p.out = out; p.prod = prod;
  p.B = B; p.D = D; p.ld = ld; p.kb = kb; p.n_mb = n_mb; p.n_nt = n_nt; p.out_amax = out_amax_bits;
  const size_t smem = (size_t)CX_STAGES * CX_STAGE_BYTES + 1024 + 256;
  TFRS_DYN_SMEM(cross_tc_kernel, (int)smem);
  long long tiles = (long long)n_mb * n_nt;
  int grid = sm_count(); if (grid > tiles) grid = (int)tiles;
  cross_tc_kernel<<<grid, CX_THREADS, smem, st>>>(p);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

extern "C" int tfrs_cross_tc_fwd_f32(const float* x0, const float* x, const void* wbuf, const float* bias, int64_t B, int D,
                                     int64_t ld, float diag_scale, float* out, float* prod, void* ws, size_t ws_bytes,
                                     void* stream) {
  return tfrs_cross_tc_fwd_ex_f32(x0, x, wbuf, bias, B, D, ld, diag_scale, out, prod, nullptr, nullptr, ws, ws_bytes, stream);
}
