// softmax_tc_bwd.cu -- K3b on the tensor cores: backward of the in-batch softmax loss of tfrs.tasks.Retrieval
//   (tape.gradient at models/base.py:77 through tasks/retrieval.py:178-210)
//   G_ij = (softmax(s)_ij - [i == j]) * w_i * grad_loss / T ;   dq = G . c  [B,d] ;   dc = G^T . q  [C,d]
// as two launches of ONE kernel (flash-attention-backward shape, deterministic -- no atomics):
//   stationary operand X (128 rows = TMEM lanes, resident in smem), streaming operand Y (128-row tiles, bulk-TMA ring)
//     S   = X . Y^T          tcgen05 SS-mode, hi/lo fp16 split operands (3 MMAs per K16), fp32 in TMEM
//     A   = (exp(S/T - lse_q) - diag) * w_q     computed by the epilogue warps from TMEM, split into fp16 hi/lo and
//                                               written back IN PLACE over S with tcgen05.st (128 fp32 cols -> 64+64)
//     dX += A . Y            tcgen05 TS-mode: A from TMEM, Y straight from the same smem tile as an MN-major operand
//   launch 1: X = q, Y = c  -> dq ;  launch 2: X = c, Y = q (lse/w become per-column vectors staged with the tile) -> dc.
// The [B,C] logits / probabilities never touch HBM; the scores are the same split products as the forward pass
// (softmax_tc.cu), so exp(s - lse) is consistent with the saved lse.  d <= 64.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_split.cuh"
#include "softmax_ext.cuh"

namespace tfrs {
namespace tc {

constexpr int SB_THREADS = 640;                // warp 0 producer, 1 MMA, 2 TMEM alloc, 4-19 epilogue (TMEM lane quad = warp & 3)
constexpr int SB_W_PROD = 0, SB_W_MMA = 1, SB_W_ALLOC = 2;
constexpr int SB_STAGES = 4;
// the tensor core's fp32 adder truncates: a chain of thousands of accumulations into the same TMEM tile drifts (5e-5 of
// the gradient scale at C = 16384), so the dX tile is drained into fp32 registers (round-to-nearest adds) every
// SB_DRAIN streamed tiles = 192 accumulation steps, the length of the forward Cross chain
constexpr int SB_DRAIN = 8;
constexpr int SB_BUFS = 3;                     // S/G accumulator buffers in TMEM (3 x 128 columns) + dX (64 columns)
constexpr int SB_Y_BYTES = 32768;              // one 128-row tile: hi 16 KB | lo 16 KB
constexpr int SB_STAGE_BYTES = SB_Y_BYTES + 1024;  // + lse[128] | w[128] of the tile (transposed launch)
constexpr float SB_LOG2E = 1.4426950408889634f;

struct SoftmaxBwdParams {
  const unsigned char* ximg; const unsigned char* yimg;
  const CxStats* xst; const CxStats* yst; const CxStats* wst;
  const float* lse_pad; const float* w_pad;   // indexed by QUERY, padded to a multiple of 128 (w already * 2^wst.exp)
  const float* grad_loss;
  const float* cbias_pad;                     // BIAS: per-CANDIDATE logit bias (natural units), padded to a multiple of 128
  long long n_x_rows, n_y_valid, n_ytiles, part_stride;
  int n_xb, parts, d;
  float inv_t;
  float* out;
  // EXT (accidental-hit removal / score_mask, see softmax_tc.cu): masked entries get G = 0
  const int* id_lo; const int* id_hi;   // candidate ids (32-bit halves, padded); query j's positive is candidate j
  const uint32_t* mbits; int mwords;    // keep-bits [stationary row][streamed column]: the (B,C) matrix for dq, its transpose for dc
};

// One thread's 64 accumulator columns -> A, split into fp16 hi (written in place over r[0..31]: output slot 2*j4+e
// is only written after inputs 4*j4.. are consumed) and lo[32].
//   non-transposed: A = (exp(s/T - lse_i) - diag) * 2^14 -- lse_r already carries the -14 ln2, the weight of row i is
//                   applied once to the dX block;   transposed: A = (exp(s/T - lse_j) - diag) * w^_j (per-column vectors).
// EDGE = the tile holds the diagonal or columns beyond the valid range (rare): masks compiled in only there.
// BIAS: logits carry a per-candidate bias b_c: the exponent is s/T + b_c - lse_q.  Non-transposed: candidates are the
// columns (cb4 = this thread's 64 biases, read through L2); transposed: the candidate is the row (bias_r).
// EXT: kill0 / kill1 = bit j set -> entry j (columns 0..31 / 32..63) is masked: A = 0 (where() blocks the gradient of a
// masked score, and exp(MIN_FLOAT - lse) = 0 for an accidental hit).
template <bool TRANSPOSED, bool EDGE, bool BIAS, bool EXT>
__device__ __forceinline__ void sb_transform(uint32_t (&r)[64], uint32_t (&lo)[32], float scale, float lse_r,
                                             const float4* __restrict__ aux4, int n_valid, int jd,
                                             const float4* __restrict__ cb4, float bias_r, uint32_t kill0, uint32_t kill1) {
#pragma unroll
  for (int j4 = 0; j4 < 16; ++j4) {
    float lq[4] = {lse_r, lse_r, lse_r, lse_r}, wq[4] = {16384.f, 16384.f, 16384.f, 16384.f};
    if (TRANSPOSED) {
      const float4 l4 = aux4[j4], w4 = aux4[32 + j4];
      lq[0] = l4.x; lq[1] = l4.y; lq[2] = l4.z; lq[3] = l4.w;
      wq[0] = w4.x; wq[1] = w4.y; wq[2] = w4.z; wq[3] = w4.w;
      if (BIAS) { lq[0] -= bias_r; lq[1] -= bias_r; lq[2] -= bias_r; lq[3] -= bias_r; }
    } else if (BIAS) {
      const float4 bb = __ldg(cb4 + j4);
      lq[0] -= bb.x; lq[1] -= bb.y; lq[2] -= bb.z; lq[3] -= bb.w;
    }
    float a[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * j4 + e;
      float pr = ex2_approx(fmaf(__uint_as_float(r[j]), scale, -lq[e]) * SB_LOG2E);
      if (TRANSPOSED) {
        if (EDGE && j == jd) pr -= 1.0f;
        pr *= wq[e];
      } else {
        if (EDGE && j == jd) pr -= 16384.f;
      }
      a[e] = (!EDGE || j < n_valid) ? pr : 0.f;
      if (EXT && (((j < 32 ? kill0 : kill1) >> (j & 31)) & 1u)) a[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const __half2 h = __floats2half2_rn(a[2 * e], a[2 * e + 1]);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(a[2 * e] - hf.x, a[2 * e + 1] - hf.y);
      r[2 * j4 + e] = *reinterpret_cast<const uint32_t*>(&h);
      lo[2 * j4 + e] = *reinterpret_cast<const uint32_t*>(&l);
    }
  }
}

template <bool TRANSPOSED, int MODE>
__global__ void __launch_bounds__(SB_THREADS, 1)
softmax_tc_bwd_kernel(const SoftmaxBwdParams p) {
  constexpr bool BIAS = MODE >= 1;
  constexpr bool EXT = MODE == 2;
  extern __shared__ __align__(1024) unsigned char sb_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sb_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sX = smem;                       // 32 KB
  unsigned char* sY = smem + 32768;               // SB_STAGES x SB_STAGE_BYTES
  uint64_t* bars = reinterpret_cast<uint64_t*>(sY + SB_STAGES * SB_STAGE_BYTES);
  uint64_t* y_full = bars;
  uint64_t* y_empty = bars + SB_STAGES;
  uint64_t* x_full = bars + 2 * SB_STAGES;
  uint64_t* s_full = x_full + 1;     // [SB_BUFS]
  uint64_t* g_ready = s_full + SB_BUFS;
  uint64_t* dx_full = g_ready + SB_BUFS;
  uint64_t* dx_drained = dx_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dx_drained + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xb = blockIdx.x % p.n_xb, part = blockIdx.x / p.n_xb;
  const long long t_begin = (long long)part * p.n_ytiles / p.parts;
  const long long t_end = (long long)(part + 1) * p.n_ytiles / p.parts;
  const int n_iter = (int)(t_end - t_begin);

  if (warp == SB_W_MMA && lane == 0) {
    for (int s = 0; s < SB_STAGES; ++s) { mbar_init(&y_full[s], 1); mbar_init(&y_empty[s], 1); }
    mbar_init(x_full, 1);
    for (int b = 0; b < SB_BUFS; ++b) { mbar_init(&s_full[b], 1); mbar_init(&g_ready[b], 8); }
    mbar_init(dx_full, 1);
    mbar_init(dx_drained, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == SB_W_ALLOC) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_dx = tmem_base + SB_BUFS * 128;

  if (warp == SB_W_PROD) {
    if (lane == 0) {
      mbar_expect_tx(x_full, 32768);
      bulk_g2s(sX, p.ximg + (long long)xb * 32768, 32768, x_full);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(&y_empty[stage], phase ^ 1);
        unsigned char* dst = sY + stage * SB_STAGE_BYTES;
        const long long tile = t_begin + it;
        mbar_expect_tx(&y_full[stage], TRANSPOSED ? SB_STAGE_BYTES : SB_Y_BYTES);
        bulk_g2s(dst, p.yimg + tile * SB_Y_BYTES, SB_Y_BYTES, &y_full[stage]);
        if (TRANSPOSED) {
          bulk_g2s(dst + SB_Y_BYTES, p.lse_pad + tile * 128, 512, &y_full[stage]);
          bulk_g2s(dst + SB_Y_BYTES + 512, p.w_pad + tile * 128, 512, &y_full[stage]);
        }
        if (++stage == SB_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == SB_W_MMA) {
    if (lane == 0) {
      mbar_wait(x_full, 0);
      tc_fence_after();
      const uint32_t x0 = smem_u32(sX);
      const uint64_t x_hi = make_smem_desc(x0), x_lo = make_smem_desc(x0 + 16384);
      // dX(u) += A(u) . Y(u):  A from TMEM buffer u&1 (hi | lo per 64-column half), Y tile as MN-major B (K = its rows)
      auto issue_dx = [&](int u) {
        const int buf = u % SB_BUFS, stage = u % SB_STAGES;
        const int chunk = u / SB_DRAIN, first = (u % SB_DRAIN) == 0;
        if (first && chunk > 0) mbar_wait(dx_drained, (uint32_t)((chunk - 1) & 1));  // epilogue took the previous chunk's sum
        mbar_wait(&g_ready[buf], (uint32_t)((u / SB_BUFS) & 1));
        tc_fence_after();
        const uint32_t y0 = smem_u32(sY + stage * SB_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t a_hi = tmem_base + (uint32_t)(buf * 128 + 64 * (j >> 2) + 8 * (j & 3));
          const uint32_t a_lo = a_hi + 32;
          const uint64_t b_hi = make_smem_desc(y0 + j * 2048), b_lo = make_smem_desc(y0 + 16384 + j * 2048);
          umma_f16_ts(tmem_dx, a_hi, b_hi, IDESC_F16_M128_N64_BMN, (uint32_t)(!first || j != 0));
          umma_f16_ts(tmem_dx, a_lo, b_hi, IDESC_F16_M128_N64_BMN, 1u);
          umma_f16_ts(tmem_dx, a_hi, b_lo, IDESC_F16_M128_N64_BMN, 1u);
        }
        umma_commit(&y_empty[stage]);
        if ((u % SB_DRAIN) == SB_DRAIN - 1 || u == n_iter - 1) umma_commit(dx_full);  // chunk complete
      };
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it % SB_BUFS, stage = it % SB_STAGES;
        mbar_wait(&y_full[stage], (uint32_t)((it / SB_STAGES) & 1));
        tc_fence_after();
        const uint32_t y0 = smem_u32(sY + stage * SB_STAGE_BYTES);
        const uint64_t y_hi = make_smem_desc(y0), y_lo = make_smem_desc(y0 + 16384);
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t o = (uint64_t)(k4 * 2);
          // same three products, in the same order, as the forward pass (q_hi c_hi, q_lo c_hi, q_hi c_lo)
          umma_f16(d_tmem, x_hi + o, y_hi + o, IDESC_F16_M128_N128, (uint32_t)(k4 != 0));
          if (TRANSPOSED) {
            umma_f16(d_tmem, x_hi + o, y_lo + o, IDESC_F16_M128_N128, 1u);
            umma_f16(d_tmem, x_lo + o, y_hi + o, IDESC_F16_M128_N128, 1u);
          } else {
            umma_f16(d_tmem, x_lo + o, y_hi + o, IDESC_F16_M128_N128, 1u);
            umma_f16(d_tmem, x_hi + o, y_lo + o, IDESC_F16_M128_N128, 1u);
          }
        }
        umma_commit(&s_full[buf]);
        if (it >= 2) issue_dx(it - 2);   // two tiles behind: the epilogue of tile it-2 has had two S-MMA times to finish
      }
      if (n_iter >= 2) issue_dx(n_iter - 2);
      issue_dx(n_iter - 1);
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int grp = ew >> 3, half = (ew >> 2) & 1, quad = ew & 3;
    const int r_local = quad * 32 + lane;
    const long long row = (long long)xb * 128 + r_local;
    const float scale = ldexpf(p.inv_t, -(p.xst->exp + p.yst->exp));  // accumulator -> logit (natural units)
    float lse_r = 0.f, w_r = 1.f;
    const float bias_r = (BIAS && TRANSPOSED) ? p.cbias_pad[row] : 0.f;  // padded: in range for every row of the block
    if (!TRANSPOSED) {  // padded arrays: in range for every row of the block
      lse_r = p.lse_pad[row] - 14.0f * 0.6931471805599453f;  // folds the 2^14 fp16 range scale into the exponent
      w_r = p.w_pad[row];                                     // w_i * 2^wst.exp, applied to the dX row at the end
    }
    // dX block: 128 rows x 64 columns of fp32 in TMEM; thread = (row, 16 columns).  Chunk k (tiles [8k, 8k+8)) is added
    // into dacc once its last dX MMA has retired; a warp does that before it waits for a tile >= 3 past the chunk end
    // (everything that chunk needs was issued before the MMA thread can block on dx_drained: no circular wait).
    const int c0 = grp * 32 + half * 16;
    const uint32_t dx_addr = tmem_dx + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
    // running fp32 sum of the drained chunks: 16 values per thread, kept in SHARED memory ([i][epilogue thread]: conflict-free)
    // -- in registers they pushed the transform loop (64 accumulators + 32 lo words live) past the 96-register budget
    float* dacc = reinterpret_cast<float*>(smem + 32768 + SB_STAGES * SB_STAGE_BYTES + 1024) + (ew * 32 + lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) dacc[i * 512] = 0.f;
    int rid_lo = 0, rid_hi = 0;   // id of the stationary row: the positive's id of query `row` (dq) / of candidate `row` (dc)
    if (EXT && p.id_lo) { rid_lo = p.id_lo[row]; rid_hi = p.id_hi[row]; }
    const int n_chunks = (n_iter + SB_DRAIN - 1) / SB_DRAIN;
    int next_chunk = 0;
    auto drain_until = [&](int t_next) {  // drain every chunk whose last tile is <= t_next - 3
      while (next_chunk < n_chunks && min(next_chunk * SB_DRAIN + SB_DRAIN - 1, n_iter - 1) + 3 <= t_next) {
        mbar_wait(dx_full, (uint32_t)(next_chunk & 1));
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld16(dx_addr, acc);
        tmem_ld_wait16(acc);
#pragma unroll
        for (int i = 0; i < 16; ++i) dacc[i * 512] += __uint_as_float(acc[i]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dx_drained);
        ++next_chunk;
      }
    };
    for (int it = grp; it < n_iter; it += 2) {
      drain_until(it);
      const int stage = it % SB_STAGES, buf = it % SB_BUFS;
      const long long col0 = (t_begin + it) * 128 + half * 64;
      mbar_wait(&s_full[buf], (uint32_t)((it / SB_BUFS) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 128 + half * 64);
      uint32_t r[64];
      tmem_ld64(taddr, r);
      tmem_ld_wait64(r);
      if (TRANSPOSED) mbar_wait(&y_full[stage], (uint32_t)((it / SB_STAGES) & 1));  // the tile's lse | w vectors (bulk-copied)
      const float4* aux4 = reinterpret_cast<const float4*>(sY + stage * SB_STAGE_BYTES + SB_Y_BYTES) + half * 16;
      const int n_valid = (int)max(0ll, min(64ll, p.n_y_valid - col0));
      // the positive (query i <-> candidate i) can only sit in the tile whose columns cover this block's rows
      const bool edge = n_valid < 64 || ((long long)xb * 128 < col0 + 64 && col0 < (long long)xb * 128 + 128);
      uint32_t lo[32];
      const float4* cb4 = (BIAS && !TRANSPOSED) ? reinterpret_cast<const float4*>(p.cbias_pad + col0) : nullptr;
      uint32_t kill0 = 0, kill1 = 0;
      if (EXT) {
        if (p.mbits) {
          const uint2 mw = __ldg(reinterpret_cast<const uint2*>(p.mbits + row * p.mwords + (t_begin + it) * 4 + half * 2));
          kill0 = ~mw.x; kill1 = ~mw.y;
        }
        if (p.id_lo) {
          const int4* il = reinterpret_cast<const int4*>(p.id_lo + col0);
#pragma unroll
          for (int j4 = 0; j4 < 16; ++j4) {
            const int4 v = __ldg(il + j4);
            if ((v.x == rid_lo) | (v.y == rid_lo) | (v.z == rid_lo) | (v.w == rid_lo)) {   // rare
              const int* ih = p.id_hi + col0 + 4 * j4;
              const long long c = col0 + 4 * j4;
              uint32_t hit = 0;
              if (v.x == rid_lo && __ldg(ih + 0) == rid_hi && c + 0 != row) hit |= 1u;
              if (v.y == rid_lo && __ldg(ih + 1) == rid_hi && c + 1 != row) hit |= 2u;
              if (v.z == rid_lo && __ldg(ih + 2) == rid_hi && c + 2 != row) hit |= 4u;
              if (v.w == rid_lo && __ldg(ih + 3) == rid_hi && c + 3 != row) hit |= 8u;
              if (j4 < 8) kill0 |= hit << (4 * j4); else kill1 |= hit << (4 * (j4 - 8));
            }
          }
        }
      }
      if (edge) {
        const int jd = (row >= col0 && row < col0 + 64) ? (int)(row - col0) : -1;
        sb_transform<TRANSPOSED, true, BIAS, EXT>(r, lo, scale, lse_r, aux4, n_valid, jd, cb4, bias_r, kill0, kill1);
      } else {
        sb_transform<TRANSPOSED, false, BIAS, EXT>(r, lo, scale, lse_r, aux4, 64, -1, cb4, bias_r, kill0, kill1);
      }
      tmem_st32(taddr, r);        // hi: columns [0, 32) of this 64-column half
      tmem_st32(taddr + 32, lo);  // lo: columns [32, 64)
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_ready[buf]);
    }
    drain_until(n_iter + 3 + SB_DRAIN);  // the remaining chunks
    if (row < p.n_x_rows) {
      const float gl = p.grad_loss ? p.grad_loss[0] : 1.0f;
      // transposed: A carried w^ = w 2^wexp per column; otherwise A carried 2^14 and the row weight is applied here
      const float fs = TRANSPOSED ? ldexpf(gl * p.inv_t, -(p.wst->exp + p.yst->exp))
                                  : ldexpf(gl * p.inv_t * w_r, -(p.wst->exp + 14 + p.yst->exp));
      float* dst = p.out + (long long)part * p.part_stride + row * p.d;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c0 + i < p.d) dst[c0 + i] = dacc[i * 512] * fs;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SB_W_ALLOC) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// max |w| (1 when w == NULL) -> the exact power-of-two scale that puts it in [2^13, 2^14)
__global__ void __launch_bounds__(1024) sb_wstats_kernel(const float* __restrict__ w, long long B, CxStats* __restrict__ st) {
  __shared__ float red[32];
  float a = w ? 0.f : 1.0f;
  if (w) for (long long i = threadIdx.x; i < B; i += 1024) a = fmaxf(a, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 32; ++i) a = fmaxf(a, red[i]);
    int x = 0;
    const bool ok = a > 0.f && a < INFINITY;
    if (ok) (void)frexpf(a, &x);
    st->amax_bits = __float_as_uint(a);
    st->exp = ok ? (CX_TARGET_EXP - x) : 0;
  }
}
__global__ void __launch_bounds__(256)
sb_prep_kernel(const float* __restrict__ lse, const float* __restrict__ w, long long B, long long Bpad,
               const CxStats* __restrict__ wst, float* __restrict__ lse_pad, float* __restrict__ w_pad) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= Bpad) return;
  lse_pad[i] = i < B ? lse[i] : 0.f;
  w_pad[i] = i < B ? ldexpf(w ? w[i] : 1.0f, wst->exp) : 0.f;
}
// out[e] = sum_z partial[z][e], z ascending (deterministic)
__global__ void __launch_bounds__(256)
sb_reduce_parts_kernel(const float* __restrict__ partial, long long elems, int parts, float* __restrict__ out) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= elems) return;
  float a = partial[e];
  for (int z = 1; z < parts; ++z) a += partial[(long long)z * elems + e];
  out[e] = a;
}

// streaming-range splits per stationary block: minimise waves x (tiles per CTA + fixed cost of ~6 tile times:
// X load, pipeline fill/drain, dX epilogue, partial reduction)
static int sb_parts(long long n_xb, long long n_ytiles) {
  int parts = 1; double best = 1e30;
  const int sms = sm_count();
  for (int c = 1; c <= 16 && c <= n_ytiles; ++c) {
    const double cost = (double)ceil_div(n_xb * c, sms) * ((double)ceil_div(n_ytiles, c) + 6.0);
    if (cost < best * 0.97) { best = cost; parts = c; }
  }
  return parts;
}

struct SbPlan {
  long long q_tiles, c_tiles; int parts_q, parts_c;
  size_t o_qst, o_cst, o_wst, o_qimg, o_cimg, o_lse, o_w, o_bias, o_partial, o_idlo, o_idhi, o_mbits, o_mbits_t, total;
};
static bool sb_plan(long long B, long long C, int d, SbPlan& pl, bool has_ids = false, bool has_mask = false) {
  if (B <= 0 || C < B || d <= 0 || d > 64) return false;
  pl.q_tiles = ceil_div(B, 128); pl.c_tiles = ceil_div(C, 128);
  pl.parts_q = sb_parts(pl.q_tiles, pl.c_tiles);   // dq: X = q blocks, Y = c tiles
  pl.parts_c = sb_parts(pl.c_tiles, pl.q_tiles);   // dc: X = c blocks, Y = q tiles
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 1024); return r; };
  pl.o_qst = take(sizeof(CxStats)); pl.o_cst = take(sizeof(CxStats)); pl.o_wst = take(sizeof(CxStats));
  pl.o_qimg = take((size_t)pl.q_tiles * 32768);
  pl.o_cimg = take((size_t)pl.c_tiles * 32768);
  pl.o_lse = take((size_t)pl.q_tiles * 128 * 4);
  pl.o_w = take((size_t)pl.q_tiles * 128 * 4);
  pl.o_bias = take((size_t)pl.c_tiles * 128 * 4);
  size_t pq = pl.parts_q > 1 ? (size_t)pl.parts_q * B * d * 4 : 0, pc = pl.parts_c > 1 ? (size_t)pl.parts_c * C * d * 4 : 0;
  pl.o_partial = take(pq > pc ? pq : pc);
  pl.o_idlo = take(has_ids ? (size_t)pl.c_tiles * 128 * 4 : 0);
  pl.o_idhi = take(has_ids ? (size_t)pl.c_tiles * 128 * 4 : 0);
  pl.o_mbits = take(has_mask ? (size_t)pl.q_tiles * 128 * pl.c_tiles * 4 * 4 : 0);     // [q rows][c words]
  pl.o_mbits_t = take(has_mask ? (size_t)pl.c_tiles * 128 * pl.q_tiles * 4 * 4 : 0);   // [c rows][q words]
  pl.total = o;
  return true;
}

}  // namespace tc
}  // namespace tfrs
using namespace tfrs;
using namespace tfrs::tc;

extern "C" size_t tfrs_inbatch_softmax_tc_bwd_workspace_bytes(int64_t B, int64_t C, int d) {
  SbPlan pl;
  return sb_plan(B, C, d, pl) ? pl.total : 0;
}

// pad[i] = src[i] for i < n, 0 on the padding
__global__ void __launch_bounds__(256) sb_pad_kernel(const float* __restrict__ src, long long n, long long npad, float* __restrict__ pad) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < npad) pad[i] = i < n ? src[i] : 0.f;
}

extern "C" size_t tfrs_inbatch_softmax_tc_bwd_ex_workspace_bytes(int64_t B, int64_t C, int d, int has_ids, int has_mask) {
  SbPlan pl;
  return sb_plan(B, C, d, pl, has_ids != 0, has_mask != 0) ? pl.total : 0;
}

extern "C" int tfrs_inbatch_softmax_tc_bwd_ex(const float* q, const float* c, int64_t B, int64_t C, int d, float inv_temperature,
                                              const float* sample_weight, const float* candidate_bias,
                                              const int64_t* candidate_ids, const uint8_t* score_mask, const float* lse,
                                              const float* grad_loss, float* dq, float* dc, void* ws, size_t ws_bytes, void* stream) {
  TFRS_CHECK_ARG(q && c && lse && dq && dc, "inbatch_softmax_tc_bwd: NULL pointer");
  SbPlan pl;
  const bool ext = candidate_ids || score_mask;
  if (!sb_plan(B, C, d, pl, candidate_ids != nullptr, score_mask != nullptr)) { set_error("inbatch_softmax_tc_bwd: shape outside the tensor-core path (need B <= C, d <= 64)"); return TFRS_ERR_UNSUPPORTED; }
  if (!ws || ws_bytes < pl.total) { set_error("inbatch_softmax_tc_bwd: workspace too small"); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  TFRS_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "inbatch_softmax_tc_bwd: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w8 = (unsigned char*)ws;
  CxStats* qst = (CxStats*)(w8 + pl.o_qst); CxStats* cst = (CxStats*)(w8 + pl.o_cst); CxStats* wst = (CxStats*)(w8 + pl.o_wst);
  unsigned char* qimg = w8 + pl.o_qimg; unsigned char* cimg = w8 + pl.o_cimg;
  float* lse_pad = (float*)(w8 + pl.o_lse); float* w_pad = (float*)(w8 + pl.o_w); float* partial = (float*)(w8 + pl.o_partial);
  TFRS_CUDA(cudaMemsetAsync(w8, 0, 3072, st));
  cx_amax_kernel<<<cx_amax_grid(B), 256, 0, st>>>(q, B, d, d, qst);
  TFRS_LAUNCH_CHECK();
  cx_amax_kernel<<<cx_amax_grid(C), 256, 0, st>>>(c, C, d, d, cst);
  TFRS_LAUNCH_CHECK();
  cx_exp_kernel<<<1, 1, 0, st>>>(qst);
  TFRS_LAUNCH_CHECK();
  cx_exp_kernel<<<1, 1, 0, st>>>(cst);
  TFRS_LAUNCH_CHECK();
  cx_split_image_kernel<false><<<(unsigned)ceil_div(pl.q_tiles * 128 * 8, 256), 256, 0, st>>>(q, B, d, d, 1, pl.q_tiles, qst, qimg);
  TFRS_LAUNCH_CHECK();
  cx_split_image_kernel<false><<<(unsigned)ceil_div(pl.c_tiles * 128 * 8, 256), 256, 0, st>>>(c, C, d, d, 1, pl.c_tiles, cst, cimg);
  TFRS_LAUNCH_CHECK();
  sb_wstats_kernel<<<1, 1024, 0, st>>>(sample_weight, B, wst);
  TFRS_LAUNCH_CHECK();
  sb_prep_kernel<<<(unsigned)ceil_div(pl.q_tiles * 128, 256), 256, 0, st>>>(lse, sample_weight, B, pl.q_tiles * 128, wst, lse_pad, w_pad);
  TFRS_LAUNCH_CHECK();
  const size_t smem = 32768 + (size_t)SB_STAGES * SB_STAGE_BYTES + 1024 + 32768 + 256;  // X | Y ring | barriers | dX sums | align slack
  TFRS_DYN_SMEM((softmax_tc_bwd_kernel<false, 0>), (int)smem);
  TFRS_DYN_SMEM((softmax_tc_bwd_kernel<true, 0>), (int)smem);
  TFRS_DYN_SMEM((softmax_tc_bwd_kernel<false, 1>), (int)smem);
  TFRS_DYN_SMEM((softmax_tc_bwd_kernel<true, 1>), (int)smem);
  TFRS_DYN_SMEM((softmax_tc_bwd_kernel<false, 2>), (int)smem);
  TFRS_DYN_SMEM((softmax_tc_bwd_kernel<true, 2>), (int)smem);
  SoftmaxBwdParams p{};
  p.xst = qst; p.yst = cst; p.wst = wst; p.lse_pad = lse_pad; p.w_pad = w_pad; p.grad_loss = grad_loss; p.d = d; p.inv_t = inv_temperature;
  const int mode = ext ? 2 : (candidate_bias ? 1 : 0);
  if (mode) {   // EXT without a bias runs on a zero bias vector
    float* cbp = (float*)(w8 + pl.o_bias);
    if (candidate_bias) sb_pad_kernel<<<(unsigned)ceil_div(pl.c_tiles * 128, 256), 256, 0, st>>>(candidate_bias, C, pl.c_tiles * 128, cbp);
    else TFRS_CUDA(cudaMemsetAsync(cbp, 0, (size_t)pl.c_tiles * 128 * 4, st));
    TFRS_LAUNCH_CHECK();
    p.cbias_pad = cbp;
  }
  uint32_t* mbits = nullptr; uint32_t* mbits_t = nullptr;
  if (candidate_ids) {
    int* lo = (int*)(w8 + pl.o_idlo); int* hi = (int*)(w8 + pl.o_idhi);
    sx_ids_split_kernel<<<(unsigned)ceil_div(pl.c_tiles * 128, 256), 256, 0, st>>>((const long long*)candidate_ids, C, pl.c_tiles * 128, lo, hi);
    TFRS_LAUNCH_CHECK();
    p.id_lo = lo; p.id_hi = hi;
  }
  if (score_mask) {
    mbits = (uint32_t*)(w8 + pl.o_mbits); mbits_t = (uint32_t*)(w8 + pl.o_mbits_t);
    const int wc = (int)(pl.c_tiles * 4), wq = (int)(pl.q_tiles * 4);
    sx_mask_pack_kernel<<<(unsigned)ceil_div(pl.q_tiles * 128 * wc, 256), 256, 0, st>>>(score_mask, B, C, pl.q_tiles * 128, wc, mbits);
    TFRS_LAUNCH_CHECK();
    sx_mask_pack_t_kernel<<<dim3((unsigned)ceil_div(pl.c_tiles * 128, 256), (unsigned)wq), 256, 0, st>>>(score_mask, B, C, pl.c_tiles * 128, wq, mbits_t);
    TFRS_LAUNCH_CHECK();
  }
  // ---- dq: X = q, Y = c
  p.ximg = qimg; p.yimg = cimg; p.n_x_rows = B; p.n_y_valid = C; p.n_ytiles = pl.c_tiles; p.n_xb = (int)pl.q_tiles; p.parts = pl.parts_q;
  p.part_stride = B * (long long)d; p.out = pl.parts_q > 1 ? partial : dq;
  p.mbits = mbits; p.mwords = (int)(pl.c_tiles * 4);
  {
    const unsigned g = (unsigned)(pl.q_tiles * pl.parts_q);
    if (mode == 2) softmax_tc_bwd_kernel<false, 2><<<g, SB_THREADS, smem, st>>>(p);
    else if (mode == 1) softmax_tc_bwd_kernel<false, 1><<<g, SB_THREADS, smem, st>>>(p);
    else softmax_tc_bwd_kernel<false, 0><<<g, SB_THREADS, smem, st>>>(p);
  }
  TFRS_LAUNCH_CHECK();
  if (pl.parts_q > 1) {
    sb_reduce_parts_kernel<<<(unsigned)ceil_div(B * d, 256), 256, 0, st>>>(partial, B * (long long)d, pl.parts_q, dq);
    TFRS_LAUNCH_CHECK();
  }
  // ---- dc: X = c, Y = q (only the B query rows exist; candidates beyond B are pure negatives)
  p.ximg = cimg; p.yimg = qimg; p.xst = cst; p.yst = qst; p.n_x_rows = C; p.n_y_valid = B; p.n_ytiles = pl.q_tiles; p.n_xb = (int)pl.c_tiles;
  p.parts = pl.parts_c; p.part_stride = C * (long long)d; p.out = pl.parts_c > 1 ? partial : dc;
  p.mbits = mbits_t; p.mwords = (int)(pl.q_tiles * 4);
  {
    const unsigned g = (unsigned)(pl.c_tiles * pl.parts_c);
    if (mode == 2) softmax_tc_bwd_kernel<true, 2><<<g, SB_THREADS, smem, st>>>(p);
    else if (mode == 1) softmax_tc_bwd_kernel<true, 1><<<g, SB_THREADS, smem, st>>>(p);
    else softmax_tc_bwd_kernel<true, 0><<<g, SB_THREADS, smem, st>>>(p);
  }
  TFRS_LAUNCH_CHECK();
  if (pl.parts_c > 1) {
    sb_reduce_parts_kernel<<<(unsigned)ceil_div(C * d, 256), 256, 0, st>>>(partial, C * (long long)d, pl.parts_c, dc);
    TFRS_LAUNCH_CHECK();
  }
  return TFRS_OK;
}

extern "C" int tfrs_inbatch_softmax_tc_bwd(const float* q, const float* c, int64_t B, int64_t C, int d, float inv_temperature,
                                           const float* sample_weight, const float* candidate_bias, const float* lse,
                                           const float* grad_loss, float* dq, float* dc, void* ws, size_t ws_bytes, void* stream) {
  return tfrs_inbatch_softmax_tc_bwd_ex(q, c, B, C, d, inv_temperature, sample_weight, candidate_bias, nullptr, nullptr, lse, grad_loss,
                                        dq, dc, ws, ws_bytes, stream);
}
