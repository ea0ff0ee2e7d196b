// softmax_tc.cu -- K3 forward on the tensor cores: in-batch sampled-softmax loss of tfrs.tasks.Retrieval
//   tasks/retrieval.py:178-180 (scores = q . c^T), :185 (labels = eye), :187-188 (/temperature), :210 + :86-87
//   loss = sum_i w_i * (logsumexp_j (q_i . c_j / T)  -  q_i . c_i / T)
// as ONE tcgen05 GEMM whose epilogue keeps an online (max, sum-exp) per query row and picks the diagonal --
// the [B,C] logits and the eye() labels never exist.  fp32 parity: hi/lo fp16 split of both operands (tc_split.cuh),
// 3 MMAs per K step, fp32 accumulation in TMEM (~2^-21 relative on a score).
//
// CTA shape = the top-K scan's: 256 query rows (two 128-row A blocks, hi+lo, resident), candidate tiles of
// 128 rows streamed through a bulk-TMA ring, 2x2 TMEM accumulator buffers, 16 epilogue warps (one row x 64
// columns per thread).  Grid = query blocks x candidate parts; every (row, part, column-half) leaves a partial
// (max, sum-exp) pair that `smtc_combine_kernel` folds into lse_i and the weighted row loss; the scalar loss is
// reduced in fixed order in fp64 (deterministic).  The backward pass (softmax.cu) consumes the same `lse`.
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "tc_split.cuh"
#include "softmax_ext.cuh"

namespace tfrs {
namespace tc {

constexpr int SX_THREADS = 640;
// candidate-tile ring depth: 4 x 32 KB at d <= 64, 1 x 64 KB at d <= 128 (A = 128 KB there)
__host__ __device__ constexpr int sx_stages(int kb) { return kb == 1 ? 4 : 1; }
constexpr float SX_LOG2E = 1.4426950408889634f;

struct SoftmaxTcParams {
  const unsigned char* qimg;  // [2*nqb tiles][kb][hi|lo][16 KB]
  const unsigned char* cimg;  // [n_ctiles][kb][hi|lo][16 KB]
  const CxStats* qst; const CxStats* cst;
  long long B, C;
  int nqb, parts, kb;
  long long n_ctiles;
  float inv_t;
  float2* partial;            // [Bp, parts, 2] (max, sum-exp) in log2 units
  float* pos;                 // [Bp] positive logit (q_i.c_i/T + bias_i, after the masks) in log2 units
  const float* cbias2;        // MODE >= 1: per-candidate logit bias in log2 units, padded to n_ctiles*128 (zeros beyond C)
  // MODE 2 (the remaining Retrieval options, each nullable):
  const int* id_lo; const int* id_hi;   // candidate ids split into 32-bit halves, padded: remove_accidental_hits
  const uint32_t* mbits; int mwords;    // score_mask as bits [Bp][mwords = n_ctiles*4] (1 = keep)
};

// MIN_FLOAT (layers/loss.py:23: float32 min / 100) in log2 units: the value a masked logit takes
constexpr float SX_MIN2 = -3.4028235e36f * SX_LOG2E;

// MODE 0: plain; 1: + per-candidate bias (sampling-probability correction); 2: + accidental-hit removal (candidate ids,
// tasks/retrieval.py:194-200, layers/loss.py:114-147: logits + dup * MIN_FLOAT == MIN_FLOAT in fp32) and score_mask
// (retrieval.py:202-203: where(mask, s, MIN_FLOAT)) applied to the accumulators in registers.
template <int KB, int MODE>
__global__ void __launch_bounds__(SX_THREADS, 1)
softmax_tc_kernel(const SoftmaxTcParams p) {
  constexpr bool BIAS = MODE >= 1;
  constexpr bool EXT = MODE == 2;
  extern __shared__ __align__(1024) unsigned char sx_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sx_raw) + 1023) & ~uintptr_t(1023));
  constexpr int SX_STAGES = sx_stages(KB);
  constexpr int A_BYTES = 2 * KB * 32768;      // two A blocks, hi+lo per K slab
  constexpr int B_BYTES = KB * 32768;          // one candidate tile, hi+lo per K slab
  unsigned char* sA = smem;
  unsigned char* sB = smem + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + SX_STAGES * B_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + SX_STAGES;
  uint64_t* a_full = bars + 2 * SX_STAGES;
  uint64_t* t_full = a_full + 1;     // [ab][buf]: the two 128-row halves signal separately, so one half's epilogue warps
  uint64_t* t_empty = t_full + 4;    //            pull from TMEM while the other half's are busy on the MUFU pipe
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x % p.nqb, part = blockIdx.x / p.nqb;
  const long long t_begin = (long long)part * p.n_ctiles / p.parts;
  const long long t_end = (long long)(part + 1) * p.n_ctiles / p.parts;
  const int n_iter = (int)(t_end - t_begin);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SX_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 1);
    for (int b = 0; b < 4; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, A_BYTES);
      bulk_g2s(sA, p.qimg + (long long)qb * A_BYTES, A_BYTES, a_full);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], B_BYTES);
        bulk_g2s(sB + stage * B_BYTES, p.cimg + (t_begin + it) * (long long)B_BYTES, B_BYTES, &full[stage]);
        if (++stage == SX_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(a_full, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        mbar_wait(&full[stage], phase);
#pragma unroll
        for (int ab = 0; ab < 2; ++ab) {
          mbar_wait(&t_empty[ab * 2 + buf], tphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)((ab * 2 + buf) * 128);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t a0 = smem_u32(sA + (ab * KB + kb) * 32768), b0 = smem_u32(sB + stage * B_BYTES + kb * 32768);
            const uint64_t a_hi = make_smem_desc(a0), a_lo = make_smem_desc(a0 + 16384);
            const uint64_t b_hi = make_smem_desc(b0), b_lo = make_smem_desc(b0 + 16384);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t o = (uint64_t)(k4 * 2);
              umma_f16(d_tmem, a_hi + o, b_hi + o, IDESC_F16_M128_N128, (uint32_t)((kb | k4) != 0));
              umma_f16(d_tmem, a_lo + o, b_hi + o, IDESC_F16_M128_N128, 1u);
              umma_f16(d_tmem, a_hi + o, b_lo + o, IDESC_F16_M128_N128, 1u);
            }
          }
          umma_commit(&t_full[ab * 2 + buf]);
        }
        umma_commit(&empty[stage]);
        if (++stage == SX_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int half = ew >> 3, ab = (ew >> 2) & 1, quad = ew & 3;
    const long long row = (long long)qb * 256 + ab * 128 + quad * 32 + lane;
    // logits in log2 units: s2 = acc * 2^-(eq+ec) * invT * log2(e)
    const float scale2 = ldexpf(p.inv_t * SX_LOG2E, -(p.qst->exp + p.cst->exp));
    float m2 = -INFINITY, l = 0.f, pos2 = 0.f;
    int rid_lo = 0, rid_hi = 0;   // id of this row's positive candidate (candidate `row`); the arrays are padded past Bp
    if (EXT && p.id_lo) { rid_lo = p.id_lo[row]; rid_hi = p.id_hi[row]; }
    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const uint32_t tphase = (it >> 1) & 1;
      const long long col0 = (t_begin + it) * 128 + half * 64;
      mbar_wait(&t_full[ab * 2 + buf], tphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ab * 2 + buf) * 128 + half * 64);
      uint32_t r[64];
      tmem_ld64(taddr, r);
      tmem_ld_wait64(r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[ab * 2 + buf]);
      const int n_valid = (int)min(64ll, p.C - col0);  // columns beyond C are zero-padded rows of the image
      if (n_valid <= 0) continue;
      // BIAS (sampling-probability correction, retrieval.py:190-192): logits s/T + b_j.  The scores are turned into
      // log2-unit logits in place (one FFMA with the per-candidate bias, read through L2), the rest runs with scale 1.
      if (BIAS) {
        const float4* b4 = reinterpret_cast<const float4*>(p.cbias2 + col0);
#pragma unroll
        for (int j4 = 0; j4 < 16; ++j4) {
          const float4 bb = __ldg(b4 + j4);
          r[4 * j4 + 0] = __float_as_uint(fmaf(__uint_as_float(r[4 * j4 + 0]), scale2, bb.x));
          r[4 * j4 + 1] = __float_as_uint(fmaf(__uint_as_float(r[4 * j4 + 1]), scale2, bb.y));
          r[4 * j4 + 2] = __float_as_uint(fmaf(__uint_as_float(r[4 * j4 + 2]), scale2, bb.z));
          r[4 * j4 + 3] = __float_as_uint(fmaf(__uint_as_float(r[4 * j4 + 3]), scale2, bb.w));
        }
      }
      if (EXT) {
        if (p.id_lo) {   // accidental hits: another candidate with the id of this row's positive -> MIN_FLOAT
          const int4* il = reinterpret_cast<const int4*>(p.id_lo + col0);
#pragma unroll
          for (int j4 = 0; j4 < 16; ++j4) {
            const int4 v = __ldg(il + j4);
            if ((v.x == rid_lo) | (v.y == rid_lo) | (v.z == rid_lo) | (v.w == rid_lo)) {   // rare
              const int* ih = p.id_hi + col0 + 4 * j4;
              const long long c = col0 + 4 * j4;
              if (v.x == rid_lo && __ldg(ih + 0) == rid_hi && c + 0 != row) r[4 * j4 + 0] = __float_as_uint(SX_MIN2);
              if (v.y == rid_lo && __ldg(ih + 1) == rid_hi && c + 1 != row) r[4 * j4 + 1] = __float_as_uint(SX_MIN2);
              if (v.z == rid_lo && __ldg(ih + 2) == rid_hi && c + 2 != row) r[4 * j4 + 2] = __float_as_uint(SX_MIN2);
              if (v.w == rid_lo && __ldg(ih + 3) == rid_hi && c + 3 != row) r[4 * j4 + 3] = __float_as_uint(SX_MIN2);
            }
          }
        }
        if (p.mbits) {   // score_mask: bit = keep
          const uint2 mw = __ldg(reinterpret_cast<const uint2*>(p.mbits + row * p.mwords + (t_begin + it) * 4 + half * 2));
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (!((mw.x >> j) & 1u)) r[j] = __float_as_uint(SX_MIN2);
            if (!((mw.y >> j) & 1u)) r[32 + j] = __float_as_uint(SX_MIN2);
          }
        }
      }
      const float sc = BIAS ? 1.0f : scale2;
      const long long blk0 = (long long)qb * 256 + ab * 128;
      const bool edge = n_valid < 64 || (blk0 < col0 + 64 && col0 < blk0 + 128);  // ragged tail, or the tile with the positives
      float m_new, acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      if (!edge) {
        // scale2 > 0: the row maximum can be taken on the raw accumulators (FMNMX3 tree), one FFMA + one MUFU per score
        float t[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float a0 = max3(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]), __uint_as_float(r[8 * g + 2]));
          const float a1 = max3(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
          t[g] = max3(a0, a1, fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
        }
        const float tmax = max3(max3(t[0], t[1], t[2]), max3(t[3], t[4], t[5]), fmaxf(t[6], t[7]));
        m_new = fmaxf(m2, tmax * sc);
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          acc0 += ex2_approx(fmaf(__uint_as_float(r[j]), sc, -m_new));
          acc1 += ex2_approx(fmaf(__uint_as_float(r[j + 1]), sc, -m_new));
          acc2 += ex2_approx(fmaf(__uint_as_float(r[j + 2]), sc, -m_new));
          acc3 += ex2_approx(fmaf(__uint_as_float(r[j + 3]), sc, -m_new));
        }
      } else {
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < n_valid) tmax = fmaxf(tmax, __uint_as_float(r[j]));
        m_new = fmaxf(m2, tmax * sc);
        if (row >= col0 && row < col0 + 64) {   // the positive of query i is candidate i (retrieval.py:185)
          const int jd = (int)(row - col0);
#pragma unroll
          for (int j = 0; j < 64; ++j) if (j == jd) pos2 = __uint_as_float(r[j]) * sc;
        }
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < n_valid) acc0 += ex2_approx(fmaf(__uint_as_float(r[j]), sc, -m_new));
      }
      l = l * ex2_approx(m2 - m_new) + ((acc0 + acc1) + (acc2 + acc3));
      m2 = m_new;
    }
    if (row < p.B) {
      p.partial[(row * p.parts + part) * 2 + half] = make_float2(m2, l);
      // exactly one (part, half) thread of the row saw the diagonal column
      const long long dt = row / 128;
      if (dt >= t_begin && dt < t_end && ((row % 128) / 64) == half) p.pos[row] = pos2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// lse_i = ln2 * (M + log2(sum_p l_p 2^(m_p - M)));  rowloss_i = w_i (lse_i - pos_i), formed as ln2 * ((M - pos2_i) + log2 L)
// so that a row whose logits are ALL MIN_FLOAT (fully masked) still yields log(C), as the max-subtracted reference does
__global__ void __launch_bounds__(256)
smtc_combine_kernel(const float2* __restrict__ partial, int n_partials, const float* __restrict__ pos,
                    const float* __restrict__ w, long long B, float* __restrict__ lse, float* __restrict__ rowloss) {
  const long long row = (long long)blockIdx.x * 256 + threadIdx.x;
  if (row >= B) return;
  const float2* pp = partial + row * n_partials;
  float M = -INFINITY;
  for (int i = 0; i < n_partials; ++i) M = fmaxf(M, pp[i].x);
  float L = 0.f;
  for (int i = 0; i < n_partials; ++i) L += pp[i].y * exp2f(pp[i].x - M);
  const float lg = log2f(L);
  lse[row] = (M + lg) * 0.6931471805599453f;
  rowloss[row] = (w ? w[row] : 1.0f) * (((M - pos[row]) + lg) * 0.6931471805599453f);
}

__global__ void __launch_bounds__(1024) smtc_reduce_loss(const float* __restrict__ rowloss, long long B, float* __restrict__ loss) {
  __shared__ double red[1024];
  double a = 0.0;
  for (long long i = threadIdx.x; i < B; i += 1024) a += (double)rowloss[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) { if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
  if (threadIdx.x == 0) loss[0] = (float)red[0];
}

struct SxPlan { int kb, nqb, parts; long long Bp, n_ctiles, idpad; size_t smem, o_qst, o_cst, o_qimg, o_cimg, o_partial, o_pos, o_rowloss, o_bias, o_idlo, o_idhi, o_mbits, total; };

static bool sx_plan(long long B, long long C, int d, SxPlan& pl, bool has_ids = false, bool has_mask = false) {
  if (B <= 0 || C < B || d <= 0 || d > 128) return false;
  pl.kb = (int)ceil_div(d, 64);
  pl.nqb = (int)ceil_div(B, 256);
  pl.Bp = (long long)pl.nqb * 256;
  pl.n_ctiles = ceil_div(C, 128);
  // candidate parts: minimise waves x (tiles per CTA + ~6 tile times of fixed cost: A load, pipeline fill, partials)
  int parts = 1; double best = 1e30;
  const int sms = sm_count();
  for (int c = 1; c <= 16 && c <= pl.n_ctiles; ++c) {
    const double cost = (double)ceil_div((long long)pl.nqb * c, sms) * ((double)ceil_div(pl.n_ctiles, c) + 6.0);
    if (cost < best * 0.97) { best = cost; parts = c; }
  }
  pl.parts = parts;
  pl.smem = (size_t)(2 + sx_stages(pl.kb)) * pl.kb * 32768 + 1024 + 256;
  if (pl.smem > 227 * 1024) return false;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 1024); return r; };
  pl.o_qst = take(sizeof(CxStats)); pl.o_cst = take(sizeof(CxStats));
  pl.o_qimg = take(cx_img_bytes(pl.Bp, d));
  pl.o_cimg = take(cx_img_bytes(pl.n_ctiles * 128, d));
  pl.o_partial = take((size_t)pl.Bp * parts * 2 * sizeof(float2));
  pl.o_pos = take((size_t)pl.Bp * 4);
  pl.o_rowloss = take((size_t)pl.Bp * 4);
  pl.o_bias = take((size_t)pl.n_ctiles * 128 * 4);
  pl.idpad = pl.Bp > pl.n_ctiles * 128 ? pl.Bp : pl.n_ctiles * 128;
  pl.o_idlo = take(has_ids ? (size_t)pl.idpad * 4 : 0);
  pl.o_idhi = take(has_ids ? (size_t)pl.idpad * 4 : 0);
  pl.o_mbits = take(has_mask ? (size_t)pl.Bp * pl.n_ctiles * 4 * 4 : 0);
  pl.total = o;
  return true;
}

}  // namespace tc
}  // namespace tfrs
using namespace tfrs;
using namespace tfrs::tc;

extern "C" size_t tfrs_inbatch_softmax_tc_workspace_bytes(int64_t B, int64_t C, int d) {
  SxPlan pl;
  return sx_plan(B, C, d, pl) ? pl.total : 0;
}

// cbias2[i] = bias[i] * log2(e) for i < C, 0 on the padding (and everywhere when bias == NULL)
__global__ void __launch_bounds__(256)
smtc_bias_kernel(const float* __restrict__ bias, long long C, long long Cpad, float* __restrict__ cbias2) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < Cpad) cbias2[i] = (bias && i < C) ? bias[i] * SX_LOG2E : 0.f;
}

extern "C" size_t tfrs_inbatch_softmax_tc_ex_workspace_bytes(int64_t B, int64_t C, int d, int has_ids, int has_mask) {
  SxPlan pl;
  return sx_plan(B, C, d, pl, has_ids != 0, has_mask != 0) ? pl.total : 0;
}

extern "C" int tfrs_inbatch_softmax_tc_fwd_ex(const float* q, const float* c, int64_t B, int64_t C, int d, float inv_temperature,
                                              const float* sample_weight, const float* candidate_bias,
                                              const int64_t* candidate_ids, const uint8_t* score_mask, float* loss, float* lse,
                                              void* ws, size_t ws_bytes, void* stream) {
  TFRS_CHECK_ARG(q && c && loss && lse, "inbatch_softmax_tc_fwd: NULL pointer");
  SxPlan pl;
  const bool ext = candidate_ids || score_mask;
  if (!(inv_temperature > 0.f)) { set_error("inbatch_softmax_tc_fwd: needs a positive temperature"); return TFRS_ERR_UNSUPPORTED; }
  if (!sx_plan(B, C, d, pl, candidate_ids != nullptr, score_mask != nullptr)) { set_error("inbatch_softmax_tc_fwd: shape outside the tensor-core path (need B <= C, d <= 128)"); return TFRS_ERR_UNSUPPORTED; }
  if (!ws || ws_bytes < pl.total) { set_error("inbatch_softmax_tc_fwd: workspace too small"); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  TFRS_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "inbatch_softmax_tc_fwd: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)ws;
  CxStats* qst = (CxStats*)(w + pl.o_qst); CxStats* cst = (CxStats*)(w + pl.o_cst);
  unsigned char* qimg = w + pl.o_qimg; unsigned char* cimg = w + pl.o_cimg;
  float2* partial = (float2*)(w + pl.o_partial);
  float* pos = (float*)(w + pl.o_pos); float* rowloss = (float*)(w + pl.o_rowloss);
  TFRS_CUDA(cudaMemsetAsync(w, 0, 2048, st));  // both stats blocks
  cx_amax_kernel<<<cx_amax_grid(B), 256, 0, st>>>(q, B, d, d, qst);
  TFRS_LAUNCH_CHECK();
  cx_amax_kernel<<<cx_amax_grid(C), 256, 0, st>>>(c, C, d, d, cst);
  TFRS_LAUNCH_CHECK();
  cx_exp_kernel<<<1, 1, 0, st>>>(qst);
  TFRS_LAUNCH_CHECK();
  cx_exp_kernel<<<1, 1, 0, st>>>(cst);
  TFRS_LAUNCH_CHECK();
  {
    long long chunks = pl.Bp / 128 * 128 * (long long)pl.kb * 8;
    cx_split_image_kernel<false><<<(unsigned)ceil_div(chunks, 256), 256, 0, st>>>(q, B, d, d, pl.kb, pl.Bp / 128, qst, qimg);
    TFRS_LAUNCH_CHECK();
    chunks = pl.n_ctiles * 128 * (long long)pl.kb * 8;
    cx_split_image_kernel<false><<<(unsigned)ceil_div(chunks, 256), 256, 0, st>>>(c, C, d, d, pl.kb, pl.n_ctiles, cst, cimg);
    TFRS_LAUNCH_CHECK();
  }
  SoftmaxTcParams p{};
  p.qimg = qimg; p.cimg = cimg; p.qst = qst; p.cst = cst; p.B = B; p.C = C; p.nqb = pl.nqb; p.parts = pl.parts; p.kb = pl.kb;
  p.n_ctiles = pl.n_ctiles; p.inv_t = inv_temperature; p.partial = partial; p.pos = pos;
  const int s1 = (int)((2 + sx_stages(1)) * 32768 + 1280), s2 = (int)((2 + sx_stages(2)) * 2 * 32768 + 1280);
  TFRS_DYN_SMEM((softmax_tc_kernel<1, 0>), s1);
  TFRS_DYN_SMEM((softmax_tc_kernel<2, 0>), s2);
  TFRS_DYN_SMEM((softmax_tc_kernel<1, 1>), s1);
  TFRS_DYN_SMEM((softmax_tc_kernel<2, 1>), s2);
  TFRS_DYN_SMEM((softmax_tc_kernel<1, 2>), s1);
  TFRS_DYN_SMEM((softmax_tc_kernel<2, 2>), s2);
  const unsigned grid = (unsigned)(pl.nqb * pl.parts);
  if (candidate_bias || ext) {
    float* cb2 = (float*)(w + pl.o_bias);
    smtc_bias_kernel<<<(unsigned)ceil_div(pl.n_ctiles * 128, 256), 256, 0, st>>>(candidate_bias, C, pl.n_ctiles * 128, cb2);
    TFRS_LAUNCH_CHECK();
    p.cbias2 = cb2;
  }
  if (candidate_ids) {
    int* lo = (int*)(w + pl.o_idlo); int* hi = (int*)(w + pl.o_idhi);
    sx_ids_split_kernel<<<(unsigned)ceil_div(pl.idpad, 256), 256, 0, st>>>((const long long*)candidate_ids, C, pl.idpad, lo, hi);
    TFRS_LAUNCH_CHECK();
    p.id_lo = lo; p.id_hi = hi;
  }
  if (score_mask) {
    uint32_t* mb = (uint32_t*)(w + pl.o_mbits);
    p.mwords = (int)(pl.n_ctiles * 4);
    sx_mask_pack_kernel<<<(unsigned)ceil_div(pl.Bp * p.mwords, 256), 256, 0, st>>>(score_mask, B, C, pl.Bp, p.mwords, mb);
    TFRS_LAUNCH_CHECK();
    p.mbits = mb;
  }
  if (ext) {
    if (pl.kb == 1) softmax_tc_kernel<1, 2><<<grid, SX_THREADS, pl.smem, st>>>(p);
    else softmax_tc_kernel<2, 2><<<grid, SX_THREADS, pl.smem, st>>>(p);
  } else if (candidate_bias) {
    if (pl.kb == 1) softmax_tc_kernel<1, 1><<<grid, SX_THREADS, pl.smem, st>>>(p);
    else softmax_tc_kernel<2, 1><<<grid, SX_THREADS, pl.smem, st>>>(p);
  } else {
    if (pl.kb == 1) softmax_tc_kernel<1, 0><<<grid, SX_THREADS, pl.smem, st>>>(p);
    else softmax_tc_kernel<2, 0><<<grid, SX_THREADS, pl.smem, st>>>(p);
  }
  TFRS_LAUNCH_CHECK();
  smtc_combine_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(partial, pl.parts * 2, pos, sample_weight, B, lse, rowloss);
  TFRS_LAUNCH_CHECK();
  smtc_reduce_loss<<<1, 1024, 0, st>>>(rowloss, B, loss);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

extern "C" int tfrs_inbatch_softmax_tc_fwd(const float* q, const float* c, int64_t B, int64_t C, int d, float inv_temperature,
                                           const float* sample_weight, const float* candidate_bias, float* loss, float* lse,
                                           void* ws, size_t ws_bytes, void* stream) {
  return tfrs_inbatch_softmax_tc_fwd_ex(q, c, B, C, d, inv_temperature, sample_weight, candidate_bias, nullptr, nullptr, loss, lse,
                                        ws, ws_bytes, stream);
}
