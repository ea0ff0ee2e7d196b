// topk_tc.cu -- K2: brute-force top-K on the 5th-gen tensor cores (tcgen05 + TMEM + bulk-TMA), sm_100a.
//
// Replaces  scores = matmul(q, c^T); top_k(scores, k)  (layers/factorized_top_k.py:603-605) for large
// corpora.  The [Q,N] score matrix never exists in HBM:
//
//   index time   tfrs_index_build : corpus fp32 -> fp16 image (exact 2^e rescale), pre-tiled as 128-row UMMA SWIZZLE_128B
//                K-major tiles (one contiguous 16 KB block per 64-wide K slab), + max row norm (scaled units).
//   query time   (0) tc_qprep     : ONE kernel, a warp per query: per-ROW power-of-two rescale (scores of different
//                    queries are never compared, so every query gets its own exponent), fp16 tile image, error margins
//                (1) tc_scan<SAMPLE> : screening GEMM over every 4th corpus tile; each epilogue thread keeps the maximum
//                    of a GROUP of tiles (a "bin") -> <= 1024 bins per query; tc_threshold (a warp per query, bins in
//                    registers) takes the K-th largest bin maximum L_q: K distinct bins hold K distinct candidates
//                    >= L_q, so L_q is a valid lower bound of the K-th best screening score
//                (2) tc_scan<FILTER> : screening GEMM over the whole corpus; epilogue compares the fp32
//                    accumulators (read from TMEM) with T_q = L_q - margin_q and appends the rare
//                    survivors (octet records) to per-(query, part, half) lists -- nothing else leaves the SM
//                (3) tc_finalize (a warp per query, no block barriers): tau = K-th best screening score; survivors
//                    within the error band of tau are re-scored EXACTLY (sequential fp32 fmaf chain on the fp32
//                    corpus) and ranked by (score desc, index asc) -> bit-identical to the exact CUDA-core path.
//                    Variants of the same kernel: EXCLUDE (query_with_exclusions, :83-115,242-288: the k+E best are
//                    re-ranked with excluded identifiers lowered by 1e5) and COUNT (the FactorizedTopK metric,
//                    metrics/factorized_top_k.py:133-192: #{candidates scoring above the positive}, no top-K list).
//                (4) overflow fallback (list capacity exceeded; adversarial inputs only): exact scan.
//
// Why the result is exact: |screen(q,c) - exact(q,c)| <= eps_q = E_REL*|q|*max|c| (fp16 rounding of both
// operands: (2u+u^2) sum|q_k c_k| with u = 2^-11, plus accumulation slack; Cauchy-Schwarz).  Any member
// of the exact top-K has screening score >= tau - 2*eps_q >= L_q - 2*eps_q, so it is in the list and in
// the re-scored band.
//
// Scan kernel shape (per CTA, 1 CTA / SM, 640 threads): 256 queries (two 128-row A blocks, resident in smem)
// x a contiguous range of 128-row corpus tiles streamed through a 4-6 stage bulk-TMA ring; warp 0 = TMA
// producer, warp 1 = MMA issuer (one thread, tcgen05.mma M=128 N=128 K=16, fp16 -> fp32 in TMEM),
// warp 2 = TMEM allocator, warps 4-19 = epilogue (one query row x 64 columns per thread, tcgen05.ld 32x32b.x64).
// TMEM holds 2 A-blocks x 2 buffers x 128 columns = all 512 columns, so tile t+1's MMAs overlap tile
// t's epilogue.  Each B tile feeds two MMAs (both A blocks): 16 KB of L2->smem traffic per 512
// tensor-core cycles keeps the chip under the ~6.3 KB/clk L2 fabric limit.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include "rowselect.cuh"
#include "tc_ptx.cuh"

namespace tfrs {
namespace tc {

constexpr int TILE_N = 128;          // corpus rows per B tile
constexpr int TILE_M = 128;          // query rows per A block
constexpr int QBLK = 256;            // queries per CTA (2 A blocks)
constexpr int KSLAB = 64;            // fp16 elements per 128-byte swizzle row
constexpr int SLAB_BYTES = TILE_N * 128;  // 16 KB: 128 rows x 128 B
constexpr int HEADER_BYTES = 1024;
constexpr int THREADS = 640;      // 4 control warps + 16 epilogue warps (4 per SM sub-partition)
constexpr int EPI_WARPS = 16;
constexpr int EPI_WARP0 = 4;
constexpr int CAND_CAP = 2048;       // octet records kept per query across all segments (sizing of cap_part)
constexpr int MAX_SAMPLE_STRIDE = 4;
constexpr int FIN_MAX_PARTS = 2 * 148;  // survivor-list segments per query: (corpus part, column half)
constexpr int MAX_BINS = 1024;       // bin maxima per query (32 per lane of the threshold warp)
constexpr int FP16_TARGET = 15;      // largest operand magnitude lands in [2^14, 2^15)
// Screening error model (operands: fp16 after an exact power-of-two rescale -- one exponent for the whole corpus,
// one per query row -- so that the largest magnitude lands in [2^14, 2^15); accumulate: fp32 in TMEM):
//   |x^ - x| <= 2^-11 |x| (+2^-25 absolute below the fp16 normal range, negligible after the rescale)
//   => |screen - exact| <= (2^-10 + 2^-22) sum|q_k c_k|  + accumulation slack (budget 2^-14) + fmaf-chain 2^-16
//   <= E_REL * |q| * |c|   (Cauchy-Schwarz), E_REL = 0.00108 including the 0.1 % norm inflation.
constexpr float E_REL = 0.00108f;
constexpr float E_ACC = 0.00013f;    // run-to-run slack between the two passes (they are bit-identical in practice)

struct SideStats {              // corpus side, written at index time
  unsigned int max_norm2_bits;  // max_i |x_i * 2^exp|^2  (SCALED units; float bits; non-negative so uint order == float order)
  unsigned int amax_bits;       // max_ij |x_ij|
  int exp;                      // rescale exponent e: x * 2^e has its largest magnitude in [2^14, 2^15)
  int pad;
};
struct IndexHeader {
  SideStats st;
  int d, d_pad, kb, pad;
  long long n, n_tiles;
};

__device__ __forceinline__ int rescale_exp(float amax) {
  int x = 0;
  if (!(amax > 0.f) || !(amax < INFINITY)) return 0;
  (void)frexpf(amax, &x);      // amax = m * 2^x, m in [0.5, 1)
  return FP16_TARGET - x;      // amax * 2^exp in [2^(target-1), 2^target)
}

// ------------------------------------------------------------------------------------------------
// image builders: fp32 [rows, d] -> fp16 128-row tiles, each K slab of 64 as one swizzled 16 KB block
//   byte offset of element (r, k) inside a tile = (k/64)*16384 + r*128 + (((k%64)/8) ^ (r%8))*16 + (k%8)*2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_image_kernel(const float* __restrict__ src, long long rows, int d, int kb, long long n_tiles,
                  const SideStats* __restrict__ st, unsigned char* __restrict__ img) {
  const int scale_exp = st->exp;
  const long long total = n_tiles * TILE_N * (long long)kb * 8;  // 16-byte chunks
  for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
    int chunk = (int)(e % (kb * 8));
    long long row = e / (kb * 8);
    int slab = chunk / 8, cj = chunk % 8;
    int r = (int)(row % TILE_N);
    long long tile = row / TILE_N;
    int k0 = slab * KSLAB + cj * 8;
    __align__(16) __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f = (row < rows && k0 + j < d) ? src[row * d + k0 + j] : 0.f;
      v[j] = __float2half_rn(ldexpf(f, scale_exp));  // exact power-of-two rescale, then one rounding to fp16
    }
    unsigned char* dst = img + tile * ((long long)kb * SLAB_BYTES) + (long long)slab * SLAB_BYTES + r * 128 + ((cj ^ (r & 7)) * 16);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
  }
}

// corpus statistics, two passes (one warp per row -> coalesced): PASS 0 max |element| -> exponent; PASS 1 max row norm^2
// of the SCALED rows (so tiny or huge corpora neither underflow nor overflow the fp32 norm)
template <int PASS>
__global__ void __launch_bounds__(256)
corpus_stats_kernel(const float* __restrict__ src, long long rows, int d, SideStats* __restrict__ st) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * 256) >> 5;
  const int e = PASS ? st->exp : 0;
  float best = 0.f;
  for (long long row = warp; row < rows; row += nwarps) {
    const float* p = src + row * d;
    float acc = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float x = p[k];
      if (PASS) { const float y = ldexpf(x, e); acc = fmaf(y, y, acc); } else acc = fmaxf(acc, fabsf(x));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, acc, o);
      acc = PASS ? acc + other : fmaxf(acc, other);
    }
    best = fmaxf(best, acc);
  }
  if (lane == 0 && best > 0.f) {
    if (PASS) atomicMax(&st->max_norm2_bits, __float_as_uint(best * 1.0001f));  // slack for the tree order
    else atomicMax(&st->amax_bits, __float_as_uint(best));
  }
}
__global__ void side_exp_kernel(SideStats* st) { st->exp = rescale_exp(__uint_as_float(st->amax_bits)); }
__global__ void header_kernel(IndexHeader* dst, IndexHeader h) { *dst = h; }

// (0) query preparation, one warp per (padded) query row: per-row exponent, scaled norm -> margins, fp16 tile image.
//   margin[row] = 2*eps + run-to-run slack  (filter threshold T = L - margin)
//   cut[row]    = 2*eps                      (band below tau that is re-scored exactly)
//   qexp[row]   = the row's exponent (COUNT mode converts the positive score to screening units with it)
// All in SCREENING units: scores scaled by 2^(exp_corpus + exp_row), an exact power of two.
__global__ void __launch_bounds__(256)
tc_qprep_kernel(const float* __restrict__ q, long long Q, long long Qp, int d, int kb, const IndexHeader* __restrict__ hdr,
                unsigned char* __restrict__ qimg, float* __restrict__ margin, float* __restrict__ cut, int* __restrict__ qexp) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (row >= Qp) return;
  const bool real = row < Q;
  const float* p = q + row * d;
  float a = 0.f;
  if (real) for (int k = lane; k < d; k += 32) a = fmaxf(a, fabsf(p[k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o));
  const int e = rescale_exp(a);
  float n2 = 0.f;
  if (real) for (int k = lane; k < d; k += 32) { const float y = ldexpf(p[k], e); n2 = fmaf(y, y, n2); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  if (lane == 0) {
    const float cn = sqrtf(__uint_as_float(hdr->st.max_norm2_bits)) * 1.001f;   // scaled corpus norm bound
    const float qn = sqrtf(n2 * 1.0001f) * 1.001f;                               // scaled query norm (+ tree-order slack)
    const float eps = E_REL * qn * cn + 1e-30f;
    margin[row] = 2.f * eps + E_ACC * qn * cn;
    cut[row] = 2.f * eps;
    qexp[row] = e;
  }
  // image: chunk (slab, cj) of this row; kb*8 <= 16 chunks, one lane each
  const int r = (int)(row % TILE_N);
  const long long tile = row / TILE_N;
  if (lane < kb * 8) {
    const int slab = lane >> 3, cj = lane & 7;
    const int k0 = slab * KSLAB + cj * 8;
    __align__(16) __half v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float f = (real && k0 + j < d) ? p[k0 + j] : 0.f;
      v[j] = __float2half_rn(ldexpf(f, e));
    }
    unsigned char* dst = qimg + tile * ((long long)kb * SLAB_BYTES) + (long long)slab * SLAB_BYTES + r * 128 + ((cj ^ (r & 7)) * 16);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
  }
}

// ------------------------------------------------------------------------------------------------
// the screening GEMM
// ------------------------------------------------------------------------------------------------
enum { MODE_SAMPLE = 0, MODE_FILTER = 1 };

struct ScanParams {
  const unsigned char* qimg;    // query tile image  [2*nqb tiles][KB][16 KB]
  const unsigned char* cimg;    // corpus tile image [n_tiles][KB][16 KB]
  long long Q, N;
  int nqb, parts, n_seq, stride;  // tile sequence: tile(u) = u * stride, u in [0, n_seq)
  long long n_tiles;
  // SAMPLE: a bin = the maximum over `group` consecutive sampled tiles x 64 columns of one epilogue thread
  float* binmax; int bins_ld;     // [Qp, bins_ld]; bin = (part * bins_per_part + it / group) * 2 + half
  int group, bins_per_part;
  // FILTER
  const float* thr;               // [Qp]
  unsigned int* count;            // [Qp, parts, 2]   records written by each (query, corpus part, half)
  // Survivor RECORDS: when any of 8 consecutive columns of a row passes the threshold, the whole octet is
  // appended (two 16-byte stores + the index of its first column); finalize drops the non-survivors.
  // One private segment per (query row, corpus part, column half): a single writer thread, no atomics.
  float* cand_s;                  // [Qp, parts, 2, cap_part, 8] screening scores of the octet
  unsigned int* cand_i;           // [Qp, parts, 2, cap_part]    local index of the octet's first column
  int cap_part;                   // records per segment
  uint32_t idesc;
};

template <int KB, int STAGES, int MODE>
__global__ void __launch_bounds__(THREADS, 1)
tc_scan_kernel(const ScanParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: [A: 2*KB slabs][B: STAGES*KB slabs][barriers]
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + 2 * KB * SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * KB * SLAB_BYTES);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* a_full = bars + 2 * STAGES;  // [1]
  uint64_t* t_full = a_full + 1;         // [2]
  uint64_t* t_empty = t_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = blockIdx.x % p.nqb, part = blockIdx.x / p.nqb;
  const int u_begin = (int)((long long)part * p.n_seq / p.parts);
  const int u_end = (int)((long long)(part + 1) * p.n_seq / p.parts);
  const int n_iter = u_end - u_begin;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(a_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== bulk-TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(a_full, 2 * KB * SLAB_BYTES);
      bulk_g2s(sA, p.qimg + (long long)qb * 2 * KB * SLAB_BYTES, 2 * KB * SLAB_BYTES, a_full);
      int stage = 0; uint32_t phase = 0;
      const uint64_t pol = l2_policy_evict_first();
      for (int it = 0; it < n_iter; ++it) {
        const long long tile = (long long)(u_begin + it) * p.stride;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], KB * SLAB_BYTES);
        if (MODE == MODE_FILTER)   // streamed once per pass: evict-first, so the survivor records stay in L2 for the select kernel
          bulk_g2s_hint(sB + stage * KB * SLAB_BYTES, p.cimg + tile * ((long long)KB * SLAB_BYTES), KB * SLAB_BYTES, &full[stage], pol);
        else
          bulk_g2s(sB + stage * KB * SLAB_BYTES, p.cimg + tile * ((long long)KB * SLAB_BYTES), KB * SLAB_BYTES, &full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      mbar_wait(a_full, 0);
      tc_fence_after();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        const uint32_t tphase = (it >> 1) & 1;
        mbar_wait(&t_empty[buf], tphase ^ 1);   // epilogue drained this accumulator buffer
        mbar_wait(&full[stage], phase);         // B tile landed
        tc_fence_after();
#pragma unroll
        for (int ab = 0; ab < 2; ++ab) {
          const uint32_t d_tmem = tmem_base + (uint32_t)((ab * 2 + buf) * TILE_N);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t a_desc = make_smem_desc(smem_u32(sA + (ab * KB + kb) * SLAB_BYTES));
            const uint64_t b_desc = make_smem_desc(smem_u32(sB + (stage * KB + kb) * SLAB_BYTES));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)  // 4 x (K=16 fp16 = 32 B) inside the 128-byte swizzle row
              umma_f16(d_tmem, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), p.idesc,
                        (uint32_t)((kb | k4) != 0));
          }
        }
        umma_commit(&empty[stage]);   // smem slot free once these MMAs retire
        umma_commit(&t_full[buf]);    // accumulators ready for the epilogue
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue: one query row per thread =====
    // 16 epilogue warps: (column half, A block, TMEM lane quadrant); a thread owns one query row x 64 columns of
    // every tile, so four warps per scheduler overlap their TMEM-load latency with each other's math.
    const int ew = warp - EPI_WARP0;       // 0..15
    const int half = ew >> 3, ab = (ew >> 2) & 1, quad = ew & 3;  // TMEM lane quadrant == warp % 4
    const long long row = (long long)qb * QBLK + ab * TILE_M + quad * 32 + lane;
    const bool row_ok = row < p.Q;
    float thr = INFINITY;
    if (MODE == MODE_FILTER && row_ok) thr = p.thr[row];
    float* my_s = nullptr; unsigned int* my_i = nullptr;
    unsigned int my_cnt = 0, my_ovf = 0;
    const unsigned int cap = (unsigned int)p.cap_part;
    if (MODE == MODE_FILTER) {
      const long long seg = (((long long)row * p.parts + part) * 2 + half) * p.cap_part;
      my_s = p.cand_s + seg * 8; my_i = p.cand_i + seg;
    }
    float binm = -INFINITY;
    int in_group = 0, bin_out = 0;
    float* my_bins = nullptr;
    if (MODE == MODE_SAMPLE) my_bins = p.binmax + row * p.bins_ld + (long long)part * p.bins_per_part * 2 + half;
    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const uint32_t tphase = (it >> 1) & 1;
      const int u = u_begin + it;
      const long long tile = (long long)u * p.stride;
      const long long col0 = tile * TILE_N;  // zero-padded rows of the last tile score 0: dropped in finalize (idx >= N)
      mbar_wait(&t_full[buf], tphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ab * 2 + buf) * TILE_N);
      {
        const int h = half;
        uint32_t r[64];
        tmem_ld64(taddr + h * 64, r);
        tmem_ld_wait64(r);
        // this warp's TMEM reads of the accumulator buffer are complete: release it before the math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[buf]);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[c2 * 32 + j]);
          // 4 quarter maxima of 8, then their max (FMNMX3 trees)
          float qmx[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            qmx[i] = max3(max3(v[8 * i], v[8 * i + 1], v[8 * i + 2]), max3(v[8 * i + 3], v[8 * i + 4], v[8 * i + 5]),
                          fmaxf(v[8 * i + 6], v[8 * i + 7]));
          const float m = fmaxf(max3(qmx[0], qmx[1], qmx[2]), qmx[3]);
          if (MODE != MODE_FILTER) {
            binm = fmaxf(binm, m);
          } else {
            // Survivors are rare.  Every branch below is WARP-UNIFORM: one vote on the chunk max, then one
            // REDUX.OR of the per-lane 4-bit quarter mask; the per-lane work is predicated stores only.
            if (__any_sync(0xffffffffu, m >= thr)) {
              unsigned int qmask = 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) qmask |= (qmx[i] >= thr) ? (1u << i) : 0u;
              const unsigned int umask = __reduce_or_sync(0xffffffffu, qmask);
              const bool room = my_cnt + 4u <= cap;    // worst case of this visit (4 octets) fits
              my_ovf |= (!room && m >= thr) ? 1u : 0u;
              const unsigned int idx0 = (unsigned int)(col0 + h * 64 + c2 * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                if (umask & (1u << i)) {             // uniform: some lane's octet i has a survivor
                  if (room && qmx[i] >= thr) {       // per lane: append my whole octet (predicated, no loop)
                    float4* dst = reinterpret_cast<float4*>(my_s + (size_t)my_cnt * 8);
                    dst[0] = make_float4(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3]);
                    dst[1] = make_float4(v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
                    my_i[my_cnt] = idx0 + 8 * i;
                    ++my_cnt;
                  }
                }
              }
            }
          }
        }
      }
      if (MODE == MODE_SAMPLE) {   // close the bin after `group` tiles (loop-uniform)
        if (++in_group == p.group || it == n_iter - 1) {
          if (row_ok) my_bins[2 * bin_out] = binm;
          ++bin_out; in_group = 0; binm = -INFINITY;
        }
      }
    }
    if (MODE == MODE_SAMPLE) {
      if (row_ok) for (; bin_out < p.bins_per_part; ++bin_out) my_bins[2 * bin_out] = -INFINITY;  // bins this part did not fill
    } else {
      p.count[((long long)row * p.parts + part) * 2 + half] = my_ovf ? (cap + 1u) : my_cnt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// threshold from the bin maxima; finalize; fallback
// ------------------------------------------------------------------------------------------------
// orderable key: larger float <=> larger unsigned (NaN sorts above +inf; -0 < +0, callers canonicalise zeros)
__device__ __forceinline__ unsigned int f2key(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
  return v;
}

// k-th largest of the keys a warp holds in registers (KPL per lane, unused slots = 0 = below every float key):
// bitwise binary search for the largest v with #{key >= v} >= k, one REDUX.SUM per bit.
template <int KPL>
__device__ __forceinline__ unsigned int warp_kth_largest_regs(const unsigned int (&key)[KPL], int k, unsigned int prefix = 0, int top_bit = 31) {
  // `prefix` = the bits above top_bit, known to be shared by every real key (unused slots are 0 and never reach it)
  for (int bit = top_bit; bit >= 0; --bit) {
    const unsigned int cand = prefix | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) c += (key[j] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) prefix = cand;
  }
  return prefix;
}
// the same over n keys in shared memory
__device__ __forceinline__ unsigned int warp_kth_largest_smem(const unsigned int* keys, int n, int k, int lane, unsigned int prefix, int top_bit) {
  for (int bit = top_bit; bit >= 0; --bit) {
    const unsigned int cand = prefix | (1u << bit);
    int c = 0;
    for (int t = lane; t < n; t += 32) c += (keys[t] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) prefix = cand;
  }
  return prefix;
}

// (1b) K-th largest bin maximum of the sampled pass -> filter threshold T = L - margin.  One warp per query, the
// query's <= 32*KPL bin maxima live in registers.
template <int KPL>
__global__ void __launch_bounds__(256)
tc_threshold_kernel(const float* __restrict__ binmax, int bins_ld, int n_bins, int k, const float* __restrict__ margin,
                    float* __restrict__ thr, unsigned int* __restrict__ overflow, long long Q) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (row >= Q) return;
  const float* src = binmax + row * bins_ld;
  unsigned int key[KPL];
#pragma unroll
  for (int j = 0; j < KPL; ++j) { const int i = j * 32 + lane; key[j] = i < n_bins ? f2key(__ldg(src + i)) : 0u; }
  const unsigned int kth = warp_kth_largest_regs<KPL>(key, k);
  if (lane == 0) { thr[row] = key2f(kth) - margin[row]; overflow[row] = 0; }
}

// (3) finalize: one WARP per query, no block-wide barriers.
enum { FIN_TOPK = 0, FIN_EXCLUDE = 1, FIN_COUNT = 2 };
constexpr int FW_WARPS = 4;                       // queries per CTA of the fallback re-rank kernel
// Capacities per query come from the plan (they grow with k); a row that overflows them takes the exact fallback.
// overflow[row]: 0 = done, 1 = exact fallback

struct FinParams {
  const float* q; const float* corpus; int d; int k; long long index_offset; long long N; long long Q;
  const unsigned int* count; const float* cand_s; const unsigned int* cand_i; int segs; int cap_part;
  const float* cut; const float* thr; unsigned int* overflow;
  int cap_keys, cap_band;                        // survivors per query (keys in shared memory) / band entries re-scored exactly
  unsigned int* band_idx; int* band_n;           // [Qp, cap_band] local indices of the band, [Qp] their number
  int allow_short;                               // sharded scan with a GLOBAL threshold: a shard may hold fewer than k survivors
  float* out_s; long long* out_i;                // TOPK: [Q, k];  EXCLUDE: [Q, k_out]
  // EXCLUDE (k = k_out + n_excl candidates are fetched, then re-ranked)
  const long long* identifiers; const long long* exclusions; int n_excl; int k_out;
  // COUNT
  const float* pos; const int* qexp; const IndexHeader* hdr; int* out_count;
};

// exact score: the canonical sequential fmaf chain on the fp32 corpus row (bit-identical to the oracle)
__device__ __forceinline__ float exact_score(const float* __restrict__ qs, const float* __restrict__ c, int d) {
  float acc = 0.f;
  if ((d & 31) == 0) {  // 8 loads (128 B) in flight per step, then the canonical chain on them
    for (int kk = 0; kk < d; kk += 32) {
      float4 cv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) cv[u] = __ldg(reinterpret_cast<const float4*>(c + kk) + u);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc = fmaf(qs[kk + 4 * u], cv[u].x, acc); acc = fmaf(qs[kk + 4 * u + 1], cv[u].y, acc);
        acc = fmaf(qs[kk + 4 * u + 2], cv[u].z, acc); acc = fmaf(qs[kk + 4 * u + 3], cv[u].w, acc);
      }
    }
  } else if ((d & 3) == 0) {
    for (int kk = 0; kk < d; kk += 4) {
      const float4 cv = __ldg(reinterpret_cast<const float4*>(c + kk));
      acc = fmaf(qs[kk], cv.x, acc); acc = fmaf(qs[kk + 1], cv.y, acc);
      acc = fmaf(qs[kk + 2], cv.z, acc); acc = fmaf(qs[kk + 3], cv.w, acc);
    }
  } else {
    for (int kk = 0; kk < d; ++kk) acc = fmaf(qs[kk], __ldg(c + kk), acc);
  }
  return acc + 0.0f;  // -0 -> +0: the key order must agree with the float order
}

// two independent chains at once (same arithmetic per chain as exact_score)
__device__ __forceinline__ void exact_score2(const float* __restrict__ qs, const float* __restrict__ c0, const float* __restrict__ c1,
                                             int d, float& s0, float& s1) {
  if ((d & 31) == 0) {
    float a0 = 0.f, a1 = 0.f;
    for (int kk = 0; kk < d; kk += 32) {
      float4 u0[8], u1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { u0[u] = __ldg(reinterpret_cast<const float4*>(c0 + kk) + u); u1[u] = __ldg(reinterpret_cast<const float4*>(c1 + kk) + u); }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float q0 = qs[kk + 4 * u], q1 = qs[kk + 4 * u + 1], q2 = qs[kk + 4 * u + 2], q3 = qs[kk + 4 * u + 3];
        a0 = fmaf(q0, u0[u].x, a0); a1 = fmaf(q0, u1[u].x, a1);
        a0 = fmaf(q1, u0[u].y, a0); a1 = fmaf(q1, u1[u].y, a1);
        a0 = fmaf(q2, u0[u].z, a0); a1 = fmaf(q2, u1[u].z, a1);
        a0 = fmaf(q3, u0[u].w, a0); a1 = fmaf(q3, u1[u].w, a1);
      }
    }
    s0 = a0 + 0.0f; s1 = a1 + 0.0f;
  } else {
    s0 = exact_score(qs, c0, d); s1 = exact_score(qs, c1, d);
  }
}

// _exclude (layers/factorized_top_k.py:83-115) on a query's kf best candidates, sorted in `srt` as
// (key(score) << 32 | ~local index): identifiers in `exclusions[row]` get score - 1e5, the k_out best ADJUSTED scores
// win (ties -> lower position), the ORIGINAL scores and indices are written.  One warp; akey = kf words of scratch.
__device__ __forceinline__ void exclude_rerank(const unsigned long long* srt, int kf, unsigned long long* akey, long long row,
                                               const FinParams& p, int lane) {
  for (int t = lane; t < kf; t += 32) {
    const unsigned long long e = srt[t];
    const float s = key2f((unsigned int)(e >> 32));
    const long long gi = (long long)(0xFFFFFFFFu - (unsigned int)e) + p.index_offset;
    const long long ident = p.identifiers ? __ldg(p.identifiers + gi) : gi;
    bool isin = false;
    for (int x = 0; x < p.n_excl; ++x) isin |= (__ldg(p.exclusions + row * p.n_excl + x) == ident);
    const float adj = (isin ? s - 1.0e5f : s) + 0.0f;   // scores - isin * 1e5 (:104-107)
    akey[t] = ((unsigned long long)f2key(adj) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)t);
  }
  __syncwarp();
  for (int t = lane; t < kf; t += 32) {
    const unsigned long long mine = akey[t];
    int rank = 0;
    for (int j = 0; j < kf; ++j) rank += (akey[j] > mine) ? 1 : 0;
    if (rank < p.k_out) {
      const unsigned long long e = srt[t];
      p.out_s[row * p.k_out + rank] = key2f((unsigned int)(e >> 32));
      p.out_i[row * p.k_out + rank] = (long long)(0xFFFFFFFFu - (unsigned int)e) + p.index_offset;
    }
  }
}

// (3a) SELECT, one warp per query, no block-wide barriers, 4 KB of shared memory per query so that ~50 queries are
// in flight per SM (the work is a chain of dependent L2 round trips: occupancy is what hides it):
//   pass 1  survivors of the octet records -> their screening keys in shared memory;  tau = k-th largest (bitwise
//           search on register-resident keys);  lim = tau - 2 eps
//   pass 2  the records again: indices of the survivors with key >= lim  -> band_idx[row, :], band_n[row]
// COUNT mode needs one pass: #{screen > pos + eps} is counted, the indices with |screen - pos| <= eps form the band.
__host__ __device__ inline size_t sel_warp_bytes(int cap_keys, int segs) {
  return (size_t)cap_keys * 4 + (size_t)cap_keys * 2 + (size_t)((segs + 1 + 3) & ~3) * 4;
}
constexpr int SEL_WARPS = 8;

// one 32-record window of a query's octet records: lane -> (first index, 8 scores); returns false past the end
__device__ __forceinline__ bool load_record(const FinParams& p, long long row, const int* soff, int rec, int total_rec,
                                            unsigned int& ix0, float (&sc)[8]) {
  if (rec >= total_rec) return false;
  int lo = 0, hi = p.segs;  // largest segment with soff[seg] <= rec
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (soff[mid] <= rec) lo = mid; else hi = mid; }
  const long long at = (row * p.segs + lo) * p.cap_part + (rec - soff[lo]);
  ix0 = __ldg(p.cand_i + at);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.cand_s + at * 8));
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.cand_s + at * 8) + 1);
  sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
  return true;
}

template <int MODE>
__global__ void __launch_bounds__(SEL_WARPS * 32)
tc_select_kernel(const FinParams p) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * SEL_WARPS + warp;
  if (row >= p.Q) return;
  unsigned char* base = fsm + (size_t)warp * sel_warp_bytes(p.cap_keys, p.segs);
  unsigned int* keys = reinterpret_cast<unsigned int*>(base);                 // [cap_keys] screening keys
  unsigned short* loc = reinterpret_cast<unsigned short*>(keys + p.cap_keys);  // [cap_keys] (flat record index << 3 | column): total records < 8192
  int* soff = reinterpret_cast<int*>(loc + p.cap_keys);                        // [segs + 1]
  const unsigned int n32 = (unsigned int)p.N;   // N < 2^31
  unsigned int* band = p.band_idx + row * p.cap_band;

  // segment counts -> exclusive prefix
  int carry = 0; bool bad = false;
  for (int s0 = 0; s0 < p.segs; s0 += 32) {
    const int s = s0 + lane;
    int c = 0;
    if (s < p.segs) {
      const unsigned int cc = __ldg(p.count + row * p.segs + s);
      if (cc > (unsigned int)p.cap_part) bad = true; else c = (int)cc;
    }
    const int inc = warp_incl_scan(c, lane);
    if (s < p.segs) soff[s + 1] = carry + inc;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) soff[0] = 0;
  if (__any_sync(0xffffffffu, bad)) {  // a segment overflowed in the filter pass: records are missing -> exact fallback
    if (lane == 0) { p.overflow[row] = 1; p.band_n[row] = 0; }
    return;
  }
  __syncwarp();
  const int total_rec = soff[p.segs];
  const float thr_row = p.thr[row];
  if (MODE != FIN_COUNT && total_rec >= 8192) {   // the 16-bit record locator holds 13 bits of record index
    if (lane == 0) { p.overflow[row] = 1; p.band_n[row] = 0; }
    return;
  }

  if (MODE == FIN_COUNT) {
    // metrics/factorized_top_k.py:181-192: in_top_k(target = the positive, k) <=> #{candidates scoring > positive} < k.
    // screen > pos + eps => exact > pos (counted as is); |screen - pos| <= eps => re-scored exactly (3b).  When the
    // positive lies below the listed range, at least K listed candidates are definite (L_q > pos + eps), so
    // min(k, count) is exact.
    const float eps = 0.5f * p.cut[row];
    const float pos_s = ldexpf(p.pos[row], p.hdr->st.exp + p.qexp[row]);
    const float hi = pos_s + eps, lo_b = pos_s - eps;
    int definite = 0, m = 0;
    for (int rb = 0; rb < total_rec; rb += 32) {
      unsigned int ix0 = 0, amb = 0; int cnt = 0; float sc[8];
      if (load_record(p, row, soff, rb + lane, total_rec, ix0, sc)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool real = ix0 + j < n32;
          definite += (real && sc[j] > hi) ? 1 : 0;
          const bool a = real && sc[j] > lo_b && !(sc[j] > hi);
          amb |= a ? (1u << j) : 0u; cnt += a ? 1 : 0;
        }
      }
      const int incl = warp_incl_scan(cnt, lane);
      const int tot = __shfl_sync(0xffffffffu, incl, 31);
      if (m + tot <= p.cap_band) {
        int at_pos = m + incl - cnt;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (amb & (1u << j)) { band[at_pos] = ix0 + j; ++at_pos; }
      }
      m += tot;
    }
    definite = __reduce_add_sync(0xffffffffu, definite);
    if (lane == 0) {
      if (definite >= p.k) { p.out_count[row] = p.k; p.band_n[row] = 0; p.overflow[row] = 0; }
      else if (m > p.cap_band) { p.overflow[row] = 1; p.band_n[row] = 0; }
      else { p.out_count[row] = definite; p.band_n[row] = m; p.overflow[row] = 0; }   // 3b adds the re-scored ones
    }
    return;
  }

  // ---- TOPK / EXCLUDE, pass 1: screening keys of the survivors (score >= filter threshold, real row)
  int n = 0;
  unsigned int kmax = 0u, kmin = 0xFFFFFFFFu;
  for (int rb = 0; rb < total_rec; rb += 32) {
    float sc[8]; unsigned int ix0 = 0, keep = 0; int cnt = 0;
    if (load_record(p, row, soff, rb + lane, total_rec, ix0, sc)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool kp = sc[j] >= thr_row && ix0 + j < n32;
        keep |= kp ? (1u << j) : 0u; cnt += kp ? 1 : 0;
      }
    }
    const int incl = warp_incl_scan(cnt, lane);
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    if (n + tot <= p.cap_keys) {   // warp-uniform: the stores below need no per-entry bound check
      int at_pos = n + incl - cnt;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (keep & (1u << j)) {
          const unsigned int key = f2key(sc[j] + 0.0f);   // -0 -> +0: key order == float order
          kmax = max(kmax, key); kmin = min(kmin, key);
          keys[at_pos] = key;
          loc[at_pos] = (unsigned short)(((rb + lane) << 3) | j);
          ++at_pos;
        }
      }
    }
    n += tot;
  }
  if (n > p.cap_keys || (n < p.k && !p.allow_short)) { if (lane == 0) { p.overflow[row] = 1; p.band_n[row] = 0; } return; }
  __syncwarp();
  if (n < p.k) {
    // Sharded scan, threshold agreed across the shards (see comm.cu): this shard holds fewer than k candidates above it.
    // Every one of them is re-scored and emitted -- members of the global top-k are survivors on their shard by construction.
    const bool fits = n <= p.cap_band;
    for (int t = lane; t < n && fits; t += 32) {
      const unsigned int l = loc[t];
      const int rec = (int)(l >> 3);
      int lo = 0, hi = p.segs;
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (soff[mid] <= rec) lo = mid; else hi = mid; }
      band[t] = __ldg(p.cand_i + (row * p.segs + lo) * p.cap_part + (rec - soff[lo])) + (l & 7u);
    }
    if (lane == 0) { p.overflow[row] = fits ? 0u : 1u; p.band_n[row] = fits ? n : 0; }
    return;
  }
  // tau = k-th best screening score: bitwise search below the bits the largest and the smallest key share
  kmax = __reduce_max_sync(0xffffffffu, kmax); kmin = __reduce_min_sync(0xffffffffu, kmin);
  unsigned int tau_key;
  {
    const unsigned int diff = kmax ^ kmin;
    const int top = diff ? (31 - __clz(diff)) : -1;           // highest bit in which the survivors differ
    const unsigned int prefix0 = top >= 31 ? 0u : (top < 0 ? kmax : (kmax & ~((2u << top) - 1u)));   // top == -1: all keys equal
    if (n <= 512) {
      unsigned int kr[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { const int t = j * 32 + lane; kr[j] = t < n ? keys[t] : 0u; }
      tau_key = warp_kth_largest_regs<16>(kr, p.k, prefix0, top);
    } else if (n <= 1024) {
      unsigned int kr[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) { const int t = j * 32 + lane; kr[j] = t < n ? keys[t] : 0u; }
      tau_key = warp_kth_largest_regs<32>(kr, p.k, prefix0, top);
    } else {
      tau_key = warp_kth_largest_smem(keys, n, p.k, lane, prefix0, top);
    }
  }
  const float lim = key2f(tau_key) - p.cut[row];
  // Self-check that makes the threshold choice a pure performance matter: the whole band [lim, inf) must
  // lie above the filter threshold, otherwise survivors could be missing -> exact fallback.
  // (with a threshold agreed across shards the local band may reach below it: what is missing there cannot be in the
  //  GLOBAL top-k, and the local list is only an input of the cross-shard merge)
  if (!p.allow_short && (!(lim >= thr_row) || !(lim > -INFINITY))) { if (lane == 0) { p.overflow[row] = 1; p.band_n[row] = 0; } return; }
  // pass 2: only the band members (key >= key(lim)) go back to their record for the corpus index
  const unsigned int lim_key = f2key(lim + 0.0f);
  const unsigned int lt_mask = (1u << lane) - 1u;
  int m = 0;
  for (int tb = 0; tb < n; tb += 32) {
    const int t = tb + lane;
    const bool kp = t < n && keys[t] >= lim_key;
    const unsigned int vote = __ballot_sync(0xffffffffu, kp);
    if (kp) {
      const int at_pos = m + __popc(vote & lt_mask);
      if (at_pos < p.cap_band) {
        const unsigned int l = loc[t];
        const int rec = (int)(l >> 3);
        int lo = 0, hi = p.segs;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (soff[mid] <= rec) lo = mid; else hi = mid; }
        band[at_pos] = __ldg(p.cand_i + (row * p.segs + lo) * p.cap_part + (rec - soff[lo])) + (l & 7u);
      }
    }
    m += __popc(vote);
  }
  if (lane == 0) {
    if (m > p.cap_band) { p.overflow[row] = 1; p.band_n[row] = 0; }   // band too crowded (massive ties)
    else { p.overflow[row] = 0; p.band_n[row] = m; }
  }
}

// (3b) RE-SCORE + RANK, one 128-thread block per query: every band candidate's fp32 corpus row is fetched at once (one
// DRAM round trip per query; ~115 MB of random 256-byte rows per cfg2 batch, the HBM-bound part of the finalize step),
// the canonical fmaf chain gives the exact score, and each thread ranks its entry against the band in shared memory
// (ranks are unique: (score desc, index asc) is a total order).  EXCLUDE re-ranks the k best, COUNT adds #{exact > pos}.
constexpr int RS_BLOCK = 128;
template <int MODE>
__global__ void __launch_bounds__(RS_BLOCK)
tc_rescore_kernel(const FinParams p) {
  extern __shared__ __align__(16) unsigned char rsm[];
  unsigned long long* sk = reinterpret_cast<unsigned long long*>(rsm);                  // [cap_band] composite keys
  float* qs = reinterpret_cast<float*>(sk + p.cap_band);                                  // [d]
  unsigned long long* srt = reinterpret_cast<unsigned long long*>(qs + ((p.d + 3) & ~3)); // EXCLUDE: [2 * k]
  __shared__ int greater_sh;
  const long long row = blockIdx.x;
  const int tid = threadIdx.x;
  if (p.overflow[row] != 0) return;
  const int m = p.band_n[row];
  if (MODE == FIN_COUNT && m == 0) return;    // the select kernel already wrote the count
  for (int t = tid; t < p.d; t += RS_BLOCK) qs[t] = p.q[row * p.d + t];
  if (tid == 0) greater_sh = 0;
  __syncthreads();
  const unsigned int* band = p.band_idx + row * p.cap_band;
  if (MODE == FIN_COUNT) {
    const float pos = p.pos[row];
    int g = 0;
    for (int t = tid; t < m; t += RS_BLOCK) g += (exact_score(qs, p.corpus + (long long)band[t] * p.d, p.d) > pos) ? 1 : 0;
    g = __reduce_add_sync(0xffffffffu, g);
    if ((tid & 31) == 0 && g) atomicAdd(&greater_sh, g);
    __syncthreads();
    if (tid == 0) { const int c = p.out_count[row] + greater_sh; p.out_count[row] = c < p.k ? c : p.k; }
    return;
  }
  if (p.d == 64) {
    // d = 64 (the headline shape): EIGHT lanes fetch one 256-byte corpus row (32 bytes each, one coalesced request per row;
    // every row of the band is in flight before the first FMA), then the canonical chain k = 0..63 walks through the eight
    // lanes with the accumulator handed on by shuffle -- same arithmetic, same order, same bits as exact_score().
    constexpr int RPB = RS_BLOCK / 8;          // rows per block round
    constexpr int MAXR = 8;                    // rounds held in registers (m <= 128); larger bands loop
    const int sub = tid & 7, grp = tid >> 3;
    for (int t0 = 0; t0 < m; t0 += RPB * MAXR) {
      float4 a[MAXR], b[MAXR];
#pragma unroll
      for (int r = 0; r < MAXR; ++r) {
        const int t = t0 + r * RPB + grp;
        if (t < m) {
          const float4* src = reinterpret_cast<const float4*>(p.corpus + (long long)band[t] * 64) + sub * 2;
          a[r] = __ldg(src); b[r] = __ldg(src + 1);
        }
      }
      const float q0 = qs[sub * 8], q1 = qs[sub * 8 + 1], q2 = qs[sub * 8 + 2], q3 = qs[sub * 8 + 3];
      const float q4 = qs[sub * 8 + 4], q5 = qs[sub * 8 + 5], q6 = qs[sub * 8 + 6], q7 = qs[sub * 8 + 7];
#pragma unroll
      for (int r = 0; r < MAXR; ++r) {
        const int t = t0 + r * RPB + grp;
        const bool live = t < m;                    // uniform per 8-lane group; the shuffles below are executed by all lanes
        float acc = 0.f;
#pragma unroll
        for (int step = 0; step < 8; ++step) {
          if (sub == step && live) {
            acc = fmaf(q0, a[r].x, acc); acc = fmaf(q1, a[r].y, acc); acc = fmaf(q2, a[r].z, acc); acc = fmaf(q3, a[r].w, acc);
            acc = fmaf(q4, b[r].x, acc); acc = fmaf(q5, b[r].y, acc); acc = fmaf(q6, b[r].z, acc); acc = fmaf(q7, b[r].w, acc);
          }
          acc = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 24) | step);   // hand the accumulator to the next lane of the group
        }
        if (live && sub == 0) {
          const unsigned int idx = band[t];
          sk[t] = ((unsigned long long)f2key(acc + 0.0f) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
        }
      }
    }
  } else {
    for (int t = tid; t < m; t += RS_BLOCK) {
      const unsigned int idx = band[t];
      const float s = exact_score(qs, p.corpus + (long long)idx * p.d, p.d);
      sk[t] = ((unsigned long long)f2key(s) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
    }
  }
  __syncthreads();
  if (MODE == FIN_TOPK && m < p.k)   // short shard list: pad with (-inf, INT64_MAX), the merge's "no entry"
    for (int t = m + tid; t < p.k; t += RS_BLOCK) { p.out_s[row * p.k + t] = -INFINITY; p.out_i[row * p.k + t] = LLONG_MAX; }
  for (int t = tid; t < m; t += RS_BLOCK) {
    const unsigned long long mine = sk[t];
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < m; ++j) rank += (sk[j] > mine) ? 1 : 0;
    if (rank < p.k) {
      if (MODE == FIN_TOPK) {
        p.out_s[row * p.k + rank] = key2f((unsigned int)(mine >> 32));
        p.out_i[row * p.k + rank] = (long long)(0xFFFFFFFFu - (unsigned int)mine) + p.index_offset;
      } else {
        srt[rank] = mine;
      }
    }
  }
  if (MODE == FIN_EXCLUDE) {
    __syncthreads();
    if (tid < 32) exclude_rerank(srt, p.k, srt + p.k, row, p, tid);
  }
}

// EXCLUDE for the rows the exact fallback produced: tmp_[s,i] [Q, k] sorted lists -> the same re-ranking
__global__ void __launch_bounds__(FW_WARPS * 32)
tc_exclude_fallback_kernel(const FinParams p, const float* __restrict__ tmp_s, const long long* __restrict__ tmp_i,
                           const unsigned int* __restrict__ was_fallback) {
  extern __shared__ __align__(16) unsigned char fsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * FW_WARPS + warp;
  if (row >= p.Q || was_fallback[row] == 0) return;
  unsigned long long* srt = reinterpret_cast<unsigned long long*>(fsm) + (size_t)warp * 2 * p.k;
  unsigned long long* akey = srt + p.k;
  for (int t = lane; t < p.k; t += 32)
    srt[t] = ((unsigned long long)f2key(tmp_s[row * p.k + t] + 0.0f) << 32) |
             (unsigned long long)(0xFFFFFFFFu - (unsigned int)(tmp_i[row * p.k + t] - p.index_offset));
  __syncwarp();
  exclude_rerank(srt, p.k, akey, row, p, lane);
}

struct FallbackProvider {
  const float* q; const float* corpus; long long N; int d; long long index_offset; const unsigned int* overflow;
  float* qs;
  __device__ void begin(int row, void* extra) {
    qs = reinterpret_cast<float*>(extra);
    if (overflow[row] != 0) for (int t = threadIdx.x; t < d; t += blockDim.x) qs[t] = q[(long long)row * d + t];
  }
  __device__ long long count(int row) const { return overflow[row] != 0 ? N : 0; }
  __device__ void get(int, long long t, float& s, long long& i) const {
    const float* c = corpus + t * d;
    float acc = 0.f;
    for (int kk = 0; kk < d; ++kk) acc = fmaf(qs[kk], __ldg(c + kk), acc);
    s = acc + 0.0f; i = index_offset + t;
  }
};

// COUNT for the rows that overflowed: exact #{candidates scoring above the positive}, clipped at k
__global__ void __launch_bounds__(256)
tc_count_fallback_kernel(const float* __restrict__ q, const float* __restrict__ corpus, long long N, int d, int k,
                         const float* __restrict__ pos, const unsigned int* __restrict__ overflow, int* __restrict__ out_count) {
  __shared__ float qs[128];
  __shared__ int total;
  const int row = blockIdx.x;
  if (overflow[row] == 0) return;
  for (int t = threadIdx.x; t < d; t += 256) qs[t] = q[(long long)row * d + t];
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const float ps = pos[row];
  int c = 0;
  for (long long t = threadIdx.x; t < N; t += 256) {
    const float* cr = corpus + t * d;
    float acc = 0.f;
    for (int kk = 0; kk < d; ++kk) acc = fmaf(qs[kk], __ldg(cr + kk), acc);
    c += (acc + 0.0f > ps) ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&total, c);
  __syncthreads();
  if (threadIdx.x == 0) out_count[row] = total < k ? total : k;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Optional per-stage device timing (bench.py's roofline leg): CUDA events recorded on the launch stream
// around the stages of the call.  Off by default; adds two event records per stage when on.
struct Prof {
  bool on = false;
  static constexpr int MAXC = 512, STAGES_ = 4;  // 0 prep, 1 sample(+threshold), 2 filter, 3 finalize(+fallback)
  cudaEvent_t ev[MAXC][STAGES_ + 1];
  int created = 0, calls = 0;
};
static Prof g_prof;
static void prof_mark(cudaStream_t st, int stage) {
  if (!g_prof.on || g_prof.calls >= Prof::MAXC) return;
  int c = g_prof.calls;
  while (g_prof.created <= c) {
    for (int s = 0; s <= Prof::STAGES_; ++s) cudaEventCreate(&g_prof.ev[g_prof.created][s]);
    ++g_prof.created;
  }
  cudaEventRecord(g_prof.ev[c][stage], st);
  if (stage == Prof::STAGES_) ++g_prof.calls;
}

struct Plan {
  int kb, stages; long long n_tiles; int nqb; long long Qp;
  int stride, n_sample, group, bins_per_part, n_bins, bins_ld, parts_sample, parts_full, cap_part, cap_keys, cap_band;
  size_t smem;
  // workspace offsets
  size_t o_qimg, o_margin, o_cut, o_thr, o_qexp, o_count, o_ovf, o_binmax, o_cand, o_tmp, o_band, o_bandn, total;
};

static bool make_plan(long long Q, long long N, int d, int k, Plan& pl) {
  if (d <= 0 || d > 128 || Q <= 0 || N <= 0 || k <= 0) return false;   // d > 128: smem budget (A blocks + ring) not laid out
  pl.kb = (d + KSLAB - 1) / KSLAB;
  pl.stages = pl.kb == 1 ? 6 : 4;
  pl.n_tiles = ceil_div(N, TILE_N);
  pl.nqb = (int)ceil_div(Q, QBLK);
  pl.Qp = (long long)pl.nqb * QBLK;
  if (k > 256 || N >= (1ll << 31)) return false;   // finalize capacities are sized for ~4k survivors + band entries per query
  // sample every stride-th tile; keep at least 4k raw (tile, half) bins so the k-th largest bin max is a tight bound
  const long long full_tiles = N / TILE_N;  // the zero-padded last tile is never sampled (its 0 scores are not candidates)
  if (full_tiles < 1) return false;
  pl.stride = MAX_SAMPLE_STRIDE;
#ifdef TFRS_DEBUG_SWITCHES
  { static int ov = getenv("TFRS_TC_SAMPLE_STRIDE") ? atoi(getenv("TFRS_TC_SAMPLE_STRIDE")) : 0; if (ov >= 1 && ov <= 16) pl.stride = ov; }
#endif
  while (pl.stride > 1 && 2 * ceil_div(full_tiles, pl.stride) < 4ll * k) pl.stride >>= 1;
  pl.n_sample = (int)ceil_div(full_tiles, pl.stride);
  if (2ll * pl.n_sample < 4ll * k) return false;  // too few bins for a useful threshold -> caller uses the exact path
  const int sms = sm_count();
  int parts = sms / pl.nqb; if (parts < 1) parts = 1; if (parts > FIN_MAX_PARTS / 2) parts = FIN_MAX_PARTS / 2;
  pl.parts_sample = parts < pl.n_sample ? parts : pl.n_sample;
  pl.parts_full = (long long)parts < pl.n_tiles ? parts : (int)pl.n_tiles;
  {
    // one bin = `group` consecutive sampled tiles x 64 columns of an epilogue thread; <= max(512, 4k) <= MAX_BINS per query
    const int iters_max = (int)ceil_div(pl.n_sample, pl.parts_sample);
    const int target = 4 * k > 512 ? 4 * k : 512;
    int g = 1;
    while ((long long)pl.parts_sample * ceil_div(iters_max, g) * 2 > target && g < iters_max) ++g;
    pl.group = g;
    pl.bins_per_part = (int)ceil_div(iters_max, g);
    pl.n_bins = pl.parts_sample * pl.bins_per_part * 2;
    if (pl.n_bins > MAX_BINS || pl.n_bins < 2 * k) return false;
    pl.bins_ld = (pl.n_bins + 31) / 32 * 32;
  }
  {
    // octet records per (part, column-half) segment: expected lambda = k * stride / segments (the threshold sits near rank
    // 1.2 k / sampled fraction, ~0.8 records per survivor); capacity = 2 lambda + 12 sqrt(lambda) + 8, a power of two in 32..512
    const double lambda = (double)k * pl.stride / (pl.parts_full * 2.0);
    const double want = 2.0 * lambda + 12.0 * sqrt(lambda) + 8.0;
    int p2 = 32; while (p2 < want && p2 < 512) p2 <<= 1;
    pl.cap_part = p2;
    // finalize capacities per query: ~1.3 k * stride survivors are expected (+60 %), the re-scored band holds ~k + the
    // candidates within 2 eps of tau
    int ck = 1024; while (ck < 1.6 * 1.3 * k * pl.stride && ck < 4096) ck <<= 1;
    pl.cap_keys = ck;                 // the re-scored band holds cap_keys / 2 >= 512 entries (~k + the candidates within 2 eps of tau)
    pl.cap_band = ck / 2;
  }
  pl.smem = (size_t)(2 + pl.stages) * pl.kb * SLAB_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 1024); return r; };
  pl.o_qimg = take((size_t)pl.nqb * 2 * pl.kb * SLAB_BYTES);
  pl.o_margin = take((size_t)pl.Qp * 4);
  pl.o_cut = take((size_t)pl.Qp * 4);
  pl.o_thr = take((size_t)pl.Qp * 4);
  pl.o_qexp = take((size_t)pl.Qp * 4);
  pl.o_count = take((size_t)pl.Qp * pl.parts_full * 2 * 4);
  pl.o_ovf = take((size_t)pl.Qp * 4);
  pl.o_binmax = take((size_t)pl.Qp * pl.bins_ld * 4);
  pl.o_cand = take((size_t)pl.Qp * pl.parts_full * 2 * pl.cap_part * (8 * 4 + 4));
  pl.o_tmp = take((size_t)Q * k * 12);   // EXCLUDE: the exact fallback's [Q, k] lists before the re-ranking
  pl.o_band = take((size_t)pl.Qp * pl.cap_band * 4);
  pl.o_bandn = take((size_t)pl.Qp * 4);
  pl.total = o;
  return true;
}

template <int KB, int STAGES>
static int launch_scans(const Plan& pl, ScanParams sp, cudaStream_t st, int mode) {
  auto ks = tc_scan_kernel<KB, STAGES, MODE_SAMPLE>;
  auto kf = tc_scan_kernel<KB, STAGES, MODE_FILTER>;
  TFRS_DYN_SMEM(ks, (int)pl.smem);
  TFRS_DYN_SMEM(kf, (int)pl.smem);
  if (mode == MODE_SAMPLE) {
    sp.parts = pl.parts_sample; sp.n_seq = pl.n_sample; sp.stride = pl.stride;
    ks<<<(unsigned)(pl.nqb * sp.parts), THREADS, pl.smem, st>>>(sp);
  } else {
    sp.parts = pl.parts_full; sp.n_seq = (int)pl.n_tiles; sp.stride = 1;
    kf<<<(unsigned)(pl.nqb * sp.parts), THREADS, pl.smem, st>>>(sp);
  }
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

static int launch_scan_mode(const Plan& pl, const ScanParams& sp, cudaStream_t st, int mode) {
  if (pl.kb == 1) return launch_scans<1, 6>(pl, sp, st, mode);
  return launch_scans<2, 4>(pl, sp, st, mode);
}

template <int MODE>
static int launch_finalize(FinParams fp, cudaStream_t st) {
  auto ksel = tc_select_kernel<MODE>;
  auto krs = tc_rescore_kernel<MODE>;
  TFRS_DYN_SMEM(ksel, (int)(SEL_WARPS * sel_warp_bytes(4096, FIN_MAX_PARTS)));
  ksel<<<(unsigned)ceil_div(fp.Q, SEL_WARPS), SEL_WARPS * 32, SEL_WARPS * sel_warp_bytes(fp.cap_keys, fp.segs), st>>>(fp);
  TFRS_LAUNCH_CHECK();
  const size_t rs_smem = (size_t)fp.cap_band * 8 + (size_t)((fp.d + 3) & ~3) * 4 + (MODE == FIN_EXCLUDE ? (size_t)fp.k * 16 : 0);
  TFRS_DYN_SMEM(krs, 64 * 1024);
  krs<<<(unsigned)fp.Q, RS_BLOCK, rs_smem, st>>>(fp);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

// One call = qprep -> sampled pass -> threshold -> filter pass -> finalize (+ retry, + exact fallback).
struct Call {
  int mode;  // FIN_*
  const float* q; long long Q; const float* corpus; const void* index_buf; long long N; int d; int k; long long index_offset;
  void* ws; size_t ws_bytes; cudaStream_t st;
  float* out_s; long long* out_i;                                                       // TOPK / EXCLUDE
  const long long* identifiers; const long long* exclusions; int n_excl; int k_out;     // EXCLUDE
  const float* pos; int* out_count;                                                     // COUNT
  // sharded scan: called between the threshold kernel and the filter pass with the device arrays that define the filter
  // thresholds; may raise thr[] (comm.cu exchanges a global lower bound of the k-th best score across the shards)
  ThrHook hook; void* hook_ctx; int allow_short;
};

static int run_call(const Call& c) {
  Plan pl;
  if (!make_plan(c.Q, c.N, c.d, c.k, pl)) {
    set_error("topk_tc: shape (Q=%lld N=%lld d=%d k=%d) is outside the tensor-core path; use tfrs_topk_scan_f32",
              c.Q, c.N, c.d, c.k);
    return TFRS_ERR_UNSUPPORTED;
  }
  if (!c.ws || c.ws_bytes < pl.total + 16) { set_error("topk_tc: workspace too small (%zu < %zu)", c.ws_bytes, pl.total + 16); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = c.st;
  unsigned char* w = (unsigned char*)(((uintptr_t)c.ws + 15) & ~(uintptr_t)15);
  unsigned char* qimg = w + pl.o_qimg;
  float* margin = (float*)(w + pl.o_margin);
  float* cut = (float*)(w + pl.o_cut);
  float* thr = (float*)(w + pl.o_thr);
  int* qexp = (int*)(w + pl.o_qexp);
  unsigned int* count = (unsigned int*)(w + pl.o_count);
  unsigned int* ovf = (unsigned int*)(w + pl.o_ovf);
  float* binmax = (float*)(w + pl.o_binmax);
  float* cand_s = (float*)(w + pl.o_cand);
  unsigned int* cand_i = (unsigned int*)(w + pl.o_cand + (size_t)pl.Qp * pl.parts_full * 2 * pl.cap_part * 32);
  float* tmp_s = (float*)(w + pl.o_tmp);
  long long* tmp_i = (long long*)(w + pl.o_tmp + align_up((size_t)c.Q * c.k * 4, 8));
  const IndexHeader* hdr = (const IndexHeader*)c.index_buf;
  const unsigned char* cimg = (const unsigned char*)c.index_buf + HEADER_BYTES;

  prof_mark(st, 0);
  // (0) per-row exponent, image and margins: one launch
  tc_qprep_kernel<<<(unsigned)ceil_div(pl.Qp * 32, 256), 256, 0, st>>>(c.q, c.Q, pl.Qp, c.d, pl.kb, hdr, qimg, margin, cut, qexp);
  TFRS_LAUNCH_CHECK();
  ScanParams sp{};
  sp.qimg = qimg; sp.cimg = cimg; sp.Q = c.Q; sp.N = c.N; sp.nqb = pl.nqb; sp.n_tiles = pl.n_tiles;
  sp.binmax = binmax; sp.bins_ld = pl.bins_ld; sp.group = pl.group; sp.bins_per_part = pl.bins_per_part;
  sp.thr = thr; sp.count = count; sp.cand_s = cand_s; sp.cand_i = cand_i; sp.cap_part = pl.cap_part;
  sp.idesc = IDESC_F16_M128_N128;
  prof_mark(st, 1);
  // (1) sampled pass -> bin maxima -> k-th largest -> threshold
  int rc = launch_scan_mode(pl, sp, st, MODE_SAMPLE);
  if (rc) return rc;
  if (pl.n_bins <= 512)
    tc_threshold_kernel<16><<<(unsigned)ceil_div(c.Q * 32, 256), 256, 0, st>>>(binmax, pl.bins_ld, pl.n_bins, c.k, margin, thr, ovf, c.Q);
  else
    tc_threshold_kernel<32><<<(unsigned)ceil_div(c.Q * 32, 256), 256, 0, st>>>(binmax, pl.bins_ld, pl.n_bins, c.k, margin, thr, ovf, c.Q);
  TFRS_LAUNCH_CHECK();
  if (c.hook) {
    rc = c.hook(c.hook_ctx, thr, margin, cut, qexp, &hdr->st.exp, c.Q, st);
    if (rc) return rc;
  }
  prof_mark(st, 2);
  // (2) full pass with the fused threshold filter
  rc = launch_scan_mode(pl, sp, st, MODE_FILTER);
  if (rc) return rc;
  prof_mark(st, 3);
  // (3) exact re-scoring + final order (a warp per query); (4) exact fallback for the rows that asked for it
  FinParams fp{};
  fp.q = c.q; fp.corpus = c.corpus; fp.d = c.d; fp.k = c.k; fp.index_offset = c.index_offset; fp.N = c.N; fp.Q = c.Q;
  fp.count = count; fp.cand_s = cand_s; fp.cand_i = cand_i; fp.segs = pl.parts_full * 2; fp.cap_part = pl.cap_part;
  fp.cut = cut; fp.thr = thr; fp.overflow = ovf; fp.out_s = c.out_s; fp.out_i = c.out_i;
  fp.cap_keys = pl.cap_keys; fp.cap_band = pl.cap_band; fp.allow_short = c.allow_short;
  fp.band_idx = (unsigned int*)(w + pl.o_band); fp.band_n = (int*)(w + pl.o_bandn);
  fp.identifiers = c.identifiers; fp.exclusions = c.exclusions; fp.n_excl = c.n_excl; fp.k_out = c.k_out;
  fp.pos = c.pos; fp.qexp = qexp; fp.hdr = hdr; fp.out_count = c.out_count;
  if (c.mode == FIN_TOPK) rc = launch_finalize<FIN_TOPK>(fp, st);
  else if (c.mode == FIN_EXCLUDE) rc = launch_finalize<FIN_EXCLUDE>(fp, st);
  else rc = launch_finalize<FIN_COUNT>(fp, st);
  if (rc) return rc;
  if (c.mode == FIN_COUNT) {
    tc_count_fallback_kernel<<<(unsigned)c.Q, 256, 0, st>>>(c.q, c.corpus, c.N, c.d, c.k, c.pos, ovf, c.out_count);
    TFRS_LAUNCH_CHECK();
  } else {
    // CTAs of non-flagged queries exit immediately
    FallbackProvider prov{c.q, c.corpus, c.N, c.d, c.index_offset, ovf, nullptr};
    int cap = rowselect_cap(c.k);
    TFRS_DYN_SMEM(row_topk_kernel<FallbackProvider>, 64 * 1024);
    float* fs = c.mode == FIN_TOPK ? c.out_s : tmp_s;
    long long* fi = c.mode == FIN_TOPK ? c.out_i : tmp_i;
    row_topk_kernel<FallbackProvider><<<(unsigned)c.Q, RS_THREADS, rowselect_smem(cap, (size_t)c.d * 4), st>>>(prov, c.k, cap, fs, fi, c.k);
    TFRS_LAUNCH_CHECK();
    if (c.mode == FIN_EXCLUDE) {
      tc_exclude_fallback_kernel<<<(unsigned)ceil_div(c.Q, FW_WARPS), FW_WARPS * 32, (size_t)FW_WARPS * 2 * c.k * 8, st>>>(fp, tmp_s, tmp_i, ovf);
      TFRS_LAUNCH_CHECK();
    }
  }
  prof_mark(st, 4);
  return TFRS_OK;
}

}  // namespace tc
}  // namespace tfrs

using namespace tfrs;
using namespace tfrs::tc;

extern "C" size_t tfrs_index_bytes(int64_t N, int d) {
  if (N <= 0 || d <= 0 || d > 128) return 0;
  int kb = (d + KSLAB - 1) / KSLAB;
  return (size_t)HEADER_BYTES + (size_t)ceil_div(N, TILE_N) * kb * SLAB_BYTES;
}

extern "C" int tfrs_index_build(const float* corpus, int64_t N, int d, void* index_buf, size_t index_bytes, void* stream) {
  TFRS_CHECK_ARG(corpus && index_buf && N > 0 && d > 0, "index_build: bad arguments");
  if (d > 128) { set_error("index_build: d=%d > 128 is not supported by the tensor-core path", d); return TFRS_ERR_UNSUPPORTED; }
  TFRS_CHECK_ARG(index_bytes >= tfrs_index_bytes(N, d), "index_build: buffer too small");
  TFRS_CHECK_ARG((reinterpret_cast<uintptr_t>(index_buf) & 15) == 0, "index_build: buffer must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  IndexHeader h{};
  h.d = d; h.kb = (d + KSLAB - 1) / KSLAB; h.d_pad = h.kb * KSLAB; h.n = N; h.n_tiles = ceil_div(N, TILE_N);
  TFRS_CUDA(cudaMemsetAsync(index_buf, 0, HEADER_BYTES, st));
  header_kernel<<<1, 1, 0, st>>>(reinterpret_cast<IndexHeader*>(index_buf), h);
  TFRS_LAUNCH_CHECK();
  SideStats* cst = &reinterpret_cast<IndexHeader*>(index_buf)->st;
  const unsigned sgrid = (unsigned)(sm_count() * 8);
  corpus_stats_kernel<0><<<sgrid, 256, 0, st>>>(corpus, N, d, cst);
  TFRS_LAUNCH_CHECK();
  side_exp_kernel<<<1, 1, 0, st>>>(cst);
  TFRS_LAUNCH_CHECK();
  corpus_stats_kernel<1><<<sgrid, 256, 0, st>>>(corpus, N, d, cst);
  TFRS_LAUNCH_CHECK();
  long long chunks = h.n_tiles * TILE_N * (long long)h.kb * 8;
  unsigned blocks = (unsigned)(ceil_div(chunks, 256) < (1 << 20) ? ceil_div(chunks, 256) : (1 << 20));
  tile_image_kernel<<<blocks, 256, 0, st>>>(corpus, N, d, h.kb, h.n_tiles, cst, (unsigned char*)index_buf + HEADER_BYTES);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

extern "C" size_t tfrs_topk_tc_workspace_bytes(int64_t Q, int64_t N, int d, int k) {
  Plan pl;
  if (!make_plan(Q, N, d, k, pl)) return 0;
  return pl.total + 16;
}

extern "C" int tfrs_topk_tc_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d,
                                int k, int64_t index_offset, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes,
                                void* stream) {
  TFRS_CHECK_ARG(q && corpus && index_buf && out_scores && out_idx, "topk_tc: NULL pointer");
  TFRS_CHECK_ARG(k <= N, "input must have at least k columns. Had %lld, needed %d", (long long)N, k);
  Call c{};
  c.mode = FIN_TOPK; c.q = q; c.Q = Q; c.corpus = corpus; c.index_buf = index_buf; c.N = N; c.d = d; c.k = k;
  c.index_offset = index_offset; c.ws = ws; c.ws_bytes = ws_bytes; c.st = (cudaStream_t)stream;
  c.out_s = out_scores; c.out_i = (long long*)out_idx;
  return run_call(c);
}

// internal (comm.cu): the local scan of the sharded call, with the threshold hook and short lists allowed
int tfrs::tc_topk_sharded_local(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d, int k,
                                int64_t index_offset, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, void* stream,
                                tfrs::ThrHook hook, void* hook_ctx) {
  Call c{};
  c.mode = FIN_TOPK; c.q = q; c.Q = Q; c.corpus = corpus; c.index_buf = index_buf; c.N = N; c.d = d; c.k = k;
  c.index_offset = index_offset; c.ws = ws; c.ws_bytes = ws_bytes; c.st = (cudaStream_t)stream;
  c.out_s = out_scores; c.out_i = (long long*)out_idx;
  c.hook = hook; c.hook_ctx = hook_ctx; c.allow_short = hook ? 1 : 0;
  return run_call(c);
}

extern "C" int tfrs_topk_tc_exclude_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d,
                                        int k, int64_t index_offset, const int64_t* identifiers, const int64_t* exclusions,
                                        int n_excl, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, void* stream) {
  TFRS_CHECK_ARG(q && corpus && index_buf && out_scores && out_idx && exclusions && n_excl > 0, "topk_tc_exclude: bad argument");
  TFRS_CHECK_ARG((long long)k + n_excl <= N, "input must have at least k columns. Had %lld, needed %d", (long long)N, k + n_excl);
  Call c{};
  c.mode = FIN_EXCLUDE; c.q = q; c.Q = Q; c.corpus = corpus; c.index_buf = index_buf; c.N = N; c.d = d; c.k = k + n_excl;
  c.index_offset = index_offset; c.ws = ws; c.ws_bytes = ws_bytes; c.st = (cudaStream_t)stream;
  c.out_s = out_scores; c.out_i = (long long*)out_idx;
  c.identifiers = (const long long*)identifiers; c.exclusions = (const long long*)exclusions; c.n_excl = n_excl; c.k_out = k;
  return run_call(c);
}

extern "C" int tfrs_topk_tc_count_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d,
                                      int k, const float* positive_scores, int32_t* out_count, void* ws, size_t ws_bytes,
                                      void* stream) {
  TFRS_CHECK_ARG(q && corpus && index_buf && positive_scores && out_count, "topk_tc_count: NULL pointer");
  TFRS_CHECK_ARG(k <= N, "input must have at least k columns. Had %lld, needed %d", (long long)N, k);
  Call c{};
  c.mode = FIN_COUNT; c.q = q; c.Q = Q; c.corpus = corpus; c.index_buf = index_buf; c.N = N; c.d = d; c.k = k;
  c.ws = ws; c.ws_bytes = ws_bytes; c.st = (cudaStream_t)stream;
  c.pos = positive_scores; c.out_count = out_count;
  return run_call(c);
}

// Debug/test introspection: where the per-query survivor counts / fallback flags of the last call live
// inside the caller's workspace (byte offsets from the 16-byte-aligned workspace base).
extern "C" int tfrs_topk_tc_layout(int64_t Q, int64_t N, int d, int k, int64_t* out8) {
  TFRS_CHECK_ARG(out8, "topk_tc_layout: NULL pointer");
  Plan pl;
  if (!make_plan(Q, N, d, k, pl)) { set_error("topk_tc_layout: unsupported shape"); return TFRS_ERR_UNSUPPORTED; }
  out8[0] = (int64_t)pl.o_count; out8[1] = (int64_t)pl.o_ovf; out8[2] = (int64_t)pl.o_thr; out8[3] = (int64_t)pl.o_cand;
  out8[4] = pl.parts_full * 2; out8[5] = pl.cap_part; out8[6] = pl.Qp; out8[7] = (int64_t)pl.o_cut;
  return TFRS_OK;
}

extern "C" int tfrs_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.calls = 0;
  return TFRS_OK;
}

// Synchronises the device, then returns the summed stage times (ms) over the calls recorded since
// tfrs_profile_enable(1): stage_ms[0..3] = prep, sample+threshold, filter, finalize+fallback.
extern "C" int tfrs_profile_read(float* stage_ms, int* calls) {
  TFRS_CHECK_ARG(stage_ms && calls, "profile_read: NULL pointer");
  TFRS_CUDA(cudaDeviceSynchronize());
  for (int s = 0; s < 4; ++s) stage_ms[s] = 0.f;
  for (int c = 0; c < g_prof.calls; ++c)
    for (int s = 0; s < 4; ++s) {
      float ms = 0.f;
      TFRS_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[c][s], g_prof.ev[c][s + 1]));
      stage_ms[s] += ms;
    }
  *calls = g_prof.calls;
  return TFRS_OK;
}
