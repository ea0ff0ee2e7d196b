// topk_tc.cu -- K2: brute-force top-K on the 5th-gen tensor cores (tcgen05 + TMEM + bulk-TMA), sm_100a.
//
// Replaces  scores = matmul(q, c^T); top_k(scores, k)  (layers/factorized_top_k.py:603-605) for large
// corpora.  The [Q,N] score matrix never exists in HBM:
//
//   index time   tfrs_index_build : corpus fp32 -> fp16 image (exact 2^e rescale), pre-tiled as 128-row UMMA SWIZZLE_128B
//                K-major tiles (one contiguous 16 KB block per 64-wide K slab), + max row norm / max |element|.
//   query time   (0) q stats/image: queries -> the same fp16 tile image; per-query error margin
//                (1) tc_scan<SAMPLE> : screening GEMM over every 4th corpus tile; epilogue keeps only the
//                    per-(query, 64-column bin) max  -> K-th largest bin max = a valid lower bound L_q of
//                    the K-th best screening score (K distinct bins hold K distinct candidates >= it)
//                (2) tc_scan<FILTER> : screening GEMM over the whole corpus; epilogue compares the fp32
//                    accumulators (read from TMEM) with T_q = L_q - margin_q and appends the rare
//                    survivors (score, index) to a per-query list -- nothing else leaves the SM
//                (3) tc_finalize : per query: tau = K-th best screening score; survivors within the error
//                    band of tau are re-scored EXACTLY (sequential fp32 fmaf chain on the fp32 corpus),
//                    sorted by (score desc, index asc) -> bit-identical to the exact CUDA-core path.
//                (4) overflow fallback (list capacity exceeded; adversarial inputs only): exact scan.
//
// Why the result is exact: |screen(q,c) - exact(q,c)| <= eps_q = E_REL*|q|*max|c| (fp16 rounding of both
// operands: (2u+u^2) sum|q_k c_k| with u = 2^-11, plus accumulation slack; Cauchy-Schwarz).  Any member
// of the exact top-K has screening score >= tau - 2*eps_q >= L_q - 2*eps_q, so it is in the list and in
// the re-scored band.
//
// Kernel shape (per CTA, 1 CTA / SM, 384 threads): 256 queries (two 128-row A blocks, resident in smem)
// x a contiguous range of 128-row corpus tiles streamed through a 4-6 stage bulk-TMA ring; warp 0 = TMA
// producer, warp 1 = MMA issuer (one thread, tcgen05.mma M=128 N=128 K=16, fp16 -> fp32 in TMEM),
// warp 2 = TMEM allocator, warps 4-11 = epilogue (one query row per thread, tcgen05.ld 32x32b.x32).
// TMEM holds 2 A-blocks x 2 buffers x 128 columns = all 512 columns, so tile t+1's MMAs overlap tile
// t's epilogue.  Each B tile feeds two MMAs (both A blocks): 16 KB of L2->smem traffic per 512
// tensor-core cycles keeps the chip under the ~6.3 KB/clk L2 fabric limit.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
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
constexpr int CAND_CAP = 2048;       // survivors kept per query
constexpr int MAX_SAMPLE_STRIDE = 4;
constexpr int FIN_MAX_PARTS = 2 * 148;  // survivor-list segments per query: (corpus part, column half)
// Screening error model (operands: fp16 after an exact power-of-two rescale of each side so that the
// largest magnitude lands in [2^14, 2^15); accumulate: fp32 in TMEM):
//   |x^ - x| <= 2^-11 |x| (+2^-25 absolute below the fp16 normal range, negligible after the rescale)
//   => |screen - exact| <= (2^-10 + 2^-22) sum|q_k c_k|  + accumulation slack (budget 2^-14) + fmaf-chain 2^-16
//   <= E_REL * |q| * |c|   (Cauchy-Schwarz), E_REL = 0.00108 including the 0.1 % norm inflation.
constexpr float E_REL = 0.00108f;
constexpr float E_ACC = 0.00013f;    // run-to-run slack between the two passes (they are bit-identical in practice)

// Debug A/B switch (env TFRS_TC_BF16=1): screen with unscaled bf16 operands instead of scaled fp16.
static int use_bf16() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TFRS_TC_BF16"); v = (e && atoi(e)) ? 1 : 0; }
  return v;
}
constexpr float E_REL_BF16 = 0.0083f;

struct SideStats {              // per operand side (corpus at index time, queries per call)
  unsigned int max_norm2_bits;  // max_i |x_i|^2   (float bits; non-negative so uint order == float order)
  unsigned int amax_bits;       // max_ij |x_ij|
  int exp;                      // rescale exponent e: x * 2^e has its largest magnitude in [2^14, 2^15)
  int pad;
};
struct IndexHeader {
  SideStats st;
  int d, d_pad, kb, pad;
  long long n, n_tiles;
};

// ------------------------------------------------------------------------------------------------
// image builders: fp32 [rows, d] -> fp16 128-row tiles, each K slab of 64 as one swizzled 16 KB block
//   byte offset of element (r, k) inside a tile = (k/64)*16384 + r*128 + (((k%64)/8) ^ (r%8))*16 + (k%8)*2
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_image_kernel(const float* __restrict__ src, long long rows, int d, int kb, long long n_tiles,
                  const SideStats* __restrict__ st, int bf16, unsigned char* __restrict__ img) {
  const int scale_exp = bf16 ? 0 : st->exp;
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
      if (bf16) { __nv_bfloat16 b = __float2bfloat16_rn(f); v[j] = *reinterpret_cast<__half*>(&b); }
    }
    unsigned char* dst = img + tile * ((long long)kb * SLAB_BYTES) + (long long)slab * SLAB_BYTES + r * 128 + ((cj ^ (r & 7)) * 16);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
  }
}

// max row norm^2 and max |element| of a [rows, d] matrix (one warp per row -> coalesced)
__global__ void __launch_bounds__(256)
side_stats_kernel(const float* __restrict__ src, long long rows, int d, SideStats* __restrict__ st,
                  float* __restrict__ row_n2 /* nullable: per-row |x|^2 */) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * 256) >> 5;
  float best_n2 = 0.f, best_a = 0.f;
  for (long long row = warp; row < rows; row += nwarps) {
    const float* p = src + row * d;
    float n2 = 0.f, a = 0.f;
    for (int k = lane; k < d; k += 32) { float x = p[k]; n2 = fmaf(x, x, n2); a = fmaxf(a, fabsf(x)); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n2 += __shfl_xor_sync(0xffffffffu, n2, o); a = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, o)); }
    best_n2 = fmaxf(best_n2, n2); best_a = fmaxf(best_a, a);
    if (row_n2 && lane == 0) row_n2[row] = n2;
  }
  if (lane == 0) {
    if (best_n2 > 0.f) atomicMax(&st->max_norm2_bits, __float_as_uint(best_n2 * 1.0001f));  // slack for the tree order
    if (best_a > 0.f) atomicMax(&st->amax_bits, __float_as_uint(best_a));
  }
}
__global__ void side_exp_kernel(SideStats* st, int target) {
  const float amax = __uint_as_float(st->amax_bits);
  int x = 0;
  if (amax > 0.f && amax < INFINITY) (void)frexpf(amax, &x);  // amax = m * 2^x, m in [0.5, 1)
  st->exp = (amax > 0.f && amax < INFINITY) ? (target - x) : 0;  // amax * 2^exp in [2^(target-1), 2^target)
}
static int fp16_target() {
  static int v = -100;
  if (v == -100) { const char* e = getenv("TFRS_TC_FP16_TARGET"); v = e ? atoi(e) : 15; }
  return v;
}

__global__ void header_kernel(IndexHeader* dst, IndexHeader h) { *dst = h; }

// per-query margins from |q| and the corpus max norm, expressed in SCREENING units (scores scaled by
// 2^(exp_q + exp_c), an exact power of two)
__global__ void __launch_bounds__(256)
qmargin_kernel(long long Q, long long Qp, const IndexHeader* __restrict__ hdr, const SideStats* __restrict__ qst, int bf16,
               float* __restrict__ margin /* in: |q|^2 per row (rows < Q); out: filter margin */, float* __restrict__ cut) {
  long long row = (long long)blockIdx.x * 256 + threadIdx.x;
  if (row >= Qp) return;
  const float n2 = row < Q ? margin[row] * 1.0001f : 0.f;  // slack for the tree-order norm
  const float cn = sqrtf(__uint_as_float(hdr->st.max_norm2_bits)) * 1.001f;
  const float qn = sqrtf(n2) * 1.001f;
  const int se = bf16 ? 0 : (hdr->st.exp + qst->exp);
  const float e = ldexpf((bf16 ? E_REL_BF16 : E_REL) * qn * cn, se) + 1e-30f;
  margin[row] = 2.f * e + ldexpf(E_ACC * qn * cn, se);
  cut[row] = 2.f * e;
}

// ------------------------------------------------------------------------------------------------
// the screening GEMM
// ------------------------------------------------------------------------------------------------
enum { MODE_SAMPLE = 0, MODE_FILTER = 1, MODE_DBG_LDONLY = 2, MODE_DBG_NOLD = 3 };  // 2,3: sample-pass experiments

struct ScanParams {
  const unsigned char* qimg;    // query tile image  [2*nqb tiles][KB][16 KB]
  const unsigned char* cimg;    // corpus tile image [n_tiles][KB][16 KB]
  long long Q, N;
  int nqb, parts, n_seq, stride;  // tile sequence: tile(u) = u * stride, u in [0, n_seq)
  long long n_tiles;
  // SAMPLE
  float* binmax; int bins_ld;     // [Qp, bins_ld], 2 bins per sampled tile
  // FILTER
  const float* thr;               // [Qp]
  unsigned int* count;            // [Qp, parts]   survivors found by each (query, corpus part)
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
      for (int it = 0; it < n_iter; ++it) {
        const long long tile = (long long)(u_begin + it) * p.stride;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], KB * SLAB_BYTES);
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
    for (int it = 0; it < n_iter; ++it) {
      const int buf = it & 1;
      const uint32_t tphase = (it >> 1) & 1;
      const int u = u_begin + it;
      const long long tile = (long long)u * p.stride;
      const long long col0 = tile * TILE_N;  // zero-padded rows of the last tile score 0: dropped in finalize (idx >= N)
      mbar_wait(&t_full[buf], tphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ab * 2 + buf) * TILE_N);
      float binm = -INFINITY;
      {
        const int h = half;
        uint32_t r[64];
        if (MODE != MODE_DBG_NOLD) {
          tmem_ld64(taddr + h * 64, r);
          tmem_ld_wait64(r);
        } else {
#pragma unroll
          for (int j = 0; j < 64; ++j) r[j] = (uint32_t)(it * 131 + j * 7 + h + lane);
        }
        // this warp's TMEM reads of the accumulator buffer are complete: release it before the math
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[buf]);
        if (MODE == MODE_DBG_LDONLY) binm = fmaxf(binm, __uint_as_float(r[0] ^ r[63]));
        else {
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
      }
      if (MODE != MODE_FILTER && row_ok) p.binmax[row * p.bins_ld + 2 * u + half] = binm;
    }
    if (MODE == MODE_FILTER) p.count[((long long)row * p.parts + part) * 2 + half] = my_ovf ? (cap + 1u) : my_cnt;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------
// threshold from the bin maxima; finalize; fallback
// ------------------------------------------------------------------------------------------------
// orderable key: larger float <=> larger unsigned (NaN sorts above +inf; -0 < +0 is harmless here)
__device__ __forceinline__ unsigned int f2key(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Block-wide k-th largest of n keys by 4 passes of 8-bit radix histograms (256 threads).
// `key_at(i)` must be cheap and repeatable.  Returns the k-th largest key (1-based k <= n) to all threads.
template <class KeyAt>
__device__ unsigned int block_kth_largest(KeyAt key_at, int n, int k, unsigned int* hist /*[256] smem*/, unsigned int* bcast /*[2] smem*/) {
  unsigned int prefix = 0, mask = 0;
  int remaining = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int t = threadIdx.x; t < 256; t += blockDim.x) hist[t] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned int key = key_at(i);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns buckets [8l, 8l+8); walk from the top bucket down until `remaining` is covered
      const int lane = threadIdx.x;
      unsigned int loc[8], tot = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) { loc[b] = hist[lane * 8 + b]; tot += loc[b]; }
      // suffix sum over lanes above
      unsigned int above = 0;
#pragma unroll
      for (int l = 31; l >= 0; --l) {  // every lane executes every shuffle
        const unsigned int tl = __shfl_sync(0xffffffffu, tot, l);
        if (l > lane) above += tl;
      }
      unsigned int run = above;
      int found = -1; unsigned int before = 0;
#pragma unroll
      for (int b = 7; b >= 0; --b) {
        if (found < 0 && run + loc[b] >= (unsigned)remaining) { found = lane * 8 + b; before = run; }
        run += loc[b];
      }
      const bool mine = (found >= 0) && (above < (unsigned)remaining);
      if (mine) { bcast[0] = (unsigned int)found; bcast[1] = before; }
    }
    __syncthreads();
    const unsigned int bucket = bcast[0];
    remaining -= (int)bcast[1];
    prefix |= bucket << shift;
    mask |= 255u << shift;
    __syncthreads();
  }
  return prefix;
}

// k-th largest bin maximum of the sampled pass -> filter threshold T = L - margin.
// The row of bin maxima is staged once in shared memory (when it fits) and radix-selected there.
constexpr int THR_SMEM_BINS = 12288;
__global__ void __launch_bounds__(256)
tc_threshold_kernel(const float* __restrict__ binmax, int bins_ld, int n_bins, int k, const float* __restrict__ margin,
                    float* __restrict__ thr, unsigned int* __restrict__ overflow) {
  extern __shared__ unsigned int thr_keys[];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int bcast[2];
  const int row = blockIdx.x;
  const float* src = binmax + (long long)row * bins_ld;
  unsigned int kth;
  if (n_bins <= THR_SMEM_BINS) {
    for (int i = threadIdx.x; i < n_bins; i += 256) thr_keys[i] = f2key(__ldg(src + i));
    __syncthreads();
    kth = block_kth_largest([&](int i) { return thr_keys[i]; }, n_bins, k, hist, bcast);
  } else {
    kth = block_kth_largest([&](int i) { return f2key(__ldg(src + i)); }, n_bins, k, hist, bcast);
  }
  if (threadIdx.x == 0) { thr[row] = key2f(kth) - margin[row]; overflow[row] = 0; }
}

constexpr int FIN_MAXM = 1024;  // survivors re-scored exactly per query (band around tau)

__global__ void __launch_bounds__(256, 5)
tc_finalize_kernel(const float* __restrict__ q, const float* __restrict__ corpus, int d, int k, long long index_offset,
                   const unsigned int* __restrict__ count, const float* __restrict__ cand_s,
                   const unsigned int* __restrict__ cand_i, int parts, int cap_part,
                   const float* __restrict__ cut, const float* __restrict__ thr, long long N,
                   unsigned int* __restrict__ overflow, float* __restrict__ out_s, long long* __restrict__ out_i) {
  extern __shared__ __align__(16) unsigned char fsm[];
  long long* ei = reinterpret_cast<long long*>(fsm);                       // [FIN_MAXM]  exact stage indices
  float* es = reinterpret_cast<float*>(fsm + (size_t)FIN_MAXM * 8);         // [FIN_MAXM]  exact scores
  float* as = es + FIN_MAXM;                                                // [CAND_CAP]  screening scores
  unsigned int* ai = reinterpret_cast<unsigned int*>(as + CAND_CAP);        // [CAND_CAP]  local indices
  float* qs = reinterpret_cast<float*>(ai + CAND_CAP);                      // [d]
  __shared__ unsigned int hist[256];
  __shared__ unsigned int bcast[2];
  __shared__ int m_sh;
  const int row = blockIdx.x, tid = threadIdx.x;
  // gather this query's per-(part, half) segments: counts -> exclusive prefix (parts <= 296, one thread each)
  __shared__ int seg_off[FIN_MAX_PARTS + 1];
  __shared__ int seg_bad;
  if (tid == 0) seg_bad = 0;
  __syncthreads();
  for (int pt = tid; pt < parts; pt += 256) {
    unsigned int c = count[(long long)row * parts + pt];
    if (c > (unsigned)cap_part) { seg_bad = 1; c = 0; }
    seg_off[pt + 1] = (int)c;
  }
  if (tid == 0) seg_off[0] = 0;
  __syncthreads();
  if (tid == 0) { int a = 0; for (int pt = 1; pt <= parts; ++pt) { a += seg_off[pt]; seg_off[pt] = a; } }
  __syncthreads();
  if (seg_bad) {  // a segment overflowed: exact fallback
    if (tid == 0) overflow[row] = 1;
    return;
  }
  for (int t = tid; t < d; t += 256) qs[t] = q[(long long)row * d + t];
  // All octet records of all segments as one flat list: a thread takes a record (one segment lookup, one index
  // load, two 16-byte score loads), keeps the true survivors (>= filter threshold, real row) and the warp
  // reserves its output slots with a shuffle prefix sum + ONE shared atomic.
  __shared__ int n_sh;
  if (tid == 0) n_sh = 0;
  __syncthreads();
  const float thr_row = thr[row];
  const int total_rec = seg_off[parts];
  const int lane_id = tid & 31;
  for (int rb = 0; rb < total_rec; rb += 256) {  // block-uniform trip count (warp shuffles inside)
    const int rec = rb + tid;
    float sc[8]; unsigned int ix0 = 0; int cnt = 0; unsigned int keep = 0;
    if (rec < total_rec) {
      int lo = 0, hi = parts;  // largest pt with seg_off[pt] <= rec
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (seg_off[mid] <= rec) lo = mid; else hi = mid; }
      const long long at = ((long long)row * parts + lo) * cap_part + (rec - seg_off[lo]);
      ix0 = __ldg(cand_i + at);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(cand_s + at * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(cand_s + at * 8) + 1);
      sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (sc[j] >= thr_row && (unsigned long long)(ix0 + j) < (unsigned long long)N) { keep |= 1u << j; ++cnt; }
    }
    int incl = cnt;  // warp inclusive prefix sum of the per-lane survivor counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane_id >= o) incl += v; }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane_id == 0 && warp_total) base = atomicAdd(&n_sh, warp_total);
    base = __shfl_sync(0xffffffffu, base, 0);
    int pos = base + incl - cnt;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (keep & (1u << j)) {
        if (pos < CAND_CAP) { as[pos] = sc[j]; ai[pos] = ix0 + j; }
        ++pos;
      }
    }
  }
  __syncthreads();
  const int n = n_sh;
  if (n > CAND_CAP || n < k) {
    if (tid == 0) overflow[row] = 1;
    return;
  }
  if (tid == 0) m_sh = 0;
  __syncthreads();
  // tau = k-th best screening score; keep the survivors inside its error band
  const unsigned int tau_key = block_kth_largest([&](int i) { return f2key(as[i]); }, n, k, hist, bcast);
  const float lim = key2f(tau_key) - cut[row];
  // Self-check that makes the threshold choice a pure performance matter: the whole band [lim, inf) must
  // lie above the filter threshold, otherwise survivors could be missing -> exact fallback.
  if (!(lim >= thr[row]) || !(lim > -INFINITY)) {
    if (tid == 0) overflow[row] = 1;
    return;
  }
  for (int tb = 0; tb < n; tb += 256) {
    const int t = tb + tid;
    const bool keep = t < n && as[t] >= lim;
    const unsigned int vote = __ballot_sync(0xffffffffu, keep);
    if (vote) {
      int base = 0;
      if ((tid & 31) == 0) base = atomicAdd(&m_sh, __popc(vote));
      base = __shfl_sync(0xffffffffu, base, 0);
      const int pos = base + __popc(vote & ((1u << (tid & 31)) - 1u));
      if (keep && pos < FIN_MAXM) ei[pos] = (long long)ai[t];
    }
  }
  __syncthreads();
  const int m = m_sh;
  if (m > FIN_MAXM) {  // band too crowded (massive ties): exact fallback
    if (tid == 0) overflow[row] = 1;
    return;
  }
  // exact re-scoring: the canonical sequential fmaf chain on the fp32 corpus
  for (int t = tid; t < m; t += 256) {
    const float* c = corpus + ei[t] * d;
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
        float4 cv = __ldg(reinterpret_cast<const float4*>(c + kk));
        acc = fmaf(qs[kk], cv.x, acc); acc = fmaf(qs[kk + 1], cv.y, acc);
        acc = fmaf(qs[kk + 2], cv.z, acc); acc = fmaf(qs[kk + 3], cv.w, acc);
      }
    } else {
      for (int kk = 0; kk < d; ++kk) acc = fmaf(qs[kk], __ldg(c + kk), acc);
    }
    es[t] = acc;
  }
  __syncthreads();
  if (m <= 512) {
    // rank sort: the band is small, so each thread ranks its entry against all others (broadcast reads, no
    // barriers) and writes it straight to its output slot.  Ranks are unique: (score, index) is a total order.
    for (int t = tid; t < m; t += 256) {
      const float s_t = es[t]; const long long i_t = ei[t];
      int rank = 0;
      for (int j = 0; j < m; ++j) rank += better(es[j], ei[j], s_t, i_t) ? 1 : 0;
      if (rank < k) {
        out_s[(long long)row * k + rank] = s_t;
        out_i[(long long)row * k + rank] = i_t + index_offset;
      }
    }
    return;
  }
  int P2 = 2; while (P2 < m) P2 <<= 1;
  for (int t = m + tid; t < P2; t += 256) { es[t] = -INFINITY; ei[t] = LLONG_MAX; }
  __syncthreads();
  bitonic_sort_desc(es, ei, P2);  // (exact score desc, index asc)
  for (int t = tid; t < k; t += 256) {
    out_s[(long long)row * k + t] = es[t];
    out_i[(long long)row * k + t] = ei[t] + index_offset;
  }
}

struct FallbackProvider {
  const float* q; const float* corpus; long long N; int d; long long index_offset; const unsigned int* overflow;
  float* qs;
  __device__ void begin(int row, void* extra) {
    qs = reinterpret_cast<float*>(extra);
    if (overflow[row]) for (int t = threadIdx.x; t < d; t += blockDim.x) qs[t] = q[(long long)row * d + t];
  }
  __device__ long long count(int row) const { return overflow[row] ? N : 0; }
  __device__ void get(int, long long t, float& s, long long& i) const {
    const float* c = corpus + t * d;
    float acc = 0.f;
    for (int kk = 0; kk < d; ++kk) acc = fmaf(qs[kk], __ldg(c + kk), acc);
    s = acc; i = index_offset + t;
  }
};

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Optional per-stage device timing (bench.py's roofline leg): CUDA events recorded on the launch stream
// around the stages of tfrs_topk_tc_f32.  Off by default; adds two event records per stage when on.
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
  int stride, n_sample, n_bins, bins_ld, parts_sample, parts_full, cap_part;
  size_t smem;
  // workspace offsets
  size_t o_qstats, o_qimg, o_margin, o_cut, o_thr, o_count, o_ovf, o_binmax, o_cand, total;
};

static bool make_plan(long long Q, long long N, int d, int k, Plan& pl) {
  if (d <= 0 || d > 256 || Q <= 0 || N <= 0 || k <= 0) return false;
  pl.kb = (d + KSLAB - 1) / KSLAB;
  pl.stages = pl.kb == 1 ? 6 : 4;
  pl.n_tiles = ceil_div(N, TILE_N);
  pl.nqb = (int)ceil_div(Q, QBLK);
  pl.Qp = (long long)pl.nqb * QBLK;
  if (pl.kb > 2) return false;          // d > 128: smem budget (A blocks + ring) not laid out yet
  if (k > 256 || N >= (1ll << 31)) return false;   // survivor capacity (CAND_CAP) is sized for ~4k + band entries per query
  // sample every stride-th tile; keep at least 4k bins so the k-th largest bin max is a tight bound
  const long long full_tiles = N / TILE_N;  // the zero-padded last tile is never sampled (its 0 scores are not candidates)
  if (full_tiles < 1) return false;
  pl.stride = MAX_SAMPLE_STRIDE;
  {
    static int ov = -1;
    if (ov < 0) { const char* e = getenv("TFRS_TC_SAMPLE_STRIDE"); ov = e ? atoi(e) : 0; }
    if (ov >= 1 && ov <= 16) pl.stride = ov;
  }
  while (pl.stride > 1 && 2 * ceil_div(full_tiles, pl.stride) < 4ll * k) pl.stride >>= 1;
  pl.n_sample = (int)ceil_div(full_tiles, pl.stride);
  pl.n_bins = pl.n_sample * 2;
  pl.bins_ld = (pl.n_bins + 3) / 4 * 4;
  if (pl.n_bins < 4 * k) return false;  // too few bins for a useful threshold -> caller uses the exact path
  const int sms = sm_count();
  int parts = sms / pl.nqb; if (parts < 1) parts = 1; if (parts > FIN_MAX_PARTS / 2) parts = FIN_MAX_PARTS / 2;
  pl.parts_sample = parts < pl.n_sample ? parts : pl.n_sample;
  pl.parts_full = (long long)parts < pl.n_tiles ? parts : (int)pl.n_tiles;
  {
    int cp = 2 * CAND_CAP / (pl.parts_full * 2);  // octet records: (part, column-half) segments add up to ~2x the per-query capacity
    int p2 = 32; while (p2 * 2 <= cp && p2 < 512) p2 <<= 1;   // 32..512 records per segment
    pl.cap_part = p2;
  }
  pl.smem = (size_t)(2 + pl.stages) * pl.kb * SLAB_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 1024); return r; };
  pl.o_qstats = take(sizeof(SideStats));
  pl.o_qimg = take((size_t)pl.nqb * 2 * pl.kb * SLAB_BYTES);
  pl.o_margin = take((size_t)pl.Qp * 4);
  pl.o_cut = take((size_t)pl.Qp * 4);
  pl.o_thr = take((size_t)pl.Qp * 4);
  pl.o_count = take((size_t)pl.Qp * pl.parts_full * 2 * 4);
  pl.o_ovf = take((size_t)pl.Qp * 4);
  pl.o_binmax = take((size_t)pl.Qp * pl.bins_ld * 4);
  pl.o_cand = take((size_t)pl.Qp * pl.parts_full * 2 * pl.cap_part * (8 * 4 + 4));
  pl.total = o;
  return true;
}

template <int KB, int STAGES>
static int launch_scans(const Plan& pl, ScanParams sp, cudaStream_t st, int mode) {
  auto ks = tc_scan_kernel<KB, STAGES, MODE_SAMPLE>;
  auto kf = tc_scan_kernel<KB, STAGES, MODE_FILTER>;
  static bool attr = false;
  if (!attr) {
    TFRS_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    TFRS_CUDA(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    attr = true;
  }
  if (mode == MODE_SAMPLE) {
    sp.parts = pl.parts_sample; sp.n_seq = pl.n_sample; sp.stride = pl.stride;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("TFRS_TC_DEBUG_MODE"); dbg = e ? atoi(e) : 0; }
    if (dbg == 2) {
      auto k2 = tc_scan_kernel<KB, STAGES, MODE_DBG_LDONLY>;
      TFRS_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      k2<<<(unsigned)(pl.nqb * sp.parts), THREADS, pl.smem, st>>>(sp);
    } else if (dbg == 3) {
      auto k3 = tc_scan_kernel<KB, STAGES, MODE_DBG_NOLD>;
      TFRS_CUDA(cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      k3<<<(unsigned)(pl.nqb * sp.parts), THREADS, pl.smem, st>>>(sp);
    } else
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
  side_stats_kernel<<<(unsigned)(148 * 8), 256, 0, st>>>(corpus, N, d, cst, nullptr);
  TFRS_LAUNCH_CHECK();
  side_exp_kernel<<<1, 1, 0, st>>>(cst, fp16_target());
  TFRS_LAUNCH_CHECK();
  long long chunks = h.n_tiles * TILE_N * (long long)h.kb * 8;
  unsigned blocks = (unsigned)(ceil_div(chunks, 256) < (1 << 20) ? ceil_div(chunks, 256) : (1 << 20));
  tile_image_kernel<<<blocks, 256, 0, st>>>(corpus, N, d, h.kb, h.n_tiles, cst, use_bf16(), (unsigned char*)index_buf + HEADER_BYTES);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

extern "C" size_t tfrs_topk_tc_workspace_bytes(int64_t Q, int64_t N, int d, int k) {
  Plan pl;
  if (!make_plan(Q, N, d, k, pl)) return 0;
  return pl.total;
}

extern "C" int tfrs_topk_tc_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d,
                                int k, int64_t index_offset, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes,
                                void* stream) {
  TFRS_CHECK_ARG(q && corpus && index_buf && out_scores && out_idx, "topk_tc: NULL pointer");
  TFRS_CHECK_ARG(k <= N, "input must have at least k columns. Had %lld, needed %d", (long long)N, k);
  Plan pl;
  if (!make_plan(Q, N, d, k, pl)) {
    set_error("topk_tc: shape (Q=%lld N=%lld d=%d k=%d) is outside the tensor-core path; use tfrs_topk_scan_f32",
              (long long)Q, (long long)N, d, k);
    return TFRS_ERR_UNSUPPORTED;
  }
  if (!ws || ws_bytes < pl.total) { set_error("topk_tc: workspace too small (%zu < %zu)", ws_bytes, pl.total); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)(((uintptr_t)ws + 15) & ~(uintptr_t)15);
  unsigned char* qimg = w + pl.o_qimg;
  float* margin = (float*)(w + pl.o_margin);
  float* cut = (float*)(w + pl.o_cut);
  float* thr = (float*)(w + pl.o_thr);
  unsigned int* count = (unsigned int*)(w + pl.o_count);
  unsigned int* ovf = (unsigned int*)(w + pl.o_ovf);
  float* binmax = (float*)(w + pl.o_binmax);
  float* cand_s = (float*)(w + pl.o_cand);
  unsigned int* cand_i = (unsigned int*)(w + pl.o_cand + (size_t)pl.Qp * pl.parts_full * 2 * pl.cap_part * 32);
  const IndexHeader* hdr = (const IndexHeader*)index_buf;
  const unsigned char* cimg = (const unsigned char*)index_buf + HEADER_BYTES;

  prof_mark(st, 0);
  // (0) query statistics, image and margins
  {
    SideStats* qst = (SideStats*)(w + pl.o_qstats);
    TFRS_CUDA(cudaMemsetAsync(qst, 0, sizeof(SideStats), st));
    side_stats_kernel<<<(unsigned)ceil_div(Q * 32, 256), 256, 0, st>>>(q, Q, d, qst, margin /* scratch: |q|^2 per row */);
    TFRS_LAUNCH_CHECK();
    side_exp_kernel<<<1, 1, 0, st>>>(qst, fp16_target());
    TFRS_LAUNCH_CHECK();
    long long chunks = (long long)pl.nqb * 2 * TILE_N * pl.kb * 8;
    tile_image_kernel<<<(unsigned)ceil_div(chunks, 256), 256, 0, st>>>(q, Q, d, pl.kb, (long long)pl.nqb * 2, qst, use_bf16(), qimg);
    TFRS_LAUNCH_CHECK();
    qmargin_kernel<<<(unsigned)ceil_div(pl.Qp, 256), 256, 0, st>>>(Q, pl.Qp, hdr, qst, use_bf16(), margin, cut);
    TFRS_LAUNCH_CHECK();
  }
  ScanParams sp{};
  sp.qimg = qimg; sp.cimg = cimg; sp.Q = Q; sp.N = N; sp.nqb = pl.nqb; sp.n_tiles = pl.n_tiles;
  sp.binmax = binmax; sp.bins_ld = pl.bins_ld; sp.thr = thr; sp.count = count; sp.cand_s = cand_s; sp.cand_i = cand_i; sp.cap_part = pl.cap_part;
  sp.idesc = IDESC_F16_M128_N128 | (use_bf16() ? ((1u << 7) | (1u << 10)) : 0u);
  prof_mark(st, 1);
  // (1) sampled pass -> bin maxima -> k-th largest -> threshold
  int rc = launch_scan_mode(pl, sp, st, MODE_SAMPLE);
  if (rc) return rc;
  {
    size_t smem = pl.n_bins <= THR_SMEM_BINS ? (size_t)pl.n_bins * 4 : 0;
    static bool attr = false;
    if (!attr) { TFRS_CUDA(cudaFuncSetAttribute(tc_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, THR_SMEM_BINS * 4)); attr = true; }
    tc_threshold_kernel<<<(unsigned)Q, 256, smem, st>>>(binmax, pl.bins_ld, pl.n_bins, k, margin, thr, ovf);
  }
  TFRS_LAUNCH_CHECK();
  prof_mark(st, 2);
  // (2) full pass with the fused threshold filter
  rc = launch_scan_mode(pl, sp, st, MODE_FILTER);
  if (rc) return rc;
  prof_mark(st, 3);
  // (3) exact re-scoring + final order
  {
    size_t smem = (size_t)FIN_MAXM * 12 + (size_t)CAND_CAP * 8 + (size_t)d * 4 + 16;
    static bool attr = false;
    if (!attr) { TFRS_CUDA(cudaFuncSetAttribute(tc_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    tc_finalize_kernel<<<(unsigned)Q, 256, smem, st>>>(q, corpus, d, k, index_offset, count, cand_s, cand_i, pl.parts_full * 2, pl.cap_part,
                                                     cut, thr, N, ovf, out_scores, (long long*)out_idx);
    TFRS_LAUNCH_CHECK();
  }
  // (4) exact fallback for overflowed queries (CTAs of non-flagged queries exit immediately)
  {
    FallbackProvider fp{q, corpus, N, d, index_offset, ovf, nullptr};
    int cap = rowselect_cap(k);
    static bool attr = false;
    if (!attr) { TFRS_CUDA(cudaFuncSetAttribute(row_topk_kernel<FallbackProvider>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    row_topk_kernel<FallbackProvider><<<(unsigned)Q, RS_THREADS, rowselect_smem(cap, (size_t)d * 4), st>>>(
        fp, k, cap, out_scores, (long long*)out_idx, k);
    TFRS_LAUNCH_CHECK();
  }
  prof_mark(st, 4);
  return TFRS_OK;
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
