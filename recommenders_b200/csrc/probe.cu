// probe.cu -- hardware probe: one CTA, one 128x128x64 UMMA tile with a caller-chosen instruction
// descriptor, raw TMEM dump.  Used to pin down undocumented layouts (e.g. fp16 accumulators) before the
// production kernels rely on them.  Not on any product path.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace tfrs {
namespace tc {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const unsigned char* __restrict__ a_img, const unsigned char* __restrict__ b_img, uint32_t idesc,
                  int n_cols, uint32_t* __restrict__ out /*[128][128] raw TMEM words*/) {
  extern __shared__ __align__(1024) unsigned char psm_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(psm_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(&bars[0], 32768);
    bulk_g2s(sA, a_img, 16384, &bars[0]);
    bulk_g2s(sB, b_img, 16384, &bars[0]);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint64_t a_desc = make_smem_desc(smem_u32(sA));
    const uint64_t b_desc = make_smem_desc(smem_u32(sB));
    for (int k4 = 0; k4 < 4; ++k4) umma_f16(tmem_base, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), idesc, (uint32_t)(k4 != 0));
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int h = 0; h < n_cols / 64; ++h) {
    uint32_t r[64];
    tmem_ld64(taddr + h * 64, r);
    tmem_ld_wait64(r);
#pragma unroll
    for (int j = 0; j < 64; ++j) out[row * 128 + h * 64 + j] = r[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

}  // namespace tc
}  // namespace tfrs
using namespace tfrs;
using namespace tfrs::tc;

// a_img / b_img: 16 KB fp16 128x64 SWIZZLE_128B K-major tile images (the first tile of a tfrs_index image,
// after its 1024-byte header).  out: 128*128 uint32.
extern "C" int tfrs_debug_umma_probe(const void* a_img, const void* b_img, uint32_t idesc, int n_cols, uint32_t* out, void* stream) {
  TFRS_CHECK_ARG(a_img && b_img && out && (n_cols == 64 || n_cols == 128), "umma_probe: bad arguments");
  static bool attr = false;
  if (!attr) { TFRS_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000)); attr = true; }
  umma_probe_kernel<<<1, 128, 40000, (cudaStream_t)stream>>>((const unsigned char*)a_img, (const unsigned char*)b_img, idesc, n_cols, out);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}

// ---- HBM access-pattern probe (tools/hbm_probe.py): what the memory system gives for the gather's access pattern --
// rows of 128 B (32 floats); 8 threads x 16 B per row; 4 independent rows in flight per thread.
//   mode 0: copy            dst[i] = src[i]                   (sequential read + sequential write)
//   mode 1: random read     sink += src[hash(i)]              (random 128-byte rows, nothing written)
//   mode 2: write only      dst[i] = const                    (sequential 128-byte rows)
//   mode 3: random read + sequential write   dst[i] = src[hash(i)]
//   mode 4: random read + strided write      dst[(i % n_rows_out) * ld + (i / n_rows_out) * 32] = src[hash(i)]   (the gather layout)
namespace tfrs {
__global__ void __launch_bounds__(256)
hbm_probe_kernel(int mode, const float4* __restrict__ src, long long src_rows, float4* __restrict__ dst, long long n,
                 long long n_rows_out, long long ld4, float* __restrict__ sink) {
  const long long total = n * 8;  // 16-byte lanes
  const long long stride = (long long)gridDim.x * 256;
  float acc = 0.f;
  for (long long w = (long long)blockIdx.x * 256 + threadIdx.x; w < total; w += stride * 4) {
    float4 v[4]; long long o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long e = w + u * stride;
      o[u] = -1;
      if (e < total) {
        const long long i = e >> 3; const int l = (int)(e & 7);
        long long r = i;
        if (mode == 1 || mode >= 3) {  // cheap 32-bit mix, scaled into [0, src_rows) (src_rows < 2^32)
          unsigned int h = (unsigned int)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
          r = (long long)__umulhi(h, (unsigned int)src_rows);
        }
        if (mode != 2) v[u] = __ldg(src + r * 8 + l); else v[u] = make_float4(1.f, 2.f, 3.f, 4.f);
        o[u] = (mode == 4) ? ((i % n_rows_out) * ld4 + (i / n_rows_out) * 8 + l) : (i * 8 + l);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (o[u] >= 0) {
        if (mode == 1) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        else dst[o[u]] = v[u];
      }
    }
  }
  if (mode == 1 && acc == 123.456f) sink[0] = acc;
}
}  // namespace tfrs

extern "C" int tfrs_debug_hbm_probe(int mode, const void* src, int64_t src_rows, void* dst, int64_t n, int64_t n_rows_out,
                                    int64_t ld_floats, float* sink, void* stream) {
  TFRS_CHECK_ARG(mode >= 0 && mode <= 4 && src && dst && sink && n > 0 && src_rows > 0, "hbm_probe: bad argument");
  long long blocks = tfrs::ceil_div(n * 8, 256 * 4);
  tfrs::hbm_probe_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(mode, (const float4*)src, src_rows, (float4*)dst, n,
                                                                               n_rows_out > 0 ? n_rows_out : 1, ld_floats / 4, sink);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}
