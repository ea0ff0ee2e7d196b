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
