// probe.cu -- hardware probe: one CTA, one 128x128x64 UMMA tile with a caller-chosen instruction
// descriptor, raw TMEM dump.  Used to pin down undocumented layouts (e.g. fp16 accumulators) before the
// production kernels rely on them.  Not on any product path.
#include "../../recommenders_b200/csrc/common.cuh"
#include "../../recommenders_b200/csrc/tc_ptx.cuh"

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

// ---- tensor-pipe / TMEM timing probe (tools/tc_rate_probe.py): cycle counts of the building blocks the softmax
// backward and the top-K filter are made of, one CTA per SM, operands are whatever the (zeroed) buffers hold.
//   mode 0: R x [16 warps: tcgen05.ld 32x32b.x64 + wait]                 -> TMEM->register rate, 32-bit columns
//   mode 1: R x [16 warps: tcgen05.ld 32x32b.x64.pack::16b + wait]       -> the same with 16-bit packing
//   mode 2: R x [12 SS MMAs  M128 N128 K16]                               (operands in shared memory)
//   mode 3: R x [24 TS MMAs  M128 N64  K16]  (A from TMEM, B MN-major)
//   mode 4: R x [12 SS (N128) then 24 TS (N64)]  interleaved, one commit at the end -> cost of SS<->TS hand-offs
//   mode 5: R x [16 warps: tcgen05.st 32x32b.x32 x2 + wait::st]
// out[blockIdx.x] = cycles (clock64) of the timed region on that SM.
namespace tfrs {
namespace tc {

__device__ __forceinline__ void tmem_ld64_pack16(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
}

__global__ void __launch_bounds__(640, 1)
tc_rate_probe_kernel(int mode, int rounds, long long* __restrict__ out, uint32_t* __restrict__ sink) {
  extern __shared__ __align__(1024) unsigned char rp_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(rp_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;            // 32 KB: hi | lo of a 128-row tile
  unsigned char* sB = smem + 32768;    // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  __shared__ long long t_begin_sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1 && lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 2) tmem_alloc(slot, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill visible to the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot;
  if (threadIdx.x == 0) t_begin_sh = clock64();
  __syncthreads();
  uint32_t keep = 0;
  if (mode == 0 || mode == 1 || mode == 5) {
    if (warp >= 4) {
      const int ew = warp - 4, quad = ew & 3;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((ew >> 2) * 128);
      for (int r0 = 0; r0 < rounds; ++r0) {
        uint32_t r[64];
        if (mode == 5) {
#pragma unroll
          for (int j = 0; j < 64; ++j) r[j] = keep + j;
          tmem_st32(taddr, r); tmem_st32(taddr + 32, r + 32);
          tmem_st_wait();
        } else {
          if (mode == 0) tmem_ld64(taddr + (r0 & 1) * 64, r); else tmem_ld64_pack16(taddr, r);
          tmem_ld_wait64(r);
          keep ^= r[0] ^ r[63];
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint64_t a_hi = make_smem_desc(a0), a_lo = make_smem_desc(a0 + 16384);
    const uint64_t b_hi = make_smem_desc(b0), b_lo = make_smem_desc(b0 + 16384);
    for (int r0 = 0; r0 < rounds; ++r0) {
      if (mode == 2 || mode == 4) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint64_t o = (uint64_t)(k4 * 2);
          umma_f16(tmem_base, a_hi + o, b_hi + o, IDESC_F16_M128_N128, (uint32_t)(k4 != 0));
          umma_f16(tmem_base, a_lo + o, b_hi + o, IDESC_F16_M128_N128, 1u);
          umma_f16(tmem_base, a_hi + o, b_lo + o, IDESC_F16_M128_N128, 1u);
        }
      }
      if (mode == 3 || mode == 4) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t ta_hi = tmem_base + 128u + (uint32_t)(64 * (j >> 2) + 8 * (j & 3)), ta_lo = ta_hi + 32;
          const uint64_t y_hi = make_smem_desc(b0 + j * 2048), y_lo = make_smem_desc(b0 + 16384 + j * 2048);
          umma_f16_ts(tmem_base + 384u, ta_hi, y_hi, IDESC_F16_M128_N64_BMN, (uint32_t)(j != 0));
          umma_f16_ts(tmem_base + 384u, ta_lo, y_hi, IDESC_F16_M128_N64_BMN, 1u);
          umma_f16_ts(tmem_base + 384u, ta_hi, y_lo, IDESC_F16_M128_N64_BMN, 1u);
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    tc_fence_after();
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t_begin_sh;
  if (keep == 0x12345678u) sink[0] = keep;
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace tfrs

extern "C" int tfrs_debug_tc_rate_probe(int mode, int rounds, int n_ctas, long long* out_cycles, uint32_t* sink, void* stream) {
  TFRS_CHECK_ARG(mode >= 0 && mode <= 5 && rounds > 0 && n_ctas > 0 && out_cycles && sink, "tc_rate_probe: bad argument");
  static bool attr = false;
  if (!attr) { TFRS_CUDA(cudaFuncSetAttribute(tfrs::tc::tc_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 2048)); attr = true; }
  tfrs::tc::tc_rate_probe_kernel<<<n_ctas, 640, 65536 + 2048, (cudaStream_t)stream>>>(mode, rounds, out_cycles, sink);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}
