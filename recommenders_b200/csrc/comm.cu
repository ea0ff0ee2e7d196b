// comm.cu -- C1: the collective of the row-sharded scan, inside the C ABI.
//
// The reference keeps the whole corpus in one variable (layers/factorized_top_k.py:571-580) and has no sharded scan; its only
// collective helper is the unused _cross_replica_concat (tasks/retrieval.py:238-321).  Here the corpus is row-sharded
// over the GPUs of one box; every rank scans its shard and ONE NCCL all-gather moves the per-shard (score, index)
// top-K lists over NVLink; every rank merges them.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the
// host framework already loaded, torch's or TensorFlow's, else the system one), so the library has no link-time
// dependency on it and a binder needs nothing but this C ABI: no torch.distributed on the data path.
#include <dlfcn.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "common.cuh"

namespace tfrs {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess == 0
enum { NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_INT64 = 4 };   // ncclDataType_t values (stable across NCCL 2.x)
enum { NCCL_MAX = 2 };                                        // ncclRedOp_t

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  char why[256] = "";
};

NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {getenv("TFRS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
      snprintf(api.why, sizeof(api.why), "%s", dlerror());
    }
    if (!api.handle) return;
#define TFRS_SYM(field, name) *(void**)(&api.field) = dlsym(api.handle, name)
    TFRS_SYM(GetUniqueId, "ncclGetUniqueId"); TFRS_SYM(CommInitRank, "ncclCommInitRank"); TFRS_SYM(CommDestroy, "ncclCommDestroy");
    TFRS_SYM(AllGather, "ncclAllGather"); TFRS_SYM(AllReduce, "ncclAllReduce"); TFRS_SYM(GroupStart, "ncclGroupStart");
    TFRS_SYM(GroupEnd, "ncclGroupEnd"); TFRS_SYM(GetErrorString, "ncclGetErrorString"); TFRS_SYM(GetVersion, "ncclGetVersion");
#undef TFRS_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GroupStart || !api.GroupEnd) {
      snprintf(api.why, sizeof(api.why), "libnccl is missing a required symbol");
      api.handle = nullptr;
    }
  });
  return api.handle ? &api : nullptr;
}

#define TFRS_NCCL(expr)                                                                                        \
  do {                                                                                                         \
    ncclResult_t r__ = (expr);                                                                                 \
    if (r__ != 0) {                                                                                            \
      NcclApi* a__ = nccl_api();                                                                               \
      ::tfrs::set_error("%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #expr, (int)r__,                \
                        (a__ && a__->GetErrorString) ? a__->GetErrorString(r__) : "?");                      \
      return TFRS_ERR_NCCL;                                                                                    \
    }                                                                                                          \
  } while (0)

__global__ void fill_pad_kernel(float* __restrict__ s, long long* __restrict__ i, long long Q, int k, int from) {
  const long long n = Q * (k - from);
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
    const long long row = t / (k - from); const int c = from + (int)(t % (k - from));
    s[row * k + c] = -INFINITY; i[row * k + c] = LLONG_MAX;
  }
}

}  // namespace
}  // namespace tfrs

constexpr int TFRS_MAX_RANKS = 16;
struct tfrs_comm {
  tfrs::ncclComm_t comm;
  int rank, world, device;
  // peer-memory exchange (tfrs_comm_enable_p2p): xbuf[r] = rank r's exchange buffer as mapped into THIS process
  // (cudaIpc over NVLink/NVSwitch; xbuf[rank] is the local allocation)
  unsigned char* xbuf[TFRS_MAX_RANKS];
  size_t xbytes;
  unsigned int epoch;
  int thr_exchange;   // exchange a global score bound before the filter pass (tfrs_comm_set_option)
};

// ---------------------------------------------------------------------------------------------------------------
// Peer-memory exchange: the collective of the sharded scan as our own kernels over NVLink P2P stores.
//
// NCCL's all-gather moves every rank's whole [Q,k] list to every rank (world x 4.9 MB received at cfg2) and every rank
// then merges all Q queries.  Here query q has an OWNER rank (contiguous blocks of Qo = ceil(Q / world) queries):
//   A. every rank stores the slice of its local lists that belongs to owner o straight into o's exchange buffer
//      (st.global on the IPC-mapped peer pointer) and raises flagA[src] there            -> (world-1)/world of ONE list leaves a rank
//   B. the owner merges the world sorted lists of its Qo queries (tfrs_topk_merge_sorted_strided on local memory)
//   C. the owner stores its final [Qo,k] block into every rank's result area and raises flagB[owner] there
//   D. every rank waits for the world flagB's and copies the result area to the caller's buffers
// 8x less NVLink traffic and 8x less merge work per rank than all-gather + replicated merge.  Flags are call epochs
// (monotone, never reset); data -> __threadfence_system() -> flag on the writer, flag -> fence -> data on the reader.
// A buffer is only overwritten by call e+1 after its readers of call e have finished: writers of step A(e+1) have passed
// their own D(e), which needs the owner's C(e), which follows the owner's B(e) reads; writers of C(e+1) have passed
// B(e+1)'s wait on every rank's A(e+1), which follows that rank's D(e) reads.
// ---------------------------------------------------------------------------------------------------------------
namespace tfrs {
namespace {
constexpr size_t XFLAGS = 8192;                 // flagA[r] at 128*r, flagT[r] at 2048 + 128*r, flagB[r] at 4096 + 128*r, block counters at the end
struct XLayout { long long Qo; size_t a_s, a_i, b_s, b_i, t, total; };
XLayout x_layout(int world, long long Q, int k) {
  XLayout L;
  L.Qo = (Q + world - 1) / world;
  size_t o = XFLAGS;
  auto take = [&](size_t b) { size_t r = o; o += align_up(b, 256); return r; };
  L.a_s = take((size_t)world * L.Qo * k * 4);
  L.a_i = take((size_t)world * L.Qo * k * 8);
  L.b_s = take((size_t)world * L.Qo * k * 4);   // result area, padded to world * Qo rows
  L.b_i = take((size_t)world * L.Qo * k * 8);
  L.t = take((size_t)world * world * L.Qo * 4);     // threshold exchange: [src rank][query] lower bounds of the k-th best score
  L.total = o;
  return L;
}

struct PeerPtrs { unsigned char* p[TFRS_MAX_RANKS]; };

__device__ __forceinline__ void spin_until(const volatile unsigned int* flag, unsigned int epoch) {
  const long long t0 = clock64();
  while ((int)(*flag - epoch) < 0) {            // wrap-safe "flag < epoch"
    if (clock64() - t0 > 20000000000ll) __trap();   // ~10 s: a peer died -- fail loudly instead of hanging the GPU
    __nanosleep(64);
  }
}

// the last CTA of a grid to get here raises this rank's flag in every peer's buffer
__device__ __forceinline__ void signal_when_grid_done(PeerPtrs peers, int world, size_t flag_off, int rank, unsigned int epoch,
                                                      unsigned int* counter) {
  __threadfence_system();            // this CTA's peer stores are ordered before its arrival
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    last = (prev == gridDim.x - 1);
    if (last) *counter = 0;
  }
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < world)
      *reinterpret_cast<volatile unsigned int*>(peers.p[threadIdx.x] + flag_off + 128 * rank) = epoch;
  }
}

// T (before the filter pass): a GLOBAL lower bound of the k-th best exact score.  Shard r knows K candidates with screening
// score >= L_r, i.e. exact score >= B_r = (L_r - eps_r) / 2^se_r; max_r B_r bounds the global k-th best from below, so shard s
// may filter at  B * 2^se_s - eps_s - slack_s : every member of the GLOBAL top-k still passes on its shard, while the
// survivors per shard drop from ~4.5 k to ~4.5 k / world -- the select / re-score work of the finalize step shrinks with
// the shard instead of staying constant.  16 KB per peer; same flag protocol as the list exchange.
__global__ void __launch_bounds__(256)
x_thr_send_kernel(PeerPtrs peers, int world, int rank, const float* __restrict__ thr, const float* __restrict__ margin,
                  const float* __restrict__ cut, const int* __restrict__ qexp, const int* __restrict__ exp_corpus, long long Q,
                  size_t t_off, long long q_cap, unsigned int epoch) {
  const int ec = *exp_corpus;
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < Q; q += (long long)gridDim.x * 256) {
    const float L = thr[q] + margin[q];
    const float B = ldexpf(L - 0.5f * cut[q], -(ec + qexp[q]));
    for (int o = 0; o < world; ++o) reinterpret_cast<float*>(peers.p[o] + t_off)[(long long)rank * q_cap + q] = B;
  }
  signal_when_grid_done(peers, world, 2048, rank, epoch, reinterpret_cast<unsigned int*>(peers.p[rank] + XFLAGS - 192));
}

__global__ void __launch_bounds__(256)
x_thr_combine_kernel(const unsigned char* local, int world, float* __restrict__ thr, const float* __restrict__ margin,
                     const float* __restrict__ cut, const int* __restrict__ qexp, const int* __restrict__ exp_corpus, long long Q,
                     size_t t_off, long long q_cap, unsigned int epoch) {
  if ((int)threadIdx.x < world) spin_until(reinterpret_cast<const volatile unsigned int*>(local + 2048 + 128 * threadIdx.x), epoch);
  __threadfence_system();
  __syncthreads();
  const int ec = *exp_corpus;
  const float* T = reinterpret_cast<const float*>(local + t_off);
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < Q; q += (long long)gridDim.x * 256) {
    float B = -INFINITY;
    for (int r = 0; r < world; ++r) B = fmaxf(B, __ldcv(T + (long long)r * q_cap + q));
    const float eps = 0.5f * cut[q];
    const float t_new = ldexpf(B, ec + qexp[q]) - eps - (margin[q] - cut[q]);
    thr[q] = fmaxf(thr[q], t_new);   // never below the shard's own rule (equal to it on the shard that set the maximum)
  }
}

// A: owner o gets rows [o*Qo, (o+1)*Qo) of this rank's local lists, at list slot `rank` of its area A
__global__ void __launch_bounds__(256)
x_scatter_kernel(PeerPtrs peers, int world, int rank, const float* __restrict__ loc_s, const long long* __restrict__ loc_i,
                 long long Q, int k, XLayout L, unsigned int epoch) {
  const long long per_owner16 = L.Qo * k * 12 / 4;   // in 4-byte words: scores (1 word) + indices (2 words) per entry
  (void)per_owner16;
  for (int o = 0; o < world; ++o) {
    const long long q0 = (long long)o * L.Qo, q1 = min(Q, q0 + L.Qo);
    if (q1 <= q0) continue;
    const long long n = (q1 - q0) * k;
    float* ds = reinterpret_cast<float*>(peers.p[o] + L.a_s) + (long long)rank * L.Qo * k;
    long long* di = reinterpret_cast<long long*>(peers.p[o] + L.a_i) + (long long)rank * L.Qo * k;
    const float* ss = loc_s + q0 * k; const long long* si = loc_i + q0 * k;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) { ds[t] = ss[t]; di[t] = si[t]; }
  }
  signal_when_grid_done(peers, world, 0, rank, epoch, reinterpret_cast<unsigned int*>(peers.p[rank] + XFLAGS - 128));
}

__global__ void x_wait_kernel(const unsigned char* local, size_t flag_off, int world, unsigned int epoch) {
  if ((int)threadIdx.x < world) spin_until(reinterpret_cast<const volatile unsigned int*>(local + flag_off + 128 * threadIdx.x), epoch);
  __threadfence_system();
}

// C: this owner's merged [Qo_mine, k] block (already in its own result area) -> every other rank's result area
__global__ void __launch_bounds__(256)
x_bcast_kernel(PeerPtrs peers, int world, int rank, long long Q, int k, XLayout L, unsigned int epoch) {
  const long long q0 = (long long)rank * L.Qo, q1 = min(Q, q0 + L.Qo);
  const long long n = q1 > q0 ? (q1 - q0) * k : 0;
  const float* ss = reinterpret_cast<const float*>(peers.p[rank] + L.b_s) + q0 * k;
  const long long* si = reinterpret_cast<const long long*>(peers.p[rank] + L.b_i) + q0 * k;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
    const float v = ss[t]; const long long ix = si[t];
    for (int o = 0; o < world; ++o) {
      if (o == rank) continue;
      reinterpret_cast<float*>(peers.p[o] + L.b_s)[q0 * k + t] = v;
      reinterpret_cast<long long*>(peers.p[o] + L.b_i)[q0 * k + t] = ix;
    }
  }
  signal_when_grid_done(peers, world, 4096, rank, epoch, reinterpret_cast<unsigned int*>(peers.p[rank] + XFLAGS - 64));
}

// D: wait for every owner's block, then hand the result area to the caller's buffers
__global__ void __launch_bounds__(256)
x_collect_kernel(const unsigned char* local, int world, long long Q, int k, XLayout L, unsigned int epoch, float* __restrict__ out_s,
                 long long* __restrict__ out_i) {
  if ((int)threadIdx.x < world) spin_until(reinterpret_cast<const volatile unsigned int*>(local + 4096 + 128 * threadIdx.x), epoch);
  __threadfence_system();
  __syncthreads();
  const float* ss = reinterpret_cast<const float*>(local + L.b_s);
  const long long* si = reinterpret_cast<const long long*>(local + L.b_i);
  const long long n = Q * k;
  for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) { out_s[t] = ss[t]; out_i[t] = si[t]; }
}
}  // namespace
}  // namespace tfrs

using namespace tfrs;

extern "C" int tfrs_comm_unique_id(void* out128) {
  TFRS_CHECK_ARG(out128, "comm_unique_id: NULL pointer");
  NcclApi* a = nccl_api();
  if (!a) { set_error("comm: NCCL is not available (%s)", nccl_api() ? "" : "dlopen libnccl.so.2 failed"); return TFRS_ERR_NCCL; }
  ncclUniqueId id;
  TFRS_NCCL(a->GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return TFRS_OK;
}

extern "C" int tfrs_comm_create(tfrs_comm_t* out, int rank, int world, const void* unique_id128) {
  TFRS_CHECK_ARG(out && unique_id128 && world > 0 && rank >= 0 && rank < world, "comm_create: bad argument");
  NcclApi* a = nccl_api();
  if (!a) { set_error("comm: NCCL is not available (dlopen libnccl.so.2 failed)"); return TFRS_ERR_NCCL; }
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  tfrs_comm* c = new tfrs_comm{};
  c->rank = rank; c->world = world; c->thr_exchange = 1;
  if (cudaGetDevice(&c->device) != cudaSuccess) c->device = 0;
  ncclResult_t r = a->CommInitRank(&c->comm, world, id, rank);   // collective: every rank of the group calls it
  if (r != 0) {
    set_error("comm_create: ncclCommInitRank -> %d (%s)", (int)r, a->GetErrorString ? a->GetErrorString(r) : "?");
    delete c;
    return TFRS_ERR_NCCL;
  }
  *out = c;
  return TFRS_OK;
}

static void x_release(tfrs_comm* c) {
  for (int r = 0; r < c->world && r < TFRS_MAX_RANKS; ++r) {
    if (!c->xbuf[r]) continue;
    if (r == c->rank) cudaFree(c->xbuf[r]); else cudaIpcCloseMemHandle(c->xbuf[r]);
    c->xbuf[r] = nullptr;
  }
  c->xbytes = 0;
}

extern "C" int tfrs_comm_destroy(tfrs_comm_t c) {
  if (!c) return TFRS_OK;
  cudaDeviceSynchronize();
  x_release(c);
  NcclApi* a = nccl_api();
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
  return TFRS_OK;
}

// Collective over the group: (re)allocates the exchange buffers for calls up to (max_Q, max_k) and maps every peer's
// buffer into this process (cudaIpc; the handles travel through one NCCL all-gather).  Synchronises the device.
extern "C" int tfrs_comm_enable_p2p(tfrs_comm_t c, int64_t max_Q, int max_k) {
  TFRS_CHECK_ARG(c && max_Q > 0 && max_k > 0, "comm_enable_p2p: bad argument");
  if (c->world > TFRS_MAX_RANKS) { set_error("comm_enable_p2p: world=%d > %d", c->world, TFRS_MAX_RANKS); return TFRS_ERR_UNSUPPORTED; }
  NcclApi* a = nccl_api();
  if (!a) { set_error("comm: NCCL is not available"); return TFRS_ERR_NCCL; }
  TFRS_CUDA(cudaDeviceSynchronize());
  x_release(c);
  const XLayout L = x_layout(c->world, max_Q, max_k);
  void* local = nullptr;
  TFRS_CUDA(cudaMalloc(&local, L.total));
  TFRS_CUDA(cudaMemset(local, 0, L.total));
  cudaIpcMemHandle_t mine;
  TFRS_CUDA(cudaIpcGetMemHandle(&mine, local));
  cudaIpcMemHandle_t* dev = nullptr;
  TFRS_CUDA(cudaMalloc((void**)&dev, sizeof(cudaIpcMemHandle_t) * (c->world + 1)));
  TFRS_CUDA(cudaMemcpy(dev + c->world, &mine, sizeof(mine), cudaMemcpyHostToDevice));
  TFRS_NCCL(a->AllGather(dev + c->world, dev, sizeof(cudaIpcMemHandle_t), NCCL_UINT8, c->comm, (cudaStream_t)0));
  TFRS_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t all[TFRS_MAX_RANKS];
  TFRS_CUDA(cudaMemcpy(all, dev, sizeof(cudaIpcMemHandle_t) * c->world, cudaMemcpyDeviceToHost));
  cudaFree(dev);
  c->xbuf[c->rank] = (unsigned char*)local;
  c->xbytes = L.total;
  c->epoch = 0;
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("comm_enable_p2p: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
      c->xbuf[r] = nullptr;
      // every rank must agree on the outcome: the all-reduce below tells the others
    } else {
      c->xbuf[r] = (unsigned char*)p;
    }
  }
  // agree: if ANY rank failed to map a peer, every rank falls back to the NCCL path
  int* ok_dev = nullptr;
  TFRS_CUDA(cudaMalloc((void**)&ok_dev, 8));
  int ok = 1;
  for (int r = 0; r < c->world; ++r) if (!c->xbuf[r]) ok = 0;
  float okf = ok ? 0.f : 1.f;   // max over ranks of "failed"
  TFRS_CUDA(cudaMemcpy(ok_dev, &okf, 4, cudaMemcpyHostToDevice));
  if (a->AllReduce) TFRS_NCCL(a->AllReduce(ok_dev, ok_dev + 1, 1, NCCL_FLOAT32, NCCL_MAX, c->comm, (cudaStream_t)0));
  TFRS_CUDA(cudaDeviceSynchronize());
  float failed = 1.f;
  TFRS_CUDA(cudaMemcpy(&failed, ok_dev + 1, 4, cudaMemcpyDeviceToHost));
  cudaFree(ok_dev);
  if (failed != 0.f) { x_release(c); if (ok) set_error("comm_enable_p2p: a peer could not map the exchange buffers"); return TFRS_ERR_UNSUPPORTED; }
  return TFRS_OK;
}

// option 0 = threshold exchange of the peer-memory path (default on).  Must be set identically on every rank.
extern "C" int tfrs_comm_set_option(tfrs_comm_t c, int option, int value) {
  TFRS_CHECK_ARG(c && option == 0, "comm_set_option: unknown option %d", option);
  c->thr_exchange = value != 0;
  return TFRS_OK;
}

extern "C" int tfrs_comm_p2p_capacity(tfrs_comm_t c, int64_t Q, int k) {
  if (!c || !c->xbytes || Q <= 0 || k <= 0) return 0;
  return x_layout(c->world, Q, k).total <= c->xbytes ? 1 : 0;
}

extern "C" int tfrs_comm_rank(tfrs_comm_t c) { return c ? c->rank : -1; }
extern "C" int tfrs_comm_world(tfrs_comm_t c) { return c ? c->world : -1; }

// all_s [world, Q, k] / all_i [world, Q, k]: every rank's lists, in rank order.  One NCCL group = one fused launch.
extern "C" int tfrs_topk_allgather(tfrs_comm_t c, const float* s, const int64_t* i, int64_t Q, int k, float* all_s,
                                   int64_t* all_i, void* stream) {
  TFRS_CHECK_ARG(c && s && i && all_s && all_i && Q >= 0 && k > 0, "topk_allgather: bad argument");
  NcclApi* a = nccl_api();
  if (!a) { set_error("comm: NCCL is not available"); return TFRS_ERR_NCCL; }
  if (Q == 0) return TFRS_OK;
  TFRS_NCCL(a->GroupStart());
  ncclResult_t r1 = a->AllGather(s, all_s, (size_t)Q * k, NCCL_FLOAT32, c->comm, (cudaStream_t)stream);
  ncclResult_t r2 = a->AllGather(i, all_i, (size_t)Q * k, NCCL_INT64, c->comm, (cudaStream_t)stream);
  TFRS_NCCL(a->GroupEnd());
  TFRS_NCCL(r1); TFRS_NCCL(r2);
  count_launch(1);
  return TFRS_OK;
}

// ---- the whole sharded BruteForce call in one entry point ---------------------------------------------------------
// workspace = [local scan scratch | send block | receive blocks]; block = [scores f32 [Q,k] | pad to 8 | indices i64 [Q,k]]
namespace {
struct ShardLayout { size_t scan, idx_off, block, o_send, o_recv, total; };
ShardLayout shard_layout(int world, int64_t Q, int64_t N_local, int d, int k) {
  ShardLayout L;
  const size_t tc = tfrs_topk_tc_workspace_bytes(Q, N_local, d, k);
  const size_t ex = tfrs_topk_scan_workspace_bytes(Q, N_local, d, k);
  L.scan = align_up((tc > ex ? tc : ex) + 256, 1024);
  L.idx_off = align_up((size_t)Q * k * 4, 8);
  L.block = align_up(L.idx_off + (size_t)Q * k * 8, 16);
  L.o_send = L.scan;
  L.o_recv = L.o_send + align_up(L.block, 1024);
  L.total = L.o_recv + (size_t)world * L.block + 1024;
  return L;
}
}  // namespace

namespace tfrs {
namespace {
struct ThrCtx { tfrs_comm* c; unsigned int epoch; XLayout X; };
int thr_hook(void* vctx, float* thr, const float* margin, const float* cut, const int* qexp, const int* exp_corpus, long long Q,
             cudaStream_t st) {
  ThrCtx* t = (ThrCtx*)vctx;
  tfrs_comm* c = t->c;
  PeerPtrs peers{};
  for (int r = 0; r < c->world; ++r) peers.p[r] = c->xbuf[r];
  const long long q_cap = (long long)c->world * t->X.Qo;
  const unsigned grid = (unsigned)ceil_div(Q, 256) < 32u ? (unsigned)ceil_div(Q, 256) : 32u;
  x_thr_send_kernel<<<grid, 256, 0, st>>>(peers, c->world, c->rank, thr, margin, cut, qexp, exp_corpus, Q, t->X.t, q_cap, t->epoch);
  TFRS_LAUNCH_CHECK();
  x_thr_combine_kernel<<<grid, 256, 0, st>>>(c->xbuf[c->rank], c->world, thr, margin, cut, qexp, exp_corpus, Q, t->X.t, q_cap, t->epoch);
  TFRS_LAUNCH_CHECK();
  return TFRS_OK;
}
__global__ void fill_neutral_kernel(float* thr, float* margin, float* cut, int* qexp, int* ec, long long Q) {
  for (long long q = (long long)blockIdx.x * 256 + threadIdx.x; q < Q; q += (long long)gridDim.x * 256) { thr[q] = -INFINITY; margin[q] = 0.f; cut[q] = 0.f; qexp[q] = 0; }
  if (blockIdx.x == 0 && threadIdx.x == 0) *ec = 0;
}
int thr_hook_neutral(ThrCtx* t, long long Q, unsigned char* scratch, cudaStream_t st) {
  // scratch: the local-scan workspace (free at this point): 4 arrays of Q floats + 1 int
  float* thr = (float*)scratch; float* margin = thr + Q; float* cut = margin + Q; int* qexp = (int*)(cut + Q); int* ec = qexp + Q;
  fill_neutral_kernel<<<32, 256, 0, st>>>(thr, margin, cut, qexp, ec, Q);
  TFRS_LAUNCH_CHECK();
  return thr_hook(t, thr, margin, cut, qexp, ec, Q, st);
}
}  // namespace
}  // namespace tfrs

extern "C" size_t tfrs_topk_sharded_workspace_bytes(int world, int64_t Q, int64_t N_local, int d, int k) {
  if (world <= 0 || Q <= 0 || N_local < 0 || d <= 0 || k <= 0) return 0;
  return shard_layout(world, Q, N_local, d, k).total;
}

// Test introspection: the block layout of the sharded call, out4 = {idx byte offset inside a block, block bytes,
// send-block offset, receive-buffer offset} (offsets from the 1024-byte-aligned workspace base).
extern "C" int tfrs_topk_sharded_layout(int world, int64_t Q, int64_t N_local, int d, int k, int64_t* out4) {
  TFRS_CHECK_ARG(out4 && world > 0 && Q > 0 && d > 0 && k > 0 && N_local >= 0, "topk_sharded_layout: bad argument");
  const ShardLayout L = shard_layout(world, Q, N_local, d, k);
  out4[0] = (int64_t)L.idx_off; out4[1] = (int64_t)L.block; out4[2] = (int64_t)L.o_send; out4[3] = (int64_t)L.o_recv;
  return TFRS_OK;
}

extern "C" int tfrs_topk_sharded_f32(tfrs_comm_t c, const float* q, int64_t Q, const float* corpus_local, const void* index_buf,
                                     int64_t N_local, int d, int k, int64_t index_offset, float* out_scores, int64_t* out_idx,
                                     void* ws, size_t ws_bytes, void* stream) {
  TFRS_CHECK_ARG(c && q && out_scores && out_idx && Q > 0 && d > 0 && k > 0 && N_local >= 0, "topk_sharded: bad argument");
  NcclApi* a = nccl_api();
  if (!a) { set_error("comm: NCCL is not available"); return TFRS_ERR_NCCL; }
  const ShardLayout L = shard_layout(c->world, Q, N_local, d, k);
  if (!ws || ws_bytes < L.total) { set_error("topk_sharded: workspace too small (%zu < %zu)", ws_bytes, L.total); return TFRS_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  unsigned char* send = w + L.o_send;
  unsigned char* recv = w + L.o_recv;
  float* send_s = (float*)send; long long* send_i = (long long*)(send + L.idx_off);
  // local scan, written straight into the send block; shards shorter than k are padded with (-inf, INT64_MAX) so that
  // EVERY rank issues the identical collective whatever path its own shard takes
  const bool p2p = c->xbytes && x_layout(c->world, Q, k).total <= c->xbytes;
  const unsigned int epoch = p2p ? ++c->epoch : 0u;
  const int k_local = (int)(N_local < k ? N_local : k);
  // the threshold exchange is itself a collective step: every rank takes the same decision (it depends only on what all
  // ranks share: Q, k, world, the option); a shard that cannot use it (no tensor-core index) takes part with "no bound"
  const bool thr_x = p2p && c->thr_exchange;
  if (k_local < k) {
    fill_pad_kernel<<<(unsigned)ceil_div(Q * (k - k_local), 256), 256, 0, st>>>(send_s, send_i, Q, k, k_local);
    TFRS_LAUNCH_CHECK();
  }
  ThrCtx tctx{c, epoch, x_layout(c->world, Q, k)};
  bool hook_ran = false;
  if (k_local > 0) {
    int rc = TFRS_ERR_UNSUPPORTED;
    if (index_buf && k_local == k && tfrs_topk_tc_workspace_bytes(Q, N_local, d, k) > 0) {
      rc = tc_topk_sharded_local(q, Q, corpus_local, index_buf, N_local, d, k, index_offset, send_s, (int64_t*)send_i, w, L.scan, st,
                                 thr_x ? thr_hook : nullptr, thr_x ? &tctx : nullptr);
      hook_ran = thr_x && rc == TFRS_OK;
    }
    if (rc == TFRS_ERR_UNSUPPORTED)
      rc = tfrs_topk_scan_f32(q, Q, corpus_local, N_local, d, k, index_offset, nullptr, nullptr, 0, send_s, (int64_t*)send_i, w, L.scan, st);
    if (rc) return rc;
  }
  if (thr_x && !hook_ran) {
    // this shard took the exact path (tiny shard): it still has to take part in the threshold exchange -- it contributes
    // "no bound" (-inf) and ignores the result
    int rc = thr_hook_neutral(&tctx, Q, w, st);
    if (rc) return rc;
  }
  if (p2p) {
    // peer-memory exchange: scatter to the owners -> owners merge their block -> owners store the results everywhere
    const XLayout X = x_layout(c->world, Q, k);
    PeerPtrs peers{};
    for (int r = 0; r < c->world; ++r) peers.p[r] = c->xbuf[r];
    unsigned char* local = c->xbuf[c->rank];
    const unsigned grid = (unsigned)(sm_count() < 64 ? sm_count() : 64);
    x_scatter_kernel<<<grid, 256, 0, st>>>(peers, c->world, c->rank, send_s, send_i, Q, k, X, epoch);
    TFRS_LAUNCH_CHECK();
    x_wait_kernel<<<1, 32, 0, st>>>(local, 0, c->world, epoch);
    TFRS_LAUNCH_CHECK();
    const long long q0 = (long long)c->rank * X.Qo, q1 = Q < q0 + X.Qo ? Q : q0 + X.Qo;
    if (q1 > q0) {
      int rc = tfrs_topk_merge_sorted_strided((const float*)(local + X.a_s), (const int64_t*)(local + X.a_i), X.Qo * k, X.Qo * k, c->world,
                                              q1 - q0, k, k, (float*)(local + X.b_s) + q0 * k, (int64_t*)(local + X.b_i) + q0 * k, st);
      if (rc) return rc;
    }
    x_bcast_kernel<<<grid, 256, 0, st>>>(peers, c->world, c->rank, Q, k, X, epoch);
    TFRS_LAUNCH_CHECK();
    x_collect_kernel<<<grid, 256, 0, st>>>(local, c->world, Q, k, X, epoch, out_scores, (long long*)out_idx);
    TFRS_LAUNCH_CHECK();
    return TFRS_OK;
  }
  TFRS_NCCL(a->AllGather(send, recv, L.block, NCCL_UINT8, c->comm, st));
  count_launch(1);
  return tfrs_topk_merge_sorted_strided((const float*)recv, (const int64_t*)(recv + L.idx_off), (int64_t)(L.block / 4),
                                        (int64_t)(L.block / 8), c->world, Q, k, k, out_scores, out_idx, st);
}
