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

struct tfrs_comm {
  tfrs::ncclComm_t comm;
  int rank, world, device;
};

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
  c->rank = rank; c->world = world;
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

extern "C" int tfrs_comm_destroy(tfrs_comm_t c) {
  if (!c) return TFRS_OK;
  NcclApi* a = nccl_api();
  if (a && c->comm) a->CommDestroy(c->comm);
  delete c;
  return TFRS_OK;
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
  const int k_local = (int)(N_local < k ? N_local : k);
  if (k_local < k) {
    fill_pad_kernel<<<(unsigned)ceil_div(Q * (k - k_local), 256), 256, 0, st>>>(send_s, send_i, Q, k, k_local);
    TFRS_LAUNCH_CHECK();
  }
  if (k_local > 0) {
    int rc = TFRS_ERR_UNSUPPORTED;
    if (index_buf && k_local == k && tfrs_topk_tc_workspace_bytes(Q, N_local, d, k) > 0)
      rc = tfrs_topk_tc_f32(q, Q, corpus_local, index_buf, N_local, d, k, index_offset, send_s, (int64_t*)send_i, w, L.scan, st);
    if (rc == TFRS_ERR_UNSUPPORTED)
      rc = tfrs_topk_scan_f32(q, Q, corpus_local, N_local, d, k, index_offset, nullptr, nullptr, 0, send_s, (int64_t*)send_i, w, L.scan, st);
    if (rc) return rc;
  }
  TFRS_NCCL(a->AllGather(send, recv, L.block, NCCL_UINT8, c->comm, st));
  count_launch(1);
  return tfrs_topk_merge_sorted_strided((const float*)recv, (const int64_t*)(recv + L.idx_off), (int64_t)(L.block / 4),
                                        (int64_t)(L.block / 8), c->world, Q, k, k, out_scores, out_idx, st);
}
