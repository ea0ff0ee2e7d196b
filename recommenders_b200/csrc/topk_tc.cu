// placeholder until the tcgen05 path lands
#include "common.cuh"
using namespace tfrs;
extern "C" size_t tfrs_index_bytes(int64_t, int) { return 0; }
extern "C" int tfrs_index_build(const float*, int64_t, int, void*, size_t, void*) { set_error("tc path not built"); return TFRS_ERR_UNSUPPORTED; }
extern "C" size_t tfrs_topk_tc_workspace_bytes(int64_t, int64_t, int, int) { return 0; }
extern "C" int tfrs_topk_tc_f32(const float*, int64_t, const float*, const void*, int64_t, int, int, int64_t, float*, int64_t*, void*, size_t, void*) { set_error("tc path not built"); return TFRS_ERR_UNSUPPORTED; }
