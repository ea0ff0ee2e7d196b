/*
 * tfrs_b200.h -- C ABI of libtfrs_b200.so: the B200 (sm_100a) kernels behind the TensorFlow
 * Recommenders retrieval / ranking hot path.
 *
 * The reference (tensorflow/recommenders v0.7.7) has no native/FFI layer: its boundary with native
 * code is "call public tf.* ops" from Python.  Each entry point below therefore replaces one TF op
 * call site of the reference (cited per function, paths relative to tensorflow_recommenders/).
 * The Python mirror of the reference API (recommenders_b200/) is the only intended caller and binds
 * these symbols with ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *  - Every function returns 0 on success or a negative TFRS_ERR_* code; tfrs_last_error() returns a
 *    thread-local message.  No C++ exception or abort crosses the boundary.
 *  - All data pointers are DEVICE pointers owned by the caller (e.g. torch storage .data_ptr());
 *    the library never frees or retains them past the call.  Pointer ARRAYS (tables / ids lists)
 *    are HOST arrays of device pointers.
 *  - No hidden allocation: scratch memory is caller-provided, its size comes from *_workspace_bytes().
 *  - Every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *    No implicit synchronisation; calls are CUDA-graph capturable.
 *  - Matrices are row-major fp32 unless stated otherwise.
 */
#ifndef TFRS_B200_H_
#define TFRS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFRS_B200_VERSION 100 /* 0.1.0 */

enum {
  TFRS_OK = 0,
  TFRS_ERR_INVALID_ARG = -1,
  TFRS_ERR_UNSUPPORTED = -2,
  TFRS_ERR_CUDA = -3,
  TFRS_ERR_WORKSPACE_TOO_SMALL = -4,
  TFRS_ERR_NCCL = -5
};

enum { TFRS_I32 = 0, TFRS_I64 = 1 };

int tfrs_version(void);
const char* tfrs_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t tfrs_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  embedding gather.  Replaces tf.keras.layers.Embedding -> tf.gather in the user towers
 * (README.md:62-66,77-78; experimental/layers/embedding/partial_tpu_embedding.py:81-85,127).
 *   out[i, out_col_off[t] .. +dims[t]) = tables[t][ids[t][i], :]     for t < n_tables, i < n
 * Writes straight into a concatenated [n, out_ld] activation (the layout Cross consumes).
 * Out-of-range ids produce zero rows.  dims[t] % 4 == 0 and 16-byte aligned rows take the
 * vectorised path; anything else a scalar path.
 * ------------------------------------------------------------------------------------------- */
int tfrs_gather_f32(const float* const* tables, const int64_t* rows, const int32_t* dims, int n_tables,
                    const void* const* ids, int ids_dtype, int64_t n, float* out, int64_t out_ld,
                    const int32_t* out_col_off, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  brute-force top-K scan.  Replaces  scores = matmul(q, c^T); top_k(scores, k)
 * (layers/factorized_top_k.py:603-605 BruteForce.call; :424-438,:440-472 Streaming.call).
 * Scores are the canonical sequential fmaf chain over k = 0..d-1; order = (score desc, index asc).
 * `state_*` (nullable, state_k entries per query) is Streaming's carried state; it takes part in
 * the selection with its own indices.  New candidates get index = index_offset + row (shard /
 * chunk offset).  Output: min(k, state_k + N) entries per query, written with row stride k.
 * k <= 2048.
 *
 * tfrs_topk_scan_f32      exact fp32 CUDA-core path (any N, d; used for small corpora/chunks).
 * tfrs_index_*            builds the tensor-core screening image of a corpus (fp16 after an exact
 *                         power-of-two rescale, UMMA SWIZZLE_128B K-major tiles + row-norm bound);
 *                         done once at index() time (BruteForce.index, factorized_top_k.py:540-584).
 * tfrs_topk_tc_f32        tcgen05 screening GEMM (fp16 in / fp32 accumulate in TMEM) with a fused
 *                         threshold filter, then exact fp32 rescoring of the survivors -- same
 *                         bit-exact result as tfrs_topk_scan_f32.  Needs d <= 128, k <= 256 and a
 *                         corpus of at least ~256*k rows (tfrs_topk_tc_workspace_bytes returns 0
 *                         outside the supported range; callers then use tfrs_topk_scan_f32).
 * ------------------------------------------------------------------------------------------- */
size_t tfrs_topk_scan_workspace_bytes(int64_t Q, int64_t N, int d, int k);
int tfrs_topk_scan_f32(const float* q, int64_t Q, const float* corpus, int64_t N, int d, int k,
                       int64_t index_offset, const float* state_scores, const int64_t* state_idx,
                       int state_k, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes,
                       void* stream);

size_t tfrs_index_bytes(int64_t N, int d);
int tfrs_index_build(const float* corpus, int64_t N, int d, void* index_buf, size_t index_bytes,
                     void* stream);
size_t tfrs_topk_tc_workspace_bytes(int64_t Q, int64_t N, int d, int k);
int tfrs_topk_tc_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N,
                     int d, int k, int64_t index_offset, float* out_scores, int64_t* out_idx, void* ws,
                     size_t ws_bytes, void* stream);

/* The same tensor-core scan with `query_with_exclusions` fused into the finalize step (layers/factorized_top_k.py
 * :242-288 + `_exclude` :83-115): the k + n_excl best candidates are selected as above, candidates whose identifier
 * (identifiers[global index], or the global index itself when identifiers == NULL) appears in exclusions[q, :] get
 * score - 1e5, the k best ADJUSTED scores win (ties -> better original position) and the ORIGINAL scores / global
 * indices are written -- bit-for-bit what the reference computes from its over-fetched list.  int64 ids.
 * Workspace: tfrs_topk_tc_workspace_bytes(Q, N, d, k + n_excl). */
int tfrs_topk_tc_exclude_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d,
                             int k, int64_t index_offset, const int64_t* identifiers, const int64_t* exclusions,
                             int n_excl, float* out_scores, int64_t* out_idx, void* ws, size_t ws_bytes, void* stream);

/* The score branch of FactorizedTopK.update_state (metrics/factorized_top_k.py:133-137,181-192) without a top-K
 * list: out_count[q] = min(k, #{candidates whose exact score is > positive_scores[q]}); the metric for any k' <= k is
 * then  count < k'  (tf.math.in_top_k: fewer than k' predictions strictly above the target).  Same screening +
 * exact-rescoring guarantees as tfrs_topk_tc_f32 (only candidates within the error band of the positive are
 * re-scored).  Workspace: tfrs_topk_tc_workspace_bytes(Q, N, d, k). */
int tfrs_topk_tc_count_f32(const float* q, int64_t Q, const float* corpus, const void* index_buf, int64_t N, int d,
                           int k, const float* positive_scores, int32_t* out_count, void* ws, size_t ws_bytes,
                           void* stream);

/* `_exclude` (layers/factorized_top_k.py:83-115) on an already fetched, sorted [Q, k_fetched] list (Streaming and
 * the exact CUDA-core path): same rule as tfrs_topk_tc_exclude_f32.  idx are global indices into `identifiers`. */
int tfrs_topk_exclude_rerank_f32(const float* scores, const int64_t* idx, int64_t Q, int k_fetched,
                                 const int64_t* identifiers, const int64_t* exclusions, int n_excl, int k_out,
                                 float* out_scores, int64_t* out_idx, void* stream);

/* out_count[q] = #{t < k : scores[q*ld + t] > positive_scores[q]} -- the in_top_k count on a retrieved list. */
int tfrs_count_above_f32(const float* scores, int64_t ld, int k, const float* positive_scores, int64_t Q,
                         int32_t* out_count, void* stream);

/* Device-resident running sums of FactorizedTopK's Mean metrics (metrics/factorized_top_k.py:186-192):
 *   acc[j] += sum_q w_q * [count_q < ks[j] && isfinite(positive_q)]  (j < n_ks);   acc[n_ks] += sum_q w_q
 * (w = 1 when sample_weight == NULL).  `ks` is a HOST array (n_ks <= 16), `acc` n_ks + 1 doubles on the device.
 * Fixed-order fp64 reduction: deterministic; the host reads acc once, in result(). */
int tfrs_topk_hits_accumulate(const int32_t* count, const float* positive_scores, const float* sample_weight, int64_t Q,
                              const int32_t* ks, int n_ks, double* acc, void* stream);

/* Test/debug introspection of tfrs_topk_tc_f32's workspace: out8 = {count offset, fallback-flag offset,
 * threshold offset, survivor-list offset, parts, cap_part, padded Q, cut offset} (byte offsets from the
 * 16-byte-aligned workspace base).  Lets the tests assert that the exact fallback was NOT what produced a
 * result. */
int tfrs_topk_tc_layout(int64_t Q, int64_t N, int d, int k, int64_t* out8);

/* Optional per-stage device timing of tfrs_topk_tc_f32 (CUDA events on the launch stream; used by
 * bench.py for the roofline figure).  tfrs_profile_read synchronises the device and returns the summed
 * times in ms of stage 0 = query image, 1 = sampled pass + threshold, 2 = full filter pass (the
 * dominant kernel), 3 = exact re-scoring (+ fallback), over `calls` recorded calls. */
int tfrs_profile_enable(int on);
int tfrs_profile_read(float* stage_ms, int* calls);

/* K2m  merge n_lists per-shard/per-chunk [Q, k_in] lists (list-major: [n_lists, Q, k_in]) into the
 * best k_out = min(k_out, n_lists*k_in) per query (Streaming reduce :440-472; shard merge after the
 * all-gather).  Order = (score desc, index asc). */
int tfrs_topk_merge(const float* scores, const int64_t* idx, int n_lists, int64_t Q, int k_in, int k_out,
                    float* out_scores, int64_t* out_idx, void* stream);

/* Same merge for lists that sit `list_stride_*` elements apart (e.g. the receive buffer of the single
 * all-gather, where every rank's block is [scores | indices]). */
int tfrs_topk_merge_strided(const float* scores, const int64_t* idx, int64_t list_stride_scores,
                            int64_t list_stride_idx, int n_lists, int64_t Q, int k_in, int k_out,
                            float* out_scores, int64_t* out_idx, void* stream);

/* The same merge when every input list is already sorted in the total order (score desc, index asc) -- which is
 * what tfrs_topk_scan_f32 / tfrs_topk_tc_f32 emit, i.e. the per-shard lists of the sharded BruteForce: merged rank
 * = own position + binary-search counts in the other lists; no sort.  n_lists * k_in <= 16384. */
int tfrs_topk_merge_sorted_strided(const float* scores, const int64_t* idx, int64_t list_stride_scores,
                                   int64_t list_stride_idx, int n_lists, int64_t Q, int k_in, int k_out,
                                   float* out_scores, int64_t* out_idx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * C1  the collective of the row-sharded scan (the reference has no sharded scan -- its corpus is one
 * variable, layers/factorized_top_k.py:571-580 -- and its only collective helper is tasks/retrieval.py:238-321).
 * One process per GPU; shard g owns a contiguous row block, so global index order == (shard, local index) order and
 * the lowest-index tie rule survives the merge.  NCCL is bound at run time (dlopen libnccl.so.2: the copy the host
 * framework already loaded, else the system one; TFRS_NCCL_LIB overrides), so binders need only this header.
 *
 *   tfrs_comm_unique_id   rank 0 creates the 128-byte NCCL id; the host framework broadcasts it by any means
 *   tfrs_comm_create      collective over the group (ncclCommInitRank on the CURRENT device); tfrs_comm_destroy frees it
 *   tfrs_topk_allgather   every rank's [Q,k] (scores, global indices) -> all_s/all_i [world, Q, k] in rank order
 *   tfrs_topk_sharded_f32 the whole sharded BruteForce.call in one entry point: local scan (tensor-core path when
 *                         `index_buf` is given and the shape allows it, exact scan otherwise; shards shorter than k
 *                         are padded with (-inf, INT64_MAX)) written straight into the send block -> ONE all-gather
 *                         of the packed blocks -> sorted-list merge on every rank.  Every rank returns the same
 *                         [Q,k] result.  All ranks must call it with the same Q, d, k.
 * Handles need external locking; calls are asynchronous on `stream`.
 * ------------------------------------------------------------------------------------------- */
typedef struct tfrs_comm* tfrs_comm_t;
int tfrs_comm_unique_id(void* out128);
int tfrs_comm_create(tfrs_comm_t* comm, int rank, int world, const void* unique_id128);
int tfrs_comm_destroy(tfrs_comm_t comm);
/* Optional: map every rank's exchange buffer into every peer (cudaIpc over NVLink / NVSwitch), sized for calls up to
 * (max_Q, max_k).  Collective; synchronises the device.  With it tfrs_topk_sharded_f32 replaces the NCCL all-gather +
 * replicated merge by its own kernels: every rank STORES the slice of its lists owned by rank o straight into o's buffer
 * (owner = contiguous block of ceil(Q / world) queries), owners merge only their block and store the result into every
 * rank's result area, epoch flags order the steps -- 1/world of the all-gather's NVLink traffic and of the merge work.
 * Returns TFRS_ERR_UNSUPPORTED on EVERY rank when any peer mapping fails (the NCCL path stays in use).
 * tfrs_comm_p2p_capacity: 1 when the mapped buffers hold a (Q, k) call. */
int tfrs_comm_enable_p2p(tfrs_comm_t comm, int64_t max_Q, int max_k);
int tfrs_comm_p2p_capacity(tfrs_comm_t comm, int64_t Q, int k);
/* option 0: exchange a GLOBAL lower bound of the k-th best score between the threshold kernel and the filter pass of the
 * peer-memory path (default 1): each shard then keeps ~1/world of the survivors, so the per-rank select / re-score work
 * shrinks with the shard.  Must be set identically on every rank. */
int tfrs_comm_set_option(tfrs_comm_t comm, int option, int value);
int tfrs_comm_rank(tfrs_comm_t comm);
int tfrs_comm_world(tfrs_comm_t comm);
int tfrs_topk_allgather(tfrs_comm_t comm, const float* s, const int64_t* i, int64_t Q, int k, float* all_s,
                        int64_t* all_i, void* stream);
size_t tfrs_topk_sharded_workspace_bytes(int world, int64_t Q, int64_t N_local, int d, int k);
/* test introspection: out4 = {index byte offset inside a block, block bytes, send-block offset, receive-buffer offset} */
int tfrs_topk_sharded_layout(int world, int64_t Q, int64_t N_local, int d, int k, int64_t* out4);
int tfrs_topk_sharded_f32(tfrs_comm_t comm, const float* q, int64_t Q, const float* corpus_local, const void* index_buf,
                          int64_t N_local, int d, int k, int64_t index_offset, float* out_scores, int64_t* out_idx,
                          void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Score helpers (exact fp32, canonical fmaf chain, one owner thread per output).
 * tfrs_sgemm_f32: C[M,N] (+)= opA(A) . opB(B); opA(m,k) = transA ? A[k*lda+m] : A[m*lda+k],
 *   opB(k,n) = transB ? B[n*ldb+k] : B[k*ldb+n].  transA=0, transB=1 is `_compute_score`
 *   = matmul(q, c^T) (layers/factorized_top_k.py:320-333; tasks/retrieval.py:178-180); the other
 *   modes serve its backward and the low-rank Cross (dcn.py:131-148,178-179).
 * tfrs_rowwise_dot_f32: out[i] = sum_k a[i,k]*b[i,k]  (positive scores,
 *   metrics/factorized_top_k.py:133-134).
 * ------------------------------------------------------------------------------------------- */
int tfrs_sgemm_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float* C, int64_t ldc, int accumulate, void* stream);
int tfrs_rowwise_dot_f32(const float* a, const float* b, int64_t rows, int d, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  in-batch softmax loss of tfrs.tasks.Retrieval (tasks/retrieval.py:178-185,187-188,210):
 *   s = (q . c^T) * inv_temperature ; loss = sum_i w_i * (logsumexp_j s_ij - s_ii)
 * C >= B (extra negatives, positives are the first B rows).  `lse` [B] is saved for backward.
 * Backward (tape.gradient at models/base.py:77): G = (softmax(s) - I) * w * grad_loss * inv_temperature,
 *   dq = G . c   [B,d] ;  dc = G^T . q   [C,d].
 * ------------------------------------------------------------------------------------------- */
size_t tfrs_inbatch_softmax_workspace_bytes(int64_t B, int64_t C, int d);
int tfrs_inbatch_softmax_fwd(const float* q, const float* c, int64_t B, int64_t C, int d,
                             float inv_temperature, const float* sample_weight, float* loss, float* lse,
                             void* ws, size_t ws_bytes, void* stream);
int tfrs_inbatch_softmax_bwd(const float* q, const float* c, int64_t B, int64_t C, int d,
                             float inv_temperature, const float* sample_weight, const float* lse,
                             const float* grad_loss, float* dq, float* dc, void* ws, size_t ws_bytes,
                             void* stream);

/* K3 forward on the tensor cores (same contract and outputs as tfrs_inbatch_softmax_fwd; d <= 128):
 * hi/lo fp16 split of q and c (|err| <= 2^-21 |q||c| on a score), tcgen05 GEMM with fp32 TMEM accumulation and
 * an online log-sum-exp epilogue -- the [B,C] logits are never written.  `candidate_bias` (nullable, [C]) is added
 * to every logit of its column after the temperature: with bias_j = -log(clip(p_j, 1e-6, 1)) it is the
 * sampling-probability correction of tasks/retrieval.py:190-192 / layers/loss.py:150-158.  Returns TFRS_ERR_UNSUPPORTED
 * (workspace_bytes == 0) outside its shape range; the caller then uses tfrs_inbatch_softmax_fwd. */
size_t tfrs_inbatch_softmax_tc_workspace_bytes(int64_t B, int64_t C, int d);
int tfrs_inbatch_softmax_tc_fwd(const float* q, const float* c, int64_t B, int64_t C, int d,
                                float inv_temperature, const float* sample_weight, const float* candidate_bias,
                                float* loss, float* lse, void* ws, size_t ws_bytes, void* stream);

/* K3b on the tensor cores (same contract and outputs as tfrs_inbatch_softmax_bwd; d <= 64): two launches of one
 * flash-attention-backward-shaped kernel -- S = X.Y^T (tcgen05, split fp16 operands), G built from TMEM by the
 * epilogue warps and written back over S (tcgen05.st), dX += G.Y with G read from TMEM and the Y tile as an MN-major
 * operand; X = q gives dq, X = c gives dc.  Deterministic (no atomics).  workspace_bytes == 0 / TFRS_ERR_UNSUPPORTED
 * outside its range. */
size_t tfrs_inbatch_softmax_tc_bwd_workspace_bytes(int64_t B, int64_t C, int d);
int tfrs_inbatch_softmax_tc_bwd(const float* q, const float* c, int64_t B, int64_t C, int d,
                                float inv_temperature, const float* sample_weight, const float* candidate_bias,
                                const float* lse, const float* grad_loss, float* dq, float* dc, void* ws,
                                size_t ws_bytes, void* stream);

/* The remaining tfrs.tasks.Retrieval loss options inside the tensor-core loss, forward and backward:
 *   candidate_ids  (nullable, int64 [C])     remove_accidental_hits: every candidate j != i whose id equals the id of query
 *                                            i's positive (candidate i) gets logit MIN_FLOAT (tasks/retrieval.py:194-200,
 *                                            layers/loss.py:114-147; `logits + dup * MIN_FLOAT` rounds to MIN_FLOAT in fp32)
 *   score_mask     (nullable, uint8 [B, C])  where(mask, s, MIN_FLOAT) (retrieval.py:202-203); row-major, nonzero = keep
 * applied after the temperature and the bias, in the reference's order.  The ids / keep-bits are tested against the fp32
 * accumulators in registers: no [B,C] logits, labels or masks are materialised (the byte mask is packed to bits once).
 * Masked entries get zero gradient.  Same shape limits as the plain entry points; *_ex_workspace_bytes sizes `ws`. */
size_t tfrs_inbatch_softmax_tc_ex_workspace_bytes(int64_t B, int64_t C, int d, int has_ids, int has_mask);
int tfrs_inbatch_softmax_tc_fwd_ex(const float* q, const float* c, int64_t B, int64_t C, int d, float inv_temperature,
                                   const float* sample_weight, const float* candidate_bias, const int64_t* candidate_ids,
                                   const uint8_t* score_mask, float* loss, float* lse, void* ws, size_t ws_bytes,
                                   void* stream);
size_t tfrs_inbatch_softmax_tc_bwd_ex_workspace_bytes(int64_t B, int64_t C, int d, int has_ids, int has_mask);
int tfrs_inbatch_softmax_tc_bwd_ex(const float* q, const float* c, int64_t B, int64_t C, int d, float inv_temperature,
                                   const float* sample_weight, const float* candidate_bias, const int64_t* candidate_ids,
                                   const uint8_t* score_mask, const float* lse, const float* grad_loss, float* dq, float* dc,
                                   void* ws, size_t ws_bytes, void* stream);

/* Multi-head queries (tasks/retrieval.py:172-176): q is [B,H,d] and  scores_ij = max_h q_ih . c_j  ("maxsim") before the same
 * loss.  Exact fp32 path: row blocks of the [B*H, C] head scores stay L2-resident, the head maximum is folded while the
 * row statistics are taken; nothing of size [B,C] reaches the caller.  The gradient goes to the head(s) attaining the maximum
 * (split evenly among exact ties, as tf.reduce_max's).  dq is [B,H,d]. */
size_t tfrs_inbatch_softmax_maxsim_workspace_bytes(int64_t B, int H, int64_t C, int d);
int tfrs_inbatch_softmax_maxsim_fwd(const float* q, const float* c, int64_t B, int H, int64_t C, int d, float inv_temperature,
                                    const float* sample_weight, float* loss, float* lse, void* ws, size_t ws_bytes, void* stream);
int tfrs_inbatch_softmax_maxsim_bwd(const float* q, const float* c, int64_t B, int H, int64_t C, int d, float inv_temperature,
                                    const float* sample_weight, const float* lse, const float* grad_loss, float* dq, float* dc,
                                    void* ws, size_t ws_bytes, void* stream);

/* Hard-negative mining inside the loss (tasks/retrieval.py:205-210, layers/loss.py:61-111) without the [B,C] logits: the
 * n + 1 best candidates of every query come from the top-K scan above (k1 = min(n + 1, C) entries per query, exact fp32
 * scores, sorted); tfrs_hardneg_loss_fwd keeps the positive (candidate i of query i, score `positive_scores[i]`) plus the
 * n best other candidates and computes  loss = sum_i w_i (logsumexp(kept logits / T) - positive / T)  and the gradient
 * coefficients `coef` [B, k1 + 2] (entry t of the list, then the positive, then the row loss; w_i / T folded in).
 * tfrs_hardneg_loss_bwd:  dq_i = g sum_t coef_it c_{j_t}  (fixed order),  dc_j += g coef_it q_i  (float atomics: the only
 * non-bit-reproducible kernel of the library; dc is zeroed by the call). */
int tfrs_hardneg_loss_fwd(const float* top_scores, const int64_t* top_idx, int64_t B, int k1, const float* positive_scores,
                          float inv_temperature, const float* sample_weight, float* loss, float* coef, void* stream);
int tfrs_hardneg_loss_bwd(const float* q, const float* c, int64_t B, int64_t C, int d, const int64_t* top_idx, int k1,
                          const float* coef, const float* grad_loss, float* dq, float* dc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  sparse Adagrad on the rows touched by a batch (optimizer.apply_gradients with IndexedSlices,
 * models/base.py:77-78; Adagrad chosen by the user, README.md:84).  Duplicate ids are summed in
 * order of occurrence, then  acc += g*g ; var -= lr*g / sqrt(acc+eps)   (eps_inside_sqrt != 0)
 *                       or   acc += g*g ; var -= lr*g / (sqrt(acc)+eps) (eps_inside_sqrt == 0).
 * Deterministic (sort + segmented reduction, no float atomics).  n < 2^24.
 * ------------------------------------------------------------------------------------------- */
size_t tfrs_sparse_adagrad_workspace_bytes(int64_t n, int d);
int tfrs_sparse_adagrad_f32(float* table, float* accum, int64_t rows, int d, const void* ids, int ids_dtype,
                            int64_t n, const float* grad_rows, float lr, float eps, int eps_inside_sqrt,
                            void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5  DCN-v2 cross layer (layers/feature_interaction/dcn.py:176-186, full-rank, no preactivation):
 *   out = x0 * (x . W + bias + diag_scale * x) + x ,  W [D,D] in Keras [in,out] layout.
 * x0, x, out have row stride ld (>= D).  `prod` (nullable) receives x.W + bias + diag_scale*x for
 * the backward pass.  Backward:
 *   dx0 = g*prod ; gp = g*x0 ; dx = gp . W^T + diag_scale*gp + g ; dW = x^T . gp ; dbias = colsum(gp).
 * dx0/dx/dW/dbias are nullable (skipped when NULL).  ws holds gp: B*D floats.
 * ------------------------------------------------------------------------------------------- */
int tfrs_cross_fwd_f32(const float* x0, const float* x, const float* W, const float* bias, int64_t B, int D,
                       int64_t ld, float diag_scale, float* out, float* prod, void* stream);
size_t tfrs_cross_bwd_workspace_bytes(int64_t B, int D);
int tfrs_cross_bwd_f32(const float* x0, const float* x, const float* W, const float* prod, const float* dout,
                       int64_t B, int D, int64_t ld, float diag_scale, float* dx0, float* dx, float* dW,
                       float* dbias, void* ws, size_t ws_bytes, void* stream);

/* K5 on the tensor cores (forward): the same cross formula as tfrs_cross_fwd_f32, computed as one tcgen05
 * GEMM on exactly-rescaled fp16 hi/lo splits of x and W (3 MMAs per K step, fp32 accumulation in TMEM;
 * ~2^-21 relative error, inside the 1e-5 bar) with the formula fused in the epilogue.
 * tfrs_cross_tc_weight_build turns W [D,D] ([in,out]) into the K-major image of W^T; rebuild it whenever
 * W changes.  `ws` holds the per-call image of x. */
size_t tfrs_cross_tc_weight_bytes(int D);
int tfrs_cross_tc_weight_build(const float* W, int D, void* wbuf, size_t bytes, void* stream);
size_t tfrs_cross_tc_workspace_bytes(int64_t B, int D);
int tfrs_cross_tc_fwd_f32(const float* x0, const float* x, const void* wbuf, const float* bias, int64_t B, int D,
                          int64_t ld, float diag_scale, float* out, float* prod, void* ws, size_t ws_bytes,
                          void* stream);
/* The same forward for a STACK of cross layers (the reference chains `x = cross(x0, x)`): `out_amax_bits` (nullable, one
 * uint32 on the device) receives max |out| as float bits, accumulated by the epilogue; passing it as `x_amax_bits` of the
 * next layer replaces that layer's pass over x for the power-of-two rescale statistic (identical bits, identical result). */
int tfrs_cross_tc_fwd_ex_f32(const float* x0, const float* x, const void* wbuf, const float* bias, int64_t B, int D,
                             int64_t ld, float diag_scale, float* out, float* prod, const unsigned int* x_amax_bits,
                             unsigned int* out_amax_bits, void* ws, size_t ws_bytes, void* stream);

/* K5b with both GEMMs on the tensor cores (same contract and outputs as tfrs_cross_bwd_f32): dx = gp.W^T + diag*gp + g
 * and dW = x^T.gp as split-fp16 tcgen05 GEMMs (dW accumulates the batch in chunks of 1024 rows, partials summed in
 * fixed order); gp = g*x0, dx0 = g*prod and dbias = colsum(gp) as in the exact path.  Deterministic. */
size_t tfrs_cross_tc_bwd_workspace_bytes(int64_t B, int D);
int tfrs_cross_tc_bwd_f32(const float* x0, const float* x, const float* W, const float* prod, const float* dout,
                          int64_t B, int D, int64_t ld, float diag_scale, float* dx0, float* dx, float* dW,
                          float* dbias, void* ws, size_t ws_bytes, void* stream);

/* General fp32-parity GEMM on the tensor cores (the same exact-rescale + fp16 hi/lo split scheme, ~2^-21 relative error):
 *   C[M,N] = opA(A) . opB(B),  opA(m,k) = transA ? A[k*lda+m] : A[m*lda+k],  opB(k,n) = transB ? B[n*ldb+k] : B[k*ldb+n].
 * Reductions longer than 1024 are accumulated in chunks of 1024 with a fixed-order sum of the partials (deterministic).
 * Serves the projections of the low-rank Cross below and large `_compute_score`-style products. */
size_t tfrs_gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int tfrs_gemm_tc_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float* C, int64_t ldc, void* ws, size_t ws_bytes, void* stream);

/* Low-rank DCN-v2 cross layer on the tensor cores (layers/feature_interaction/dcn.py:131-148,178-179 `projection_dim`;
 * the layer of MultiLayerDCN, multi_layer_dcn.py:146-148):
 *   t = x . U  [B,p] ;  out = x0 * (t . V + bias + diag_scale * x) + x       U [D,p], V [p,D] in Keras [in,out] layout
 * Two tcgen05 GEMMs; the cross formula is the epilogue of the second one (no [B,D] product round trip).  `t` (required)
 * and `prod` (nullable) are kept for the backward pass:
 *   gp = g*x0 ; dx0 = g*prod ; dt = gp.V^T ; dV = t^T.gp ; dU = x^T.dt ; dx = dt.U^T + diag_scale*gp + g ; dbias = colsum(gp)
 * -- four tensor-core GEMMs, deterministic.  D, p <= 1024.  dx0 / dx / dU / dV / dbias are nullable. */
size_t tfrs_cross_lowrank_tc_workspace_bytes(int64_t B, int D, int p);
int tfrs_cross_lowrank_tc_fwd_f32(const float* x0, const float* x, const float* U, const float* V, const float* bias, int64_t B,
                                  int D, int p, int64_t ld, float diag_scale, float* out, float* prod, float* t, void* ws,
                                  size_t ws_bytes, void* stream);
size_t tfrs_cross_lowrank_tc_bwd_workspace_bytes(int64_t B, int D, int p);
int tfrs_cross_lowrank_tc_bwd_f32(const float* x0, const float* x, const float* U, const float* V, const float* t,
                                  const float* prod, const float* dout, int64_t B, int D, int p, int64_t ld, float diag_scale,
                                  float* dx0, float* dx, float* dU, float* dV, float* dbias, void* ws, size_t ws_bytes,
                                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * DLRM DotInteraction (layers/feature_interaction/dot_interaction.py:53-104): feats [B,F,d] ->
 * pairwise dots e_i.e_j of every sample; output = lower triangle in (i,j) row-major order without
 * (self_interaction=0) or with the diagonal, [B, out_dim], or the full [B,F*F] matrix with the excluded part
 * zeroed (skip_gather=1).  Every dot is the canonical sequential fmaf chain.  F <= 64.
 * Backward: dfeats[b,i,:] = sum_j G'(i,j) feats[b,j,:] with G' the symmetrised upstream gradient.
 * ------------------------------------------------------------------------------------------- */
int tfrs_dot_interaction_out_dim(int F, int self_interaction, int skip_gather);
int tfrs_dot_interaction_fwd_f32(const float* feats, int64_t B, int F, int d, int self_interaction, int skip_gather,
                                 float* out, void* stream);
int tfrs_dot_interaction_bwd_f32(const float* feats, const float* gout, int64_t B, int F, int d, int self_interaction,
                                 int skip_gather, float* dfeats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TFRS_B200_H_ */
