/*
 * tfrs_oracle.c -- CPU restatement of the TensorFlow Recommenders retrieval hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under recommenders_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / the CPU arm -- never as the product.
 *
 * Parity pinning: the reference is pure Python on TensorFlow (not installable here), so this
 * oracle is pinned by the reference's OWN known-answer tests, re-expressed in
 * tests/test_oracle_golden.py (layers/factorized_top_k_test.py:31-147,
 * metrics/factorized_top_k_test.py:31-131, tasks/retrieval_test.py:31-298,
 * layers/loss_test.py:29-130, layers/feature_interaction/dcn_test.py:29-101).
 * The summation order of TF's SGEMM and the numerics of tf-keras Adagrad have no in-tree test:
 * for those two "parity unpinned" applies.
 *
 * Canonical arithmetic (what "bit-exact" means in this repo): every query x candidate score is
 * the sequential chain  acc = fmaf(q[k], c[k], acc), k = 0..d-1, acc starting at +0.0f.
 * Top-K order: score descending, equal scores -> lower candidate index first
 * (tf.math.top_k contract relied on by layers/factorized_top_k.py:605 and
 * metrics/factorized_top_k.py:155-161).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- canonical dot product: layers/factorized_top_k.py:320-333 (_compute_score = matmul(q, c^T)) ---- */
static inline float dot_chain(const float* q, const float* c, int d) {
  float acc = 0.0f;
  for (int k = 0; k < d; ++k) acc = fmaf(q[k], c[k], acc);
  return acc;
}

/* scores[Q,N] = q[Q,d] . c[N,d]^T  (tasks/retrieval.py:178-180, layers/factorized_top_k.py:333) */
void orc_scores(const float* q, int64_t Q, const float* c, int64_t N, int d, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < Q; ++i)
    for (int64_t j = 0; j < N; ++j) out[i * N + j] = dot_chain(q + i * d, c + j * d, d);
}

/* total order used everywhere: (score desc, index asc).  -0.0f == +0.0f as in TF's float compare. */
static inline int better(float sa, int64_t ia, float sb, int64_t ib) {
  if (sa > sb) return 1;
  if (sa < sb) return 0;
  return ia < ib;
}

typedef struct { float s; int64_t i; } pair_t;

static void sift_down(pair_t* h, int n, int p) {
  /* min-heap on the total order: root = worst of the kept set */
  for (;;) {
    int l = 2 * p + 1, r = l + 1, w = p;
    if (l < n && better(h[w].s, h[w].i, h[l].s, h[l].i)) w = l;
    if (r < n && better(h[w].s, h[w].i, h[r].s, h[r].i)) w = r;
    if (w == p) return;
    pair_t t = h[p]; h[p] = h[w]; h[w] = t; p = w;
  }
}

static int cmp_desc(const void* a, const void* b) {
  const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
  if (better(x->s, x->i, y->s, y->i)) return -1;
  if (better(y->s, y->i, x->s, x->i)) return 1;
  return 0;
}

/*
 * Brute-force scan with optional carried state (Streaming) and index offset (shards):
 *   BruteForce.call   layers/factorized_top_k.py:586-607  (matmul -> top_k -> gather ids)
 *   Streaming.call    layers/factorized_top_k.py:404-509  (per-chunk top_k, then merge with state;
 *                     state is concatenated BEFORE the new chunk, :462-463, so on equal scores the
 *                     earlier row wins -- identical to (score desc, index asc) with running indices)
 * Output k_out = min(k, state_k + N) entries per query, sorted; returns k_out.
 */
int orc_topk_scan(const float* q, int64_t Q, const float* c, int64_t N, int d, int k,
                  int64_t index_offset, const float* st_s, const int64_t* st_i, int st_k,
                  float* out_s, int64_t* out_i) {
  int64_t tot = (int64_t)st_k + N;
  int k_out = (int)(k < tot ? k : tot);
  if (k_out <= 0) return 0;
#pragma omp parallel
  {
    pair_t* h = (pair_t*)malloc(sizeof(pair_t) * (size_t)k_out);
#pragma omp for schedule(dynamic, 4)
    for (int64_t i = 0; i < Q; ++i) {
      int n = 0;
      for (int64_t t = 0; t < tot; ++t) {
        pair_t p;
        if (t < st_k) { p.s = st_s[i * st_k + t]; p.i = st_i[i * st_k + t]; }
        else { int64_t j = t - st_k; p.s = dot_chain(q + i * d, c + j * d, d); p.i = index_offset + j; }
        if (n < k_out) {
          h[n++] = p;
          if (n == k_out) for (int r = n / 2 - 1; r >= 0; --r) sift_down(h, n, r);
        } else if (better(p.s, p.i, h[0].s, h[0].i)) {
          h[0] = p; sift_down(h, n, 0);
        }
      }
      qsort(h, (size_t)n, sizeof(pair_t), cmp_desc);
      for (int r = 0; r < k_out; ++r) { out_s[i * k_out + r] = h[r].s; out_i[i * k_out + r] = h[r].i; }
    }
    free(h);
  }
  return k_out;
}

/* Merge n_lists sorted-or-not [Q,k_in] lists into the best k_out (shard merge / Streaming.reduce :440-472). */
int orc_topk_merge(const float* s, const int64_t* idx, int n_lists, int64_t Q, int k_in, int k_out,
                   float* out_s, int64_t* out_i) {
  int tot = n_lists * k_in;
  if (k_out > tot) k_out = tot;
  pair_t* buf = (pair_t*)malloc(sizeof(pair_t) * (size_t)tot);
  for (int64_t i = 0; i < Q; ++i) {
    for (int l = 0; l < n_lists; ++l)
      for (int r = 0; r < k_in; ++r) {
        buf[l * k_in + r].s = s[((int64_t)l * Q + i) * k_in + r];
        buf[l * k_in + r].i = idx[((int64_t)l * Q + i) * k_in + r];
      }
    qsort(buf, (size_t)tot, sizeof(pair_t), cmp_desc);
    for (int r = 0; r < k_out; ++r) { out_s[i * k_out + r] = buf[r].s; out_i[i * k_out + r] = buf[r].i; }
  }
  free(buf);
  return k_out;
}

/* Embedding lookup: tf.keras.layers.Embedding -> tf.gather (README.md:62-66,77-78). */
void orc_gather(const float* table, int64_t rows, int d, const int64_t* ids, int64_t n,
                float* out, int64_t out_ld, int64_t col_off) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t r = ids[i];
    if (r < 0 || r >= rows) { memset(out + i * out_ld + col_off, 0, sizeof(float) * (size_t)d); continue; }
    memcpy(out + i * out_ld + col_off, table + r * d, sizeof(float) * (size_t)d);
  }
}

/*
 * Sparse Adagrad on touched rows (optimizer chosen by the user, README.md:84; applied at
 * models/base.py:77-78).  tf-keras semantics restated (third party, "parity unpinned"):
 * duplicate ids are summed first (IndexedSlices dedupe), in order of occurrence; then
 *   acc += g*g ;  var -= lr * g / sqrt(acc + eps)        (eps_inside_sqrt = 1, Keras optimizers)
 *   acc += g*g ;  var -= lr * g / (sqrt(acc) + eps)      (eps_inside_sqrt = 0, legacy optimizers)
 * Each step is one IEEE fp32 operation, written without contraction.
 */
void orc_sparse_adagrad(float* table, float* accum, int64_t rows, int d, const int64_t* ids, int64_t n,
                        const float* grad, float lr, float eps, int eps_inside_sqrt) {
  /* visited[] marks ids already folded; O(n^2) worst case is fine for test sizes */
  char* done = (char*)calloc((size_t)n, 1);
  float* g = (float*)malloc(sizeof(float) * (size_t)d);
  for (int64_t i = 0; i < n; ++i) {
    if (done[i]) continue;
    int64_t r = ids[i];
    if (r < 0 || r >= rows) continue;
    for (int c = 0; c < d; ++c) g[c] = grad[i * d + c];
    for (int64_t j = i + 1; j < n; ++j)
      if (ids[j] == r) { done[j] = 1; for (int c = 0; c < d; ++c) g[c] = g[c] + grad[j * d + c]; }
    for (int c = 0; c < d; ++c) {
      float a = accum[r * d + c];
      float gg = g[c] * g[c];
      a = a + gg;
      accum[r * d + c] = a;
      float den = eps_inside_sqrt ? sqrtf(a + eps) : (sqrtf(a) + eps);
      float num = lr * g[c];
      table[r * d + c] = table[r * d + c] - num / den;
    }
  }
  free(done); free(g);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
