"""CPU oracle for the TFRS retrieval/ranking hot path (NumPy + the C restatement in tfrs_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product (recommenders_b200/) never does.

Every function cites the reference file:line it restates (paths relative to
/root/reference/tensorflow_recommenders/).  The reference is pure Python on TensorFlow which
cannot be installed here, so this oracle is pinned by the reference's own known-answer tests
(tests/test_oracle_golden.py).  TF SGEMM summation order and tf-keras Adagrad numerics have no
in-tree test: "parity unpinned" for those two.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MIN_FLOAT = np.float32(np.finfo(np.float32).min / 100.0)  # layers/loss.py:23, tasks/retrieval.py:25
MAX_FLOAT = np.float32(np.finfo(np.float32).max / 100.0)  # layers/loss.py:22


def build(force: bool = False) -> str:
  """Compiles oracle/tfrs_oracle.c -> libtfrs_oracle.so (gcc); returns the path."""
  so = os.path.join(_HERE, "libtfrs_oracle.so")
  src = os.path.join(_HERE, "tfrs_oracle.c")
  if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["make", "-s", "-C", _HERE, "libtfrs_oracle.so"])
  return so


def _lib():
  global _LIB
  if _LIB is None:
    _LIB = ctypes.CDLL(build())
    _LIB.orc_topk_scan.restype = ctypes.c_int
    _LIB.orc_topk_merge.restype = ctypes.c_int
    _LIB.orc_num_threads.restype = ctypes.c_int
  return _LIB


def _p(a: Optional[np.ndarray]):
  return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a) -> np.ndarray:
  return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
  return int(_lib().orc_num_threads())


# ----------------------------------------------------------------------------------------------
# scores / top-k   (layers/factorized_top_k.py)
# ----------------------------------------------------------------------------------------------
def scores(q, c) -> np.ndarray:
  """`_compute_score` = matmul(q, c^T)  (factorized_top_k.py:320-333), canonical fmaf chain."""
  q, c = _f32(q), _f32(c)
  out = np.empty((q.shape[0], c.shape[0]), np.float32)
  _lib().orc_scores(_p(q), ctypes.c_int64(q.shape[0]), _p(c), ctypes.c_int64(c.shape[0]),
                    ctypes.c_int(q.shape[1]), _p(out))
  return out


def topk_scan(q, c, k: int, index_offset: int = 0, state: Optional[Tuple[np.ndarray, np.ndarray]] = None
              ) -> Tuple[np.ndarray, np.ndarray]:
  """BruteForce.call (:586-607) / one Streaming step (:424-472): exact top-k, (score desc, idx asc)."""
  q, c = _f32(q), _f32(c)
  Q, N = q.shape[0], c.shape[0]
  if state is not None and state[0].shape[1] > 0:
    st_s = _f32(state[0]); st_i = np.ascontiguousarray(state[1], np.int64); st_k = st_s.shape[1]
  else:
    st_s = st_i = None; st_k = 0
  k_out = min(k, st_k + N)
  out_s = np.empty((Q, k_out), np.float32); out_i = np.empty((Q, k_out), np.int64)
  if k_out == 0:
    return out_s, out_i
  r = _lib().orc_topk_scan(_p(q), ctypes.c_int64(Q), _p(c), ctypes.c_int64(N), ctypes.c_int(q.shape[1]),
                           ctypes.c_int(k), ctypes.c_int64(index_offset), _p(st_s), _p(st_i),
                           ctypes.c_int(st_k), _p(out_s), _p(out_i))
  assert r == k_out
  return out_s, out_i


def topk_merge(score_lists: np.ndarray, idx_lists: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
  """Merge [L,Q,k_in] per-shard lists -> [Q,min(k,L*k_in)] (shard merge; Streaming.reduce :440-472)."""
  s = _f32(score_lists); i = np.ascontiguousarray(idx_lists, np.int64)
  L, Q, k_in = s.shape
  k_out = min(k, L * k_in)
  out_s = np.empty((Q, k_out), np.float32); out_i = np.empty((Q, k_out), np.int64)
  _lib().orc_topk_merge(_p(s), _p(i), ctypes.c_int(L), ctypes.c_int64(Q), ctypes.c_int(k_in),
                        ctypes.c_int(k_out), _p(out_s), _p(out_i))
  return out_s, out_i


def top_k_rows(values: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
  """tf.math.top_k on a materialised matrix: sorted desc, equal values -> lower index first."""
  v = np.asarray(values)
  order = np.argsort(-v.astype(np.float64), axis=1, kind="stable")[:, :k]
  return np.take_along_axis(v, order, 1), order


def brute_force(q, candidates, identifiers=None, k: int = 10):
  """BruteForce.index + call (:540-607): default identifiers = range(N) int32 (:544-545)."""
  c = _f32(candidates)
  if c.ndim != 2:
    raise ValueError(f"The candidates tensor must be 2D (got {c.shape}).")
  ids = np.arange(c.shape[0], dtype=np.int32) if identifiers is None else np.asarray(identifiers)
  if ids.shape[0] != c.shape[0]:
    raise ValueError("The candidates and identifiers tensors must have the same number of rows")
  s, i = topk_scan(q, c, k)
  return s, ids[i]


def streaming(q, chunks: Iterable, k: int = 10, handle_incomplete_batches: bool = True):
  """Streaming.call (:404-509): per-chunk top-k then running merge; ids = running int32 counter
  (:474-485) or the dataset's identifiers (:486-488)."""
  q = _f32(q)
  state_s = np.zeros((q.shape[0], 0), np.float32)
  state_i = np.zeros((q.shape[0], 0), np.int64)
  counter = 0
  id_chunks = []
  has_ids = None
  for el in chunks:
    if isinstance(el, tuple):
      ids, emb = el; has_ids = True
      id_chunks.append(np.asarray(ids))
    else:
      emb = el; has_ids = False
    emb = _f32(emb)
    if not handle_incomplete_batches and (emb.shape[0] < k):
      raise ValueError(f"Tried to retrieve k={k} top items, but the candidate dataset batch size is too small.")
    state_s, state_i = topk_scan(q, emb, k, index_offset=counter, state=(state_s, state_i))
    counter += emb.shape[0]
  if has_ids:
    all_ids = np.concatenate(id_chunks, 0)
    return state_s, all_ids[state_i]
  return state_s, state_i.astype(np.int32)


def exclude(scores_: np.ndarray, identifiers: np.ndarray, exclude_ids: np.ndarray, k: int):
  """`_exclude` (:83-115): subtract 1e5 from excluded ids' scores, top_k(min(k, cols)) on the
  adjusted scores, return ORIGINAL scores/ids at those positions."""
  isin = (identifiers[:, :, None] == exclude_ids[:, None, :]).any(-1)
  adjusted = scores_ - isin.astype(np.float32) * np.float32(1.0e5)
  k = min(k, scores_.shape[1])
  _, idx = top_k_rows(adjusted, k)
  return np.take_along_axis(scores_, idx, 1), np.take_along_axis(identifiers, idx, 1)


def query_with_exclusions(layer_fn, q, exclusions: np.ndarray, k: int):
  """TopK.query_with_exclusions (:242-288): over-fetch k+E then `_exclude`."""
  s, i = layer_fn(q, k + exclusions.shape[1])
  return exclude(s, i, exclusions, k)


# ----------------------------------------------------------------------------------------------
# metrics   (metrics/factorized_top_k.py)
# ----------------------------------------------------------------------------------------------
def in_top_k(targets: np.ndarray, predictions: np.ndarray, k: int) -> np.ndarray:
  """tf.math.in_top_k: true iff prediction[target] is finite and #{pred > pred[target]} < k."""
  p = np.asarray(predictions, np.float32)
  t = p[np.arange(p.shape[0]), targets]
  return np.isfinite(t) & ((p > t[:, None]).sum(1) < k)


def factorized_top_k_update(q, true_emb, topk_fn, ks: Sequence[int], true_ids=None, sample_weight=None):
  """FactorizedTopK.update_state (:91-194).  Returns per-k (weighted_sum, weight_sum) increments.

  positive score = reduce_sum(q * c_true) (:133-134) -- restated with the canonical chain so that
  the positive's score compares consistently with the retrieved scores."""
  q = _f32(q); true_emb = _f32(true_emb)
  pos = np.array([[np.float32(scores(q[i:i + 1], true_emb[i:i + 1])[0, 0])] for i in range(q.shape[0])],
                 np.float32).reshape(-1, 1)
  top_s, top_ids = topk_fn(q, max(ks))
  w = np.ones((q.shape[0],), np.float32) if sample_weight is None else _f32(sample_weight).reshape(-1)
  out = []
  if true_ids is not None:
    true_ids = np.asarray(true_ids).reshape(-1, 1)
    nan_pad = np.isnan(top_s)
    match = ((true_ids == top_ids) & ~nan_pad).astype(np.float32)
    for k in ks:
      found = np.clip(match[:, :k].sum(1), 0.0, 1.0)
      out.append((float((found * w).sum()), float(w.sum())))
  else:
    y = np.concatenate([pos, top_s], 1)
    for k in ks:
      acc = in_top_k(np.zeros(q.shape[0], np.int64), y, k).astype(np.float32)
      out.append((float((acc * w).sum()), float(w.sum())))
  return out


# ----------------------------------------------------------------------------------------------
# loss transforms   (layers/loss.py)
# ----------------------------------------------------------------------------------------------
def sampling_probability_correction(logits, prob):
  """SamplingProbablityCorrection (loss.py:150-158)."""
  return _f32(logits) - np.log(np.clip(_f32(prob), np.float32(1e-6), np.float32(1.0)))


def remove_accidental_hits(labels, logits, candidate_ids):
  """RemoveAccidentalHits (loss.py:114-147)."""
  labels = _f32(labels); logits = _f32(logits); ids = np.asarray(candidate_ids)
  pos = labels.argmax(1)
  dup = (ids[pos][:, None] == ids[None, :]).astype(np.float32) - labels
  return logits + dup * MIN_FLOAT


def hard_negative_mining(logits, labels, num_hard_negatives: int):
  """HardNegativeMining (loss.py:61-111): keep positive + n hardest negatives per row."""
  logits = _f32(logits); labels = _f32(labels)
  n = min(num_hard_negatives + 1, logits.shape[1])
  _, cols = top_k_rows(logits + labels * MAX_FLOAT, n)
  return np.take_along_axis(logits, cols, 1), np.take_along_axis(labels, cols, 1)


# ----------------------------------------------------------------------------------------------
# Retrieval task   (tasks/retrieval.py:121-235)
# ----------------------------------------------------------------------------------------------
def retrieval_scores(q, c, temperature=None, candidate_sampling_probability=None, candidate_ids=None,
                     remove_accidental_hits_=False, score_mask=None, num_hard_negatives=None):
  q = np.asarray(q, np.float32); c = _f32(c)
  if q.ndim == 3:  # multi-head maxsim, retrieval.py:172-176
    s = np.stack([scores(q[:, h, :], c) for h in range(q.shape[1])], 1).max(1)
  else:
    s = scores(q, c)
  labels = np.eye(s.shape[0], s.shape[1], dtype=np.float32)  # :185
  if temperature is not None:
    s = s / np.float32(temperature)  # :187-188
  if candidate_sampling_probability is not None:
    s = sampling_probability_correction(s, candidate_sampling_probability)  # :190-192
  if remove_accidental_hits_:
    if candidate_ids is None:
      raise ValueError("When accidental hit removal is enabled, candidate ids must be supplied.")
    s = remove_accidental_hits(labels, s, candidate_ids)  # :194-200
  if score_mask is not None:
    s = np.where(np.asarray(score_mask, bool), s, MIN_FLOAT)  # :202-203
  if num_hard_negatives is not None:
    s, labels = hard_negative_mining(s, labels, num_hard_negatives)  # :205-208
  return s.astype(np.float32), labels


def softmax_xent_sum(logits, labels, sample_weight=None) -> float:
  """CategoricalCrossentropy(from_logits=True, reduction=SUM) (retrieval.py:86-87,210) in float64:
  sum_i w_i * (logsumexp(s_i) - sum_j y_ij s_ij)."""
  s = np.asarray(logits, np.float64); y = np.asarray(labels, np.float64)
  m = s.max(1, keepdims=True)
  lse = m[:, 0] + np.log(np.exp(s - m).sum(1))
  per = lse * y.sum(1) - (y * s).sum(1)
  if sample_weight is not None:
    per = per * np.asarray(sample_weight, np.float64).reshape(-1)
  return float(per.sum())


def retrieval_loss(q, c, sample_weight=None, **kw) -> float:
  s, y = retrieval_scores(q, c, **kw)
  return softmax_xent_sum(s, y, sample_weight)


def retrieval_loss_grads(q, c, sample_weight=None, temperature=None):
  """d loss / d q, d c for the default path, float64 (G = (softmax - I) * w / T; dq = G c; dc = G^T q)."""
  q64 = np.asarray(q, np.float64); c64 = np.asarray(c, np.float64)
  s = q64 @ c64.T
  t = 1.0 if temperature is None else float(temperature)
  s = s / t
  m = s.max(1, keepdims=True)
  p = np.exp(s - m); p /= p.sum(1, keepdims=True)
  g = p - np.eye(*s.shape)
  w = np.ones(s.shape[0]) if sample_weight is None else np.asarray(sample_weight, np.float64).reshape(-1)
  g = g * w[:, None] / t
  return g @ c64, g.T @ q64


def retrieval_loss_and_grads_general(q, c, sample_weight=None, temperature=None, candidate_sampling_probability=None,
                                     candidate_ids=None, remove_accidental_hits_=False, score_mask=None,
                                     num_hard_negatives=None):
  """float64 loss, d loss/d q, d loss/d c of Retrieval.call with every option (tasks/retrieval.py:178-210, layers/loss.py),
  2-D queries.  Gradient rules follow the TF ops: `logits + dup * MIN_FLOAT` passes the gradient (the probability of such
  an entry is 0), `tf.where(mask, s, MIN_FLOAT)` blocks it, hard-negative mining gathers: only the selected columns count."""
  q64 = np.asarray(q, np.float64); c64 = np.asarray(c, np.float64)
  B, C = q64.shape[0], c64.shape[0]
  t = 1.0 if temperature is None else float(temperature)
  s = (q64 @ c64.T) / t
  eye = np.eye(B, C)
  if candidate_sampling_probability is not None:
    s = s - np.log(np.clip(np.asarray(candidate_sampling_probability, np.float64), 1e-6, 1.0))[None, :]
  blocked = np.zeros((B, C), bool)
  if remove_accidental_hits_:
    ids = np.asarray(candidate_ids)
    dup = (ids[:B][:, None] == ids[None, :]) & (eye == 0)
    s = s + dup * float(MIN_FLOAT)
  if score_mask is not None:
    keep = np.asarray(score_mask, bool)
    s = np.where(keep, s, float(MIN_FLOAT))
    blocked |= ~keep
  active = np.ones((B, C), bool)
  if num_hard_negatives is not None:
    n = min(num_hard_negatives + 1, C)
    key = s + eye * float(MAX_FLOAT)
    order = np.lexsort((np.broadcast_to(np.arange(C), (B, C)), -key), axis=1)[:, :n]   # value desc, index asc
    active = np.zeros((B, C), bool)
    np.put_along_axis(active, order, True, 1)
  sm = np.where(active, s, -np.inf)
  m = sm.max(1, keepdims=True)
  e = np.exp(sm - m)
  z = e.sum(1, keepdims=True)
  w = np.ones(B) if sample_weight is None else np.asarray(sample_weight, np.float64).reshape(-1)
  loss = float((w * ((m[:, 0] - (eye * s).sum(1)) + np.log(z[:, 0]))).sum())   # max-subtracted, as TF's fused op
  g = (e / z - eye) * active
  g = np.where(blocked, 0.0, g) * (w[:, None] / t)
  return loss, g @ c64, g.T @ q64


# ----------------------------------------------------------------------------------------------
# Cross layer   (layers/feature_interaction/dcn.py:151-186)
# ----------------------------------------------------------------------------------------------
def cross(x0, x, W, bias=None, diag_scale: float = 0.0, U=None, V=None, preactivation=None):
  """y = x0 * (act(x @ W + b) + diag_scale * x) + x ; W is [in, out] (Keras Dense layout).
  Low-rank: x @ U[D,p] @ V[p,D] (dcn.py:131-148,178-179).  float64 accumulate -> float32."""
  x0_ = np.asarray(x0, np.float64); x_ = x0_ if x is None else np.asarray(x, np.float64)
  if x0_.shape[-1] != x_.shape[-1]:
    raise ValueError("`x0` and `x` dimension mismatch!")
  if U is not None:
    prod = (x_ @ np.asarray(U, np.float64)) @ np.asarray(V, np.float64)
  else:
    prod = x_ @ np.asarray(W, np.float64)
  if bias is not None:
    prod = prod + np.asarray(bias, np.float64)
  if preactivation is not None:
    prod = preactivation(prod)
  if diag_scale:
    prod = prod + diag_scale * x_
  return (x0_ * prod + x_).astype(np.float32)


def multi_layer_dcn(x0, Us, Vs, biases=None):
  """x_{l+1} = x0 * ((x_l @ U_l) @ V_l + b_l) + x_l  (layers/feature_interaction/multi_layer_dcn.py:136-153); float64."""
  x0_ = np.asarray(x0, np.float64); xl = x0_
  for l in range(len(Us)):
    prod = (xl @ np.asarray(Us[l], np.float64)) @ np.asarray(Vs[l], np.float64)
    if biases is not None:
      prod = prod + np.asarray(biases[l], np.float64)
    xl = x0_ * prod + xl
  return xl.astype(np.float32)


def dot_interaction(inputs, self_interaction: bool = False, skip_gather: bool = False):
  """DLRM dot interaction (layers/feature_interaction/dot_interaction.py:53-104): concat -> [B, F, d] (:75-77),
  xactions = feats @ feats^T (:83), lower triangle with / without the diagonal via boolean_mask in row-major order
  (:85-102), or the full matrix with the upper part zeroed when skip_gather (:95-100).  Every dot product is the
  canonical sequential float32 fmaf chain (dots())."""
  dims = {np.asarray(t).shape[1] for t in inputs}
  if len(dims) != 1:
    raise ValueError("Input tensors` dimensions must be equal")
  feats = np.stack([np.asarray(t, np.float32) for t in inputs], 1)  # [B, F, d]
  B, F, d = feats.shape
  x = np.zeros((B, F, F), np.float32)
  for b in range(B):
    x[b] = scores(feats[b], feats[b])
  i, j = np.meshgrid(np.arange(F), np.arange(F), indexing="ij")
  lower = (j <= i) if self_interaction else (j < i)
  if skip_gather:
    return np.where(lower[None], x, np.float32(0)).reshape(B, F * F)
  return x[:, lower]  # boolean mask walks (i, j) row-major


def dot_interaction_grads(inputs, gout, self_interaction: bool = False, skip_gather: bool = False):
  """d out / d feats [B, F, d] in float64."""
  feats = np.stack([np.asarray(t, np.float64) for t in inputs], 1)
  B, F, d = feats.shape
  i, j = np.meshgrid(np.arange(F), np.arange(F), indexing="ij")
  lower = (j <= i) if self_interaction else (j < i)
  G = np.zeros((B, F, F))
  g = np.asarray(gout, np.float64)
  if skip_gather:
    G = np.where(lower[None], g.reshape(B, F, F), 0.0)
  else:
    G[:, lower] = g
  return np.einsum("bij,bjd->bid", G + G.transpose(0, 2, 1), feats)


def cross_grads(x0, x, W, bias, dout, diag_scale: float = 0.0):
  """Backward of the full-rank Cross without preactivation (float64)."""
  x0_ = np.asarray(x0, np.float64); x_ = np.asarray(x, np.float64); W_ = np.asarray(W, np.float64)
  g = np.asarray(dout, np.float64)
  prod = x_ @ W_ + (0 if bias is None else np.asarray(bias, np.float64)) + diag_scale * x_
  dx0 = g * prod
  gp = g * x0_
  dx = gp @ W_.T + diag_scale * gp + g
  dW = x_.T @ gp
  db = gp.sum(0)
  return dx0, dx, dW, db


# ----------------------------------------------------------------------------------------------
# embedding gather + sparse Adagrad
# ----------------------------------------------------------------------------------------------
def gather(table, ids) -> np.ndarray:
  """tf.keras.layers.Embedding lookup -> tf.gather (README.md:62-66)."""
  t = _f32(table); i = np.ascontiguousarray(ids, np.int64).reshape(-1)
  out = np.empty((i.shape[0], t.shape[1]), np.float32)
  _lib().orc_gather(_p(t), ctypes.c_int64(t.shape[0]), ctypes.c_int(t.shape[1]), _p(i),
                    ctypes.c_int64(i.shape[0]), _p(out), ctypes.c_int64(t.shape[1]), ctypes.c_int64(0))
  return out


def sparse_adagrad(table, accum, ids, grad_rows, lr: float, eps: float = 1e-7, eps_inside_sqrt: bool = True):
  """In-place sparse Adagrad on copies; returns (table, accum).  See tfrs_oracle.c for the rule."""
  t = _f32(table).copy(); a = _f32(accum).copy()
  i = np.ascontiguousarray(ids, np.int64).reshape(-1); g = _f32(grad_rows)
  _lib().orc_sparse_adagrad(_p(t), _p(a), ctypes.c_int64(t.shape[0]), ctypes.c_int(t.shape[1]), _p(i),
                            ctypes.c_int64(i.shape[0]), _p(g), ctypes.c_float(lr), ctypes.c_float(eps),
                            ctypes.c_int(1 if eps_inside_sqrt else 0))
  return t, a


# ----------------------------------------------------------------------------------------------
# CPU arm for bench.py: the reference's op sequence with a BLAS sgemm (what TF-CPU dispatches to)
# ----------------------------------------------------------------------------------------------
def brute_force_blas(q: np.ndarray, c: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
  """matmul -> top_k -> (ids = indices), factorized_top_k.py:603-607, materialising [Q,N] like TF does."""
  s = q @ c.T
  part = np.argpartition(-s, k - 1, axis=1)[:, :k]
  ps = np.take_along_axis(s, part, 1)
  order = np.lexsort((part, -ps), axis=1)
  return np.take_along_axis(ps, order, 1), np.take_along_axis(part, order, 1)


def brute_force_torch(q: np.ndarray, c: np.ndarray, k: int, chunk: int = 512, threads: Optional[int] = None):
  """SURVEY 8d's CPU protocol for the reference's BruteForce.call (factorized_top_k.py:603-607): torch CPU sgemm
  (MKL/oneDNN, all cores) -> torch.topk(sorted=True) -> indices, queries chunked by 512 to bound the [Q,N] buffer (the
  reference's own tutorial chunks by 1000).  A SPEED baseline: torch.topk does not promise the lowest-index tie rule, so
  this is never used as the checker."""
  import torch
  if threads:
    torch.set_num_threads(int(threads))
  qt = torch.from_numpy(np.ascontiguousarray(q, np.float32)); ct = torch.from_numpy(np.ascontiguousarray(c, np.float32))
  vs, ix = [], []
  with torch.no_grad():
    for lo in range(0, qt.shape[0], chunk):
      s = qt[lo:lo + chunk] @ ct.T
      v, i = torch.topk(s, k, dim=1, sorted=True)
      vs.append(v); ix.append(i)
  return torch.cat(vs).numpy(), torch.cat(ix).numpy()
