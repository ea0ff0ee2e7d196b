"""Functional wrappers over the C ABI (one Python function per entry point of include/tfrs_b200.h).

Everything here takes/returns CUDA torch tensors; torch only provides memory, streams and the
autograd tape.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _ffi
from ._ffi import c_f, c_i, c_l, c_sz, check, f32c, lib, ptr, require_cuda, stream, workspace

# Corpora at least this large go through the tensor-core screening path when an index image exists.
TC_MIN_N = 16384
TC_MAX_K = 256
TC_MAX_Q_PER_CALL = 8192


# ------------------------------------------------------------------------------------------------
# K1 gather
# ------------------------------------------------------------------------------------------------
def gather(tables: Sequence[torch.Tensor], ids: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None,
           col_offsets: Optional[Sequence[int]] = None) -> torch.Tensor:
  """out[i, off_t:off_t+dim_t] = tables[t][ids[t][i]]  -- concatenated multi-table embedding lookup."""
  nt = len(tables)
  if nt == 0 or len(ids) != nt:
    raise ValueError("gather: need as many id tensors as tables")
  n = ids[0].numel()
  dims = [int(t.shape[1]) for t in tables]
  if col_offsets is None:
    col_offsets, o = [], 0
    for d in dims:
      col_offsets.append(o); o += d
    width = o
  else:
    width = max(o + d for o, d in zip(col_offsets, dims))
  dev = tables[0].device
  tabs = [f32c(t, "table") for t in tables]
  code = _ffi.ids_dtype_code(ids[0])
  idl = []
  for x in ids:
    require_cuda(x, "ids")
    if _ffi.ids_dtype_code(x) != code or x.numel() != n:
      raise ValueError("gather: all id tensors must share dtype and length")
    idl.append(x.contiguous().view(-1))
  if out is None:
    out = torch.empty((n, width), dtype=torch.float32, device=dev)
  else:
    require_cuda(out, "out")
    if out.dtype != torch.float32 or out.stride(-1) != 1 or out.shape[0] != n:
      raise ValueError("gather: out must be float32 [n, >=width] with unit inner stride")
  out_ld = out.stride(0) if out.dim() == 2 else width
  tp = (ctypes.c_void_p * nt)(*[t.data_ptr() for t in tabs])
  ip = (ctypes.c_void_p * nt)(*[x.data_ptr() for x in idl])
  rows = (ctypes.c_int64 * nt)(*[int(t.shape[0]) for t in tabs])
  dm = (ctypes.c_int32 * nt)(*dims)
  co = (ctypes.c_int32 * nt)(*[int(c) for c in col_offsets])
  check(lib().tfrs_gather_f32(tp, rows, dm, nt, ip, code, n, ptr(out), out_ld, co, stream()), "gather")
  return out


# ------------------------------------------------------------------------------------------------
# K2 top-k
# ------------------------------------------------------------------------------------------------
def topk_scan(q: torch.Tensor, corpus: torch.Tensor, k: int, index_offset: int = 0,
              state: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
  """Exact brute-force top-k (CUDA-core path) with optional carried state. Returns ([Q,k_out] f32, [Q,k_out] i64)."""
  q = f32c(q, "queries"); corpus = f32c(corpus, "candidates")
  if q.dim() != 2 or corpus.dim() != 2 or q.shape[1] != corpus.shape[1]:
    raise ValueError(f"topk_scan: shape mismatch {tuple(q.shape)} vs {tuple(corpus.shape)}")
  Q, d = q.shape; N = corpus.shape[0]
  st_s = st_i = None; st_k = 0
  if state is not None and state[0].shape[1] > 0:
    st_s = f32c(state[0], "state scores"); st_i = require_cuda(state[1], "state idx").to(torch.int64).contiguous()
    st_k = st_s.shape[1]
  k_out = min(k, st_k + N)
  if out is None:
    out_s = torch.empty((Q, k), dtype=torch.float32, device=q.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=q.device)
  else:
    out_s, out_i = out  # contiguous [Q, k] f32 / i64 views supplied by the caller
  if Q == 0 or k_out == 0:
    return out_s[:, :0], out_i[:, :0]
  wsb = lib().tfrs_topk_scan_workspace_bytes(Q, N, d, k)
  ws = workspace(wsb, q.device, "scan")
  check(lib().tfrs_topk_scan_f32(ptr(q), Q, ptr(corpus), N, d, k, index_offset, ptr(st_s), ptr(st_i), st_k,
                                 ptr(out_s), ptr(out_i), ptr(ws), ws.numel(), stream()), "topk_scan")
  return out_s[:, :k_out], out_i[:, :k_out]


def index_build(corpus: torch.Tensor, reuse_slot: Optional[str] = None) -> torch.Tensor:
  """Builds the tensor-core screening image (fp16 UMMA tiles + norm bound) of a corpus.  `reuse_slot` builds it in a
  per-stream scratch buffer instead of a fresh allocation (Streaming's per-chunk images)."""
  corpus = f32c(corpus, "candidates")
  N, d = corpus.shape
  nb = lib().tfrs_index_bytes(N, d)
  if nb == 0:
    raise NotImplementedError("tensor-core index not available for this shape")
  buf = torch.empty(nb, dtype=torch.uint8, device=corpus.device) if reuse_slot is None else workspace(nb, corpus.device, reuse_slot)
  check(lib().tfrs_index_build(ptr(corpus), N, d, ptr(buf), nb, stream()), "index_build")
  return buf


def topk_tc(q: torch.Tensor, corpus: torch.Tensor, index_buf: torch.Tensor, k: int, index_offset: int = 0,
            out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
  """tcgen05 screening + exact rescoring; bit-identical to topk_scan."""
  q = f32c(q, "queries"); corpus = f32c(corpus, "candidates")
  Q, d = q.shape; N = corpus.shape[0]
  if out is None:
    out_s = torch.empty((Q, k), dtype=torch.float32, device=q.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=q.device)
  else:
    out_s, out_i = out
  if Q == 0:
    return out_s, out_i
  if Q > TC_MAX_Q_PER_CALL:  # bound the workspace (bin maxima + survivor records scale with Q): query chunks
    for lo in range(0, Q, TC_MAX_Q_PER_CALL):
      hi = min(Q, lo + TC_MAX_Q_PER_CALL)
      topk_tc(q[lo:hi], corpus, index_buf, k, index_offset, out=(out_s[lo:hi], out_i[lo:hi]))
    return out_s, out_i
  wsb = lib().tfrs_topk_tc_workspace_bytes(Q, N, d, k)
  ws = workspace(wsb, q.device, "tc")
  check(lib().tfrs_topk_tc_f32(ptr(q), Q, ptr(corpus), ptr(index_buf), N, d, k, index_offset, ptr(out_s),
                               ptr(out_i), ptr(ws), ws.numel(), stream()), "topk_tc")
  return out_s, out_i


def _i64(t, name: str, device) -> torch.Tensor:
  if not isinstance(t, torch.Tensor):
    t = torch.as_tensor(t)
  return t.to(device=device, dtype=torch.int64).contiguous()


def topk_tc_exclude(q: torch.Tensor, corpus: torch.Tensor, index_buf: torch.Tensor, k: int, exclusions: torch.Tensor,
                    identifiers: Optional[torch.Tensor] = None, index_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
  """`query_with_exclusions` fused into the tensor-core scan's finalize step: ([Q,k] f32 original scores,
  [Q,k] i64 global indices).  `identifiers` (integer tensor covering the corpus, or None = the row index) and
  `exclusions` [Q,E] are compared as int64."""
  q = f32c(q, "queries"); corpus = f32c(corpus, "candidates")
  Q, d = q.shape; N = corpus.shape[0]
  ex = _i64(exclusions, "exclusions", q.device)
  E = int(ex.shape[1])
  ids = None if identifiers is None else _i64(identifiers, "identifiers", q.device)
  out_s = torch.empty((Q, k), dtype=torch.float32, device=q.device)
  out_i = torch.empty((Q, k), dtype=torch.int64, device=q.device)
  if Q == 0:
    return out_s, out_i
  for lo in range(0, Q, TC_MAX_Q_PER_CALL):
    hi = min(Q, lo + TC_MAX_Q_PER_CALL)
    wsb = lib().tfrs_topk_tc_workspace_bytes(hi - lo, N, d, k + E)
    ws = workspace(wsb, q.device, "tc")
    check(lib().tfrs_topk_tc_exclude_f32(ptr(q[lo:hi]), hi - lo, ptr(corpus), ptr(index_buf), N, d, k, index_offset, ptr(ids),
                                         ptr(ex[lo:hi]), E, ptr(out_s[lo:hi]), ptr(out_i[lo:hi]), ptr(ws), ws.numel(), stream()),
          "topk_tc_exclude")
  return out_s, out_i


def topk_tc_count(q: torch.Tensor, corpus: torch.Tensor, index_buf: torch.Tensor, k: int, positive_scores: torch.Tensor
                  ) -> torch.Tensor:
  """min(k, #{candidates scoring strictly above the positive}) per query, int32 [Q] -- the fused score branch of
  FactorizedTopK (no top-K list is produced)."""
  q = f32c(q, "queries"); corpus = f32c(corpus, "candidates")
  pos = f32c(positive_scores, "positive_scores").view(-1)
  Q, d = q.shape; N = corpus.shape[0]
  out = torch.empty((Q,), dtype=torch.int32, device=q.device)
  for lo in range(0, Q, TC_MAX_Q_PER_CALL):
    hi = min(Q, lo + TC_MAX_Q_PER_CALL)
    wsb = lib().tfrs_topk_tc_workspace_bytes(hi - lo, N, d, k)
    ws = workspace(wsb, q.device, "tc")
    check(lib().tfrs_topk_tc_count_f32(ptr(q[lo:hi]), hi - lo, ptr(corpus), ptr(index_buf), N, d, k, ptr(pos[lo:hi]),
                                       ptr(out[lo:hi]), ptr(ws), ws.numel(), stream()), "topk_tc_count")
  return out


def exclude_rerank(scores: torch.Tensor, idx: torch.Tensor, exclusions: torch.Tensor, k: int,
                   identifiers: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
  """`_exclude` (factorized_top_k.py:83-115) on a fetched [Q, kf] list of (score, global index): returns the original
  scores / indices of the min(k, kf) best after lowering excluded identifiers by 1e5."""
  scores = f32c(scores, "scores"); idx = _i64(idx, "idx", scores.device)
  Q, kf = scores.shape
  ex = _i64(exclusions, "exclusions", scores.device)
  ids = None if identifiers is None else _i64(identifiers, "identifiers", scores.device)
  k_out = min(k, kf)
  out_s = torch.empty((Q, k_out), dtype=torch.float32, device=scores.device)
  out_i = torch.empty((Q, k_out), dtype=torch.int64, device=scores.device)
  if Q and k_out:
    check(lib().tfrs_topk_exclude_rerank_f32(ptr(scores), ptr(idx), Q, kf, ptr(ids), ptr(ex), int(ex.shape[1]), k_out,
                                             ptr(out_s), ptr(out_i), stream()), "exclude_rerank")
  return out_s, out_i


def count_above(scores: torch.Tensor, positive_scores: torch.Tensor) -> torch.Tensor:
  """#{t : scores[q, t] > positive[q]}, int32 [Q] (tf.math.in_top_k's count on a retrieved list)."""
  scores = f32c(scores, "scores"); pos = f32c(positive_scores, "positive_scores").view(-1)
  Q, k = scores.shape
  out = torch.empty((Q,), dtype=torch.int32, device=scores.device)
  check(lib().tfrs_count_above_f32(ptr(scores), scores.stride(0), k, ptr(pos), Q, ptr(out), stream()), "count_above")
  return out


def hits_accumulate(count: torch.Tensor, positive_scores: torch.Tensor, sample_weight: Optional[torch.Tensor],
                    ks: Sequence[int], acc: torch.Tensor) -> None:
  """acc[j] += sum_q w_q [count_q < ks[j], positive finite]; acc[len(ks)] += sum_q w_q.  acc: float64 [len(ks)+1] on
  the device; nothing is synchronised."""
  Q = count.numel()
  w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
  if w is not None and w.numel() != Q:
    raise ValueError(f"sample_weight must have one entry per query (got {w.numel()}, expected {Q})")
  karr = (ctypes.c_int32 * len(ks))(*[int(x) for x in ks])
  check(lib().tfrs_topk_hits_accumulate(ptr(count), ptr(f32c(positive_scores, "positive_scores").view(-1)), ptr(w), Q, karr,
                                        len(ks), ptr(acc), stream()), "hits_accumulate")


def topk_merge_sorted(scores: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
  """Merge [L,Q,k_in] lists that are each sorted (score desc, index asc): rank by binary search, no sort."""
  scores = f32c(scores, "scores"); idx = require_cuda(idx, "idx").to(torch.int64).contiguous()
  L, Q, k_in = scores.shape
  k_out = min(k, L * k_in)
  out_s = torch.empty((Q, k_out), dtype=torch.float32, device=scores.device)
  out_i = torch.empty((Q, k_out), dtype=torch.int64, device=scores.device)
  check(lib().tfrs_topk_merge_sorted_strided(ptr(scores), ptr(idx), Q * k_in, Q * k_in, L, Q, k_in, k_out, ptr(out_s), ptr(out_i),
                                             stream()), "topk_merge_sorted")
  return out_s, out_i


def topk_sharded(comm, q: torch.Tensor, corpus_local: torch.Tensor, index_buf: Optional[torch.Tensor], k: int,
                 index_offset: int) -> Tuple[torch.Tensor, torch.Tensor]:
  """The row-sharded BruteForce call through the C ABI: local scan -> one NCCL all-gather -> merge, on every rank."""
  q = f32c(q, "queries"); corpus_local = f32c(corpus_local, "candidates")
  Q, d = q.shape; N = corpus_local.shape[0]
  out_s = torch.empty((Q, k), dtype=torch.float32, device=q.device)
  out_i = torch.empty((Q, k), dtype=torch.int64, device=q.device)
  comm.ensure_p2p(Q, k)
  wsb = lib().tfrs_topk_sharded_workspace_bytes(comm.world, Q, N, d, k)
  ws = workspace(wsb + 1024, q.device, "sharded")
  check(lib().tfrs_topk_sharded_f32(comm.handle, ptr(q), Q, ptr(corpus_local), ptr(index_buf), N, d, k, index_offset,
                                    ptr(out_s), ptr(out_i), ptr(ws), ws.numel(), stream()), "topk_sharded")
  return out_s, out_i


def topk_merge_packed(gathered: torch.Tensor, n_lists: int, Q: int, k_in: int, k: int, idx_byte_offset: int,
                      block_bytes: int, sorted_lists: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """Merge the receive buffer of the sharded scan's single all-gather: `n_lists` blocks of `block_bytes`, each
  [scores f32 [Q,k_in] | pad | indices i64 [Q,k_in] at idx_byte_offset].  `sorted_lists` (what the scans emit:
  every list in (score desc, index asc) order) selects the rank-by-binary-search merge instead of the sort."""
  k_out = min(k, n_lists * k_in)
  out_s = torch.empty((Q, k_out), dtype=torch.float32, device=gathered.device)
  out_i = torch.empty((Q, k_out), dtype=torch.int64, device=gathered.device)
  base = gathered.data_ptr()
  fn = lib().tfrs_topk_merge_sorted_strided if sorted_lists else lib().tfrs_topk_merge_strided
  check(fn(ctypes.c_void_p(base), ctypes.c_void_p(base + idx_byte_offset), block_bytes // 4, block_bytes // 8, n_lists, Q,
           k_in, k_out, ptr(out_s), ptr(out_i), stream()), "topk_merge_strided")
  return out_s, out_i


def tc_supported(Q: int, N: int, d: int, k: int) -> bool:
  """True when (Q, N, d, k) is inside the tensor-core path (otherwise callers use topk_scan)."""
  return lib().tfrs_topk_tc_workspace_bytes(Q, N, d, k) > 0


def tc_last_call_stats(Q: int, N: int, d: int, k: int, device=None) -> dict:
  """Survivor / fallback statistics of the most recent topk_tc call with this shape (reads the cached
  workspace; synchronises).  Used by tests to prove the tensor-core path -- not the exact fallback --
  produced the result."""
  out = (ctypes.c_int64 * 8)()
  check(lib().tfrs_topk_tc_layout(Q, N, d, k, out), "topk_tc_layout")
  o_count, o_ovf, o_thr, o_cand, parts, cap, Qp, o_cut = [int(x) for x in out]
  dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
  ws = workspace(0, dev, "tc")
  base = (-ws.data_ptr()) % 16
  torch.cuda.synchronize()
  counts = ws[base + o_count: base + o_count + Qp * parts * 4].view(torch.int32).view(Qp, parts)[:Q]
  ovf = ws[base + o_ovf: base + o_ovf + Q * 4].view(torch.int32)
  per_query = counts.sum(1)
  return {"fallback_queries": int((ovf != 0).sum()), "survivors_mean": float(per_query.float().mean()),
          "survivors_max": int(per_query.max()), "parts": parts, "cap_part": cap,
          "part_max": int(counts.max())}


def profile_enable(on: bool) -> None:
  check(lib().tfrs_profile_enable(int(on)), "profile_enable")


def profile_read():
  """-> (stage_ms[4], calls): prep, sample+threshold, filter, finalize; synchronises the device."""
  ms = (ctypes.c_float * 4)(); calls = ctypes.c_int(0)
  check(lib().tfrs_profile_read(ms, ctypes.byref(calls)), "profile_read")
  return [float(x) for x in ms], int(calls.value)


def topk_merge(scores: torch.Tensor, idx: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
  """Merge [L,Q,k_in] lists into the best min(k, L*k_in) per query."""
  scores = f32c(scores, "scores"); idx = require_cuda(idx, "idx").to(torch.int64).contiguous()
  L, Q, k_in = scores.shape
  k_out = min(k, L * k_in)
  out_s = torch.empty((Q, k_out), dtype=torch.float32, device=scores.device)
  out_i = torch.empty((Q, k_out), dtype=torch.int64, device=scores.device)
  check(lib().tfrs_topk_merge(ptr(scores), ptr(idx), L, Q, k_in, k_out, ptr(out_s), ptr(out_i), stream()), "topk_merge")
  return out_s, out_i


# ------------------------------------------------------------------------------------------------
# exact matmul / scores (autograd-aware)
# ------------------------------------------------------------------------------------------------
def sgemm(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False,
          out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
  a = f32c(a, "a"); b = f32c(b, "b")
  M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
  Kb, N = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
  if K != Kb:
    raise ValueError(f"sgemm: inner dimensions differ ({K} vs {Kb})")
  if out is None:
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
  check(lib().tfrs_sgemm_f32(int(trans_a), int(trans_b), M, N, K, ptr(a), a.stride(0), ptr(b), b.stride(0),
                             ptr(out), out.stride(0), int(accumulate), stream()), "sgemm")
  return out


class _Scores(torch.autograd.Function):
  """`_compute_score` = matmul(q, c^T) (layers/factorized_top_k.py:320-333) with exact backward."""

  @staticmethod
  def forward(ctx, q, c):
    q = f32c(q, "queries"); c = f32c(c, "candidates")
    ctx.save_for_backward(q, c)
    return sgemm(q, c, False, True)

  @staticmethod
  def backward(ctx, g):
    q, c = ctx.saved_tensors
    g = f32c(g, "grad")
    dq = sgemm(g, c, False, False) if ctx.needs_input_grad[0] else None
    dc = sgemm(g, q, True, False) if ctx.needs_input_grad[1] else None
    return dq, dc


def scores(q: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
  return _Scores.apply(q, c)


class _Matmul(torch.autograd.Function):
  """x @ w with our exact SGEMM (used by the low-rank Cross path)."""

  @staticmethod
  def forward(ctx, x, w):
    x = f32c(x, "x"); w = f32c(w, "w")
    ctx.save_for_backward(x, w)
    return sgemm(x, w, False, False)

  @staticmethod
  def backward(ctx, g):
    x, w = ctx.saved_tensors
    g = f32c(g, "grad")
    dx = sgemm(g, w, False, True) if ctx.needs_input_grad[0] else None
    dw = sgemm(x, g, True, False) if ctx.needs_input_grad[1] else None
    return dx, dw


def matmul(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
  return _Matmul.apply(x, w)


def rowwise_dot(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
  a = f32c(a, "a"); b = f32c(b, "b")
  out = torch.empty((a.shape[0],), dtype=torch.float32, device=a.device)
  check(lib().tfrs_rowwise_dot_f32(ptr(a), ptr(b), a.shape[0], a.shape[1], ptr(out), stream()), "rowwise_dot")
  return out


# ------------------------------------------------------------------------------------------------
# K3 in-batch softmax loss
# ------------------------------------------------------------------------------------------------
SOFTMAX_TC_MIN_B = 512  # below this the exact CUDA-core forward is launch-latency bound anyway


def inbatch_softmax_tc(q: torch.Tensor, c: torch.Tensor, sample_weight: Optional[torch.Tensor] = None,
                       inv_temperature: float = 1.0, candidate_bias: Optional[torch.Tensor] = None):
  """Tensor-core forward only (any B): returns (loss scalar, lse [B]).  Raises NotImplementedError when d > 128.
  `candidate_bias` [C] is added to every logit of its column (after the temperature)."""
  q = f32c(q, "query_embeddings"); c = f32c(c, "candidate_embeddings")
  B, d = q.shape; C = c.shape[0]
  w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
  loss = torch.empty((1,), dtype=torch.float32, device=q.device)
  lse = torch.empty((B,), dtype=torch.float32, device=q.device)
  cb = None if candidate_bias is None else f32c(candidate_bias, "candidate_bias").view(-1)
  ws = workspace(max(lib().tfrs_inbatch_softmax_tc_workspace_bytes(B, C, d), 256), q.device, "softmax_tc")
  check(lib().tfrs_inbatch_softmax_tc_fwd(ptr(q), ptr(c), B, C, d, c_f(inv_temperature), ptr(w), ptr(cb), ptr(loss), ptr(lse),
                                          ptr(ws), ws.numel(), stream()), "inbatch_softmax_tc_fwd")
  return loss.view(()), lse


def _ids_i64(candidate_ids, C: int, device) -> torch.Tensor:
  """candidate ids of any type -> int64 [C] on the device.  Integer tensors are used as they are; anything else (strings,
  NumPy object arrays, float ids) is factorised on the host -- only EQUALITY of ids matters to accidental-hit removal."""
  if isinstance(candidate_ids, torch.Tensor) and not candidate_ids.dtype.is_floating_point and candidate_ids.dtype != torch.bool:
    t = candidate_ids.reshape(-1).to(device=device, dtype=torch.int64).contiguous()
  else:
    import numpy as np
    arr = candidate_ids.detach().cpu().numpy() if isinstance(candidate_ids, torch.Tensor) else np.asarray(candidate_ids)
    _, inv = np.unique(arr.reshape(-1), return_inverse=True)
    t = torch.as_tensor(inv.astype(np.int64), device=device)
  if t.numel() != C:
    raise ValueError(f"candidate_ids must have one entry per candidate (got {t.numel()}, expected {C})")
  return t


class _InBatchSoftmax(torch.autograd.Function):

  @staticmethod
  def forward(ctx, q, c, sample_weight, inv_temperature, candidate_bias=None, candidate_ids=None, score_mask=None):
    q = f32c(q, "query_embeddings"); c = f32c(c, "candidate_embeddings")
    B, d = q.shape; C = c.shape[0]
    w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
    cb = None if candidate_bias is None else f32c(candidate_bias, "candidate_bias").view(-1)
    if cb is not None and cb.numel() != C:
      raise ValueError(f"candidate_bias must have one entry per candidate (got {cb.numel()}, expected {C})")
    if w is not None and w.numel() != B:
      raise ValueError(f"sample_weight must have one entry per query (got {w.numel()}, expected {B})")
    ids = None if candidate_ids is None else _ids_i64(candidate_ids, C, q.device)
    mask = None
    if score_mask is not None:
      mask = require_cuda(score_mask, "score_mask")
      if tuple(mask.shape) != (B, C):
        raise ValueError(f"score_mask must be [{B},{C}], got {tuple(mask.shape)}")
      mask = (mask if mask.dtype == torch.bool else mask != 0).contiguous().view(torch.uint8)
    ext = ids is not None or mask is not None
    if (cb is not None or ext) and not inbatch_softmax_bias_supported(B, C, d):
      raise NotImplementedError("inbatch_softmax_loss: candidate_bias / candidate_ids / score_mask need the tensor-core path "
                                f"(B >= {SOFTMAX_TC_MIN_B}, d <= 64); got B={B}, d={d}")
    loss = torch.empty((1,), dtype=torch.float32, device=q.device)
    lse = torch.empty((B,), dtype=torch.float32, device=q.device)
    tcb = lib().tfrs_inbatch_softmax_tc_ex_workspace_bytes(B, C, d, int(ids is not None), int(mask is not None)) \
        if B >= SOFTMAX_TC_MIN_B else 0
    if tcb:  # tensor-core forward (hi/lo fp16 split, fp32 accumulate, online log-sum-exp epilogue)
      ws = workspace(tcb, q.device, "softmax_tc")
      check(lib().tfrs_inbatch_softmax_tc_fwd_ex(ptr(q), ptr(c), B, C, d, c_f(inv_temperature), ptr(w), ptr(cb), ptr(ids), ptr(mask),
                                                 ptr(loss), ptr(lse), ptr(ws), ws.numel(), stream()), "inbatch_softmax_tc_fwd")
    else:
      wsb = lib().tfrs_inbatch_softmax_workspace_bytes(B, C, d)
      ws = workspace(wsb, q.device, "softmax")
      check(lib().tfrs_inbatch_softmax_fwd(ptr(q), ptr(c), B, C, d, c_f(inv_temperature), ptr(w), ptr(loss), ptr(lse),
                                           ptr(ws), ws.numel(), stream()), "inbatch_softmax_fwd")
    empty = torch.empty(0, device=q.device)
    ctx.save_for_backward(q, c, lse, w if w is not None else empty, cb if cb is not None else empty,
                          ids if ids is not None else empty, mask if mask is not None else empty)
    ctx.has_w = w is not None
    ctx.has_cb = cb is not None
    ctx.has_ids = ids is not None
    ctx.has_mask = mask is not None
    ctx.inv_t = inv_temperature
    ctx.used_tc = bool(tcb)
    return loss.view(())

  @staticmethod
  def backward(ctx, g):
    q, c, lse, w, cb, ids, mask = ctx.saved_tensors
    B, d = q.shape; C = c.shape[0]
    g = f32c(g, "grad").view(1)
    dq = torch.empty_like(q); dc = torch.empty_like(c)
    tcb = lib().tfrs_inbatch_softmax_tc_bwd_ex_workspace_bytes(B, C, d, int(ctx.has_ids), int(ctx.has_mask)) if ctx.used_tc else 0
    if tcb:  # tensor-core backward: same split products as the forward pass that produced `lse`
      ws = workspace(tcb, q.device, "softmax_tc_bwd")
      check(lib().tfrs_inbatch_softmax_tc_bwd_ex(ptr(q), ptr(c), B, C, d, c_f(ctx.inv_t), ptr(w) if ctx.has_w else None,
                                                 ptr(cb) if ctx.has_cb else None, ptr(ids) if ctx.has_ids else None,
                                                 ptr(mask) if ctx.has_mask else None, ptr(lse), ptr(g), ptr(dq), ptr(dc), ptr(ws),
                                                 ws.numel(), stream()), "inbatch_softmax_tc_bwd")
      return dq, dc, None, None, None, None, None
    if ctx.has_cb or ctx.has_ids or ctx.has_mask:
      raise NotImplementedError("inbatch_softmax_loss backward with loss options needs the tensor-core path")
    wsb = lib().tfrs_inbatch_softmax_workspace_bytes(B, C, d)
    ws = workspace(wsb, q.device, "softmax")
    check(lib().tfrs_inbatch_softmax_bwd(ptr(q), ptr(c), B, C, d, c_f(ctx.inv_t), ptr(w) if ctx.has_w else None,
                                         ptr(lse), ptr(g), ptr(dq), ptr(dc), ptr(ws), ws.numel(), stream()),
          "inbatch_softmax_bwd")
    return dq, dc, None, None, None, None, None


def inbatch_softmax_bwd_exact(q, c, lse, sample_weight=None, inv_temperature: float = 1.0, grad_loss=None):
  """Exact fp32 CUDA-core backward for a given saved `lse` (the anchor the tensor-core backward is tested against)."""
  q = f32c(q, "query_embeddings"); c = f32c(c, "candidate_embeddings"); lse = f32c(lse, "lse")
  B, d = q.shape; C = c.shape[0]
  w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
  g = None if grad_loss is None else f32c(grad_loss, "grad").view(1)
  dq = torch.empty_like(q); dc = torch.empty_like(c)
  ws = workspace(lib().tfrs_inbatch_softmax_workspace_bytes(B, C, d), q.device, "softmax")
  check(lib().tfrs_inbatch_softmax_bwd(ptr(q), ptr(c), B, C, d, c_f(inv_temperature), ptr(w), ptr(lse), ptr(g), ptr(dq), ptr(dc),
                                       ptr(ws), ws.numel(), stream()), "inbatch_softmax_bwd")
  return dq, dc


def inbatch_softmax_tc_bwd(q, c, lse, sample_weight=None, inv_temperature: float = 1.0, grad_loss=None, candidate_bias=None):
  """Tensor-core backward only (any B, d <= 64): returns (dq, dc) for the given saved `lse`."""
  q = f32c(q, "query_embeddings"); c = f32c(c, "candidate_embeddings"); lse = f32c(lse, "lse")
  B, d = q.shape; C = c.shape[0]
  w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
  g = None if grad_loss is None else f32c(grad_loss, "grad").view(1)
  dq = torch.empty_like(q); dc = torch.empty_like(c)
  ws = workspace(max(lib().tfrs_inbatch_softmax_tc_bwd_workspace_bytes(B, C, d), 256), q.device, "softmax_tc_bwd")
  cb = None if candidate_bias is None else f32c(candidate_bias, "candidate_bias").view(-1)
  check(lib().tfrs_inbatch_softmax_tc_bwd(ptr(q), ptr(c), B, C, d, c_f(inv_temperature), ptr(w), ptr(cb), ptr(lse), ptr(g),
                                          ptr(dq), ptr(dc), ptr(ws), ws.numel(), stream()), "inbatch_softmax_tc_bwd")
  return dq, dc


def inbatch_softmax_bias_supported(B: int, C: int, d: int) -> bool:
  """True when the loss with a per-candidate logit bias can run fused (tensor-core forward AND backward)."""
  return (B >= SOFTMAX_TC_MIN_B and lib().tfrs_inbatch_softmax_tc_workspace_bytes(B, C, d) > 0 and
          lib().tfrs_inbatch_softmax_tc_bwd_workspace_bytes(B, C, d) > 0)


def inbatch_softmax_loss(q: torch.Tensor, c: torch.Tensor, sample_weight: Optional[torch.Tensor] = None,
                         temperature: Optional[float] = None, candidate_bias: Optional[torch.Tensor] = None,
                         candidate_ids=None, score_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
  """sum_i w_i * (logsumexp_j(l_ij) - l_ii),  l_ij = q_i.c_j / T + b_j  -- tasks/retrieval.py:178-210;
  b = candidate_bias (e.g. -log(clip(p, 1e-6, 1)): the sampling-probability correction, :190-192), no gradient;
  candidate_ids: accidental-hit removal (:194-200) -- l_ij = MIN_FLOAT where id_j == id_i, j != i;
  score_mask [B,C]: l_ij = MIN_FLOAT where the mask is False (:202-203)."""
  inv_t = 1.0 if temperature is None else 1.0 / float(temperature)
  return _InBatchSoftmax.apply(q, c, sample_weight, inv_t, candidate_bias, candidate_ids, score_mask)


class _MaxSimSoftmax(torch.autograd.Function):
  """In-batch softmax loss on multi-head queries [B,H,d]: scores = max over heads (tasks/retrieval.py:172-176)."""

  @staticmethod
  def forward(ctx, q, c, sample_weight, inv_temperature):
    q = f32c(q, "query_embeddings"); c = f32c(c, "candidate_embeddings")
    B, H, d = q.shape; C = c.shape[0]
    w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
    if w is not None and w.numel() != B:
      raise ValueError(f"sample_weight must have one entry per query (got {w.numel()}, expected {B})")
    loss = torch.empty((1,), dtype=torch.float32, device=q.device)
    lse = torch.empty((B,), dtype=torch.float32, device=q.device)
    ws = workspace(lib().tfrs_inbatch_softmax_maxsim_workspace_bytes(B, H, C, d), q.device, "softmax")
    check(lib().tfrs_inbatch_softmax_maxsim_fwd(ptr(q), ptr(c), B, H, C, d, c_f(inv_temperature), ptr(w), ptr(loss), ptr(lse),
                                                ptr(ws), ws.numel(), stream()), "inbatch_softmax_maxsim_fwd")
    ctx.save_for_backward(q, c, lse, w if w is not None else torch.empty(0, device=q.device))
    ctx.has_w = w is not None
    ctx.inv_t = inv_temperature
    return loss.view(())

  @staticmethod
  def backward(ctx, g):
    q, c, lse, w = ctx.saved_tensors
    B, H, d = q.shape; C = c.shape[0]
    g = f32c(g, "grad").view(1)
    dq = torch.empty_like(q); dc = torch.empty_like(c)
    ws = workspace(lib().tfrs_inbatch_softmax_maxsim_workspace_bytes(B, H, C, d), q.device, "softmax")
    check(lib().tfrs_inbatch_softmax_maxsim_bwd(ptr(q), ptr(c), B, H, C, d, c_f(ctx.inv_t), ptr(w) if ctx.has_w else None, ptr(lse),
                                                ptr(g), ptr(dq), ptr(dc), ptr(ws), ws.numel(), stream()), "inbatch_softmax_maxsim_bwd")
    return dq, dc, None, None


def inbatch_softmax_maxsim_loss(q: torch.Tensor, c: torch.Tensor, sample_weight: Optional[torch.Tensor] = None,
                                temperature: Optional[float] = None) -> torch.Tensor:
  """sum_i w_i (logsumexp_j(max_h q_ih.c_j / T) - max_h q_ih.c_i / T) for q [B,H,d] (tasks/retrieval.py:172-210)."""
  inv_t = 1.0 if temperature is None else 1.0 / float(temperature)
  return _MaxSimSoftmax.apply(q, c, sample_weight, inv_t)


# ------------------------------------------------------------------------------------------------
# hard-negative mining loss (top-K scan + sparse softmax)
# ------------------------------------------------------------------------------------------------
def hard_negative_supported(B: int, C: int, d: int, num_hard_negatives: int) -> bool:
  """True when Retrieval(num_hard_negatives=n) can run on the top-K scan instead of the [B,C] logits."""
  return num_hard_negatives >= 1 and C >= B and min(num_hard_negatives + 1, C) <= 2048


class _HardNegativeSoftmax(torch.autograd.Function):

  @staticmethod
  def forward(ctx, q, c, num_hard_negatives, sample_weight, inv_temperature):
    q = f32c(q, "query_embeddings"); c = f32c(c, "candidate_embeddings")
    B, d = q.shape; C = c.shape[0]
    if not inv_temperature > 0:
      raise NotImplementedError("hard_negative_softmax_loss needs a positive temperature")
    k1 = min(int(num_hard_negatives) + 1, C)
    w = None if sample_weight is None else f32c(sample_weight, "sample_weight").view(-1)
    if w is not None and w.numel() != B:
      raise ValueError(f"sample_weight must have one entry per query (got {w.numel()}, expected {B})")
    qd, cd = q.detach(), c.detach()
    if C >= TC_MIN_N and d <= 128 and tc_supported(B, C, d, k1):
      top_s, top_i = topk_tc(qd, cd, index_build(cd, reuse_slot="hardneg_index"), k1)
    else:
      top_s, top_i = topk_scan(qd, cd, k1)
    pos = rowwise_dot(qd, cd[:B])
    loss = torch.empty((1,), dtype=torch.float32, device=q.device)
    coef = torch.empty((B, k1 + 2), dtype=torch.float32, device=q.device)
    check(lib().tfrs_hardneg_loss_fwd(ptr(top_s), ptr(top_i), B, k1, ptr(pos), c_f(inv_temperature), ptr(w), ptr(loss), ptr(coef),
                                      stream()), "hardneg_loss_fwd")
    ctx.save_for_backward(q, c, top_i, coef)
    ctx.k1 = k1
    return loss.view(())

  @staticmethod
  def backward(ctx, g):
    q, c, top_i, coef = ctx.saved_tensors
    B, d = q.shape; C = c.shape[0]
    g = f32c(g, "grad").view(1)
    dq = torch.empty_like(q); dc = torch.empty_like(c)
    check(lib().tfrs_hardneg_loss_bwd(ptr(q), ptr(c), B, C, d, ptr(top_i), ctx.k1, ptr(coef), ptr(g), ptr(dq), ptr(dc), stream()),
          "hardneg_loss_bwd")
    return dq, dc, None, None, None


def hard_negative_softmax_loss(q: torch.Tensor, c: torch.Tensor, num_hard_negatives: int,
                               sample_weight: Optional[torch.Tensor] = None, temperature: Optional[float] = None) -> torch.Tensor:
  """Retrieval loss with HardNegativeMining(n) (tasks/retrieval.py:205-210, layers/loss.py:61-111): softmax cross-entropy
  over the positive and the n highest-scoring other candidates of each query."""
  inv_t = 1.0 if temperature is None else 1.0 / float(temperature)
  return _HardNegativeSoftmax.apply(q, c, int(num_hard_negatives), sample_weight, inv_t)


# ------------------------------------------------------------------------------------------------
# K4 sparse Adagrad
# ------------------------------------------------------------------------------------------------
def sparse_adagrad_(table: torch.Tensor, accum: torch.Tensor, ids: torch.Tensor, grad_rows: torch.Tensor,
                    lr: float, eps: float = 1e-7, eps_inside_sqrt: bool = True) -> None:
  require_cuda(table, "table"); require_cuda(accum, "accum")
  if table.dtype != torch.float32 or not table.is_contiguous() or accum.dtype != torch.float32 or not accum.is_contiguous():
    raise ValueError("sparse_adagrad_: table/accum must be contiguous float32")
  ids = require_cuda(ids, "ids").contiguous().view(-1)
  g = f32c(grad_rows, "grad_rows")
  n = ids.numel(); d = table.shape[1]
  if g.shape != (n, d):
    raise ValueError(f"sparse_adagrad_: grad_rows must be [{n},{d}], got {tuple(g.shape)}")
  wsb = lib().tfrs_sparse_adagrad_workspace_bytes(n, d)
  ws = workspace(wsb, table.device, "adagrad")
  check(lib().tfrs_sparse_adagrad_f32(ptr(table), ptr(accum), table.shape[0], d, ptr(ids), _ffi.ids_dtype_code(ids), n,
                                      ptr(g), c_f(lr), c_f(eps), int(eps_inside_sqrt), ptr(ws), ws.numel(), stream()),
        "sparse_adagrad")


# ------------------------------------------------------------------------------------------------
# K5 cross
# ------------------------------------------------------------------------------------------------
# Cross layers at least this large run their forward GEMM on the tensor cores (fp16 hi/lo split, fp32 accumulate).
CROSS_TC_MIN_B = 1024
CROSS_TC_MIN_D = 64
def cross_weight_image(W: torch.Tensor) -> torch.Tensor:
  """K-major fp16 hi/lo image of W^T for the tensor-core Cross kernel, rebuilt on every call: the split is
  O(D^2) against the O(B*D^2) GEMM (2.9 MB at D=845), and a cache keyed on (data_ptr, _version) can serve a stale
  image after the allocator recycles an address or after an in-place `.data` update."""
  D = W.shape[0]
  nb = lib().tfrs_cross_tc_weight_bytes(D)
  buf = workspace(nb, W.device, f"cross_w{D}")
  check(lib().tfrs_cross_tc_weight_build(ptr(W), D, ptr(buf), nb, stream()), "cross_tc_weight_build")
  return buf


class _Cross(torch.autograd.Function):

  @staticmethod
  def forward(ctx, x0, x, W, bias, diag_scale, x_amax):
    x0 = f32c(x0, "x0"); x = f32c(x, "x"); W = f32c(W, "kernel")
    b = None if bias is None else f32c(bias, "bias")
    B, D = x0.shape
    out = torch.empty_like(x0)
    need_grad = any(ctx.needs_input_grad[:4])
    prod = torch.empty_like(x0) if need_grad else None
    ctx.used_tc = B >= CROSS_TC_MIN_B and D >= CROSS_TC_MIN_D and W.shape == (D, D)
    # max |out| as float bits, produced by the tensor-core epilogue: the next layer of a stack uses it as its rescale
    # statistic instead of a pass over its input (0 on the exact path: "unknown")
    out_amax = torch.zeros((1,), dtype=torch.int32, device=x0.device)
    if ctx.used_tc:
      wimg = cross_weight_image(W)
      wsb = lib().tfrs_cross_tc_workspace_bytes(B, D)
      ws = workspace(wsb, x0.device, "cross_tc")
      check(lib().tfrs_cross_tc_fwd_ex_f32(ptr(x0), ptr(x), ptr(wimg), ptr(b), B, D, D, c_f(diag_scale), ptr(out), ptr(prod),
                                           ptr(x_amax), ptr(out_amax), ptr(ws), ws.numel(), stream()), "cross_tc_fwd")
    else:
      check(lib().tfrs_cross_fwd_f32(ptr(x0), ptr(x), ptr(W), ptr(b), B, D, D, c_f(diag_scale), ptr(out), ptr(prod),
                                     stream()), "cross_fwd")
    if need_grad:
      ctx.save_for_backward(x0, x, W, prod)
    ctx.diag = diag_scale
    ctx.has_bias = b is not None
    ctx.mark_non_differentiable(out_amax)
    return out, out_amax

  @staticmethod
  def backward(ctx, g, _g_amax=None):
    x0, x, W, prod = ctx.saved_tensors
    g = f32c(g, "grad")
    B, D = x0.shape
    n0, n1, n2, n3 = ctx.needs_input_grad[:4]
    dx0 = torch.empty_like(x0) if n0 else None
    dx = torch.empty_like(x) if n1 else None
    dW = torch.empty_like(W) if n2 else None
    db = torch.empty((D,), dtype=torch.float32, device=x0.device) if (n3 and ctx.has_bias) else None
    if ctx.used_tc:  # the two GEMMs (dx, dW) on the tensor cores
      ws = workspace(lib().tfrs_cross_tc_bwd_workspace_bytes(B, D), x0.device, "cross_tc_bwd")
      check(lib().tfrs_cross_tc_bwd_f32(ptr(x0), ptr(x), ptr(W), ptr(prod), ptr(g), B, D, D, c_f(ctx.diag), ptr(dx0), ptr(dx),
                                        ptr(dW), ptr(db), ptr(ws), ws.numel(), stream()), "cross_tc_bwd")
      return dx0, dx, dW, db, None, None
    wsb = lib().tfrs_cross_bwd_workspace_bytes(B, D)
    ws = workspace(wsb, x0.device, "cross")
    check(lib().tfrs_cross_bwd_f32(ptr(x0), ptr(x), ptr(W), ptr(prod), ptr(g), B, D, D, c_f(ctx.diag), ptr(dx0), ptr(dx),
                                   ptr(dW), ptr(db), ptr(ws), ws.numel(), stream()), "cross_bwd")
    return dx0, dx, dW, db, None, None


def cross(x0: torch.Tensor, x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], diag_scale: float = 0.0
          ) -> torch.Tensor:
  """x0 * (x @ W + bias + diag_scale * x) + x   (layers/feature_interaction/dcn.py:176-186).

  Stacked layers (`x = cross(x0, x)`): the output carries max |out| from the kernel's epilogue (`_tfrs_amax`), and a later
  call whose `x` is that very tensor, unmodified, skips its statistics pass over x -- same bits, one HBM pass less."""
  hint = getattr(x, "_tfrs_amax", None)
  x_amax = None
  if (hint is not None and hint[1] == x._version and hint[2] == x.data_ptr() and x.is_cuda and x.dtype == torch.float32 and
      x.is_contiguous() and x.shape == x0.shape):
    x_amax = hint[0]
  out, out_amax = _Cross.apply(x0, x, W, bias, float(diag_scale), x_amax)
  B, D = out.shape
  if B >= CROSS_TC_MIN_B and D >= CROSS_TC_MIN_D:
    out._tfrs_amax = (out_amax, out._version, out.data_ptr())
  return out


def gemm_tc(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
  """op(a) @ op(b) on the tensor cores with fp32 parity (split fp16, ~2^-21 relative error); any shape."""
  a = f32c(a, "a"); b = f32c(b, "b")
  M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
  N = b.shape[0] if trans_b else b.shape[1]
  if (b.shape[1] if trans_b else b.shape[0]) != K:
    raise ValueError(f"gemm_tc: inner dimensions differ ({tuple(a.shape)} x {tuple(b.shape)})")
  out = torch.empty((M, N), dtype=torch.float32, device=a.device)
  ws = workspace(lib().tfrs_gemm_tc_workspace_bytes(M, N, K), a.device, "gemm_tc")
  check(lib().tfrs_gemm_tc_f32(int(trans_a), int(trans_b), M, N, K, ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), N, ptr(ws),
                               ws.numel(), stream()), "gemm_tc")
  return out


class _CrossLowRank(torch.autograd.Function):
  """x0 * ((x @ U) @ V + bias + diag * x) + x on the tensor cores (forward: 2 GEMMs, formula fused; backward: 4 GEMMs)."""

  @staticmethod
  def forward(ctx, x0, x, U, V, bias, diag_scale):
    x0 = f32c(x0, "x0"); x = f32c(x, "x"); U = f32c(U, "kernel_u"); V = f32c(V, "kernel_v")
    b = None if bias is None else f32c(bias, "bias")
    B, D = x0.shape; p = U.shape[1]
    if U.shape != (D, p) or V.shape != (p, D):
      raise ValueError(f"cross_lowrank: kernels must be [{D},p] and [p,{D}], got {tuple(U.shape)} and {tuple(V.shape)}")
    out = torch.empty_like(x0)
    need_grad = any(ctx.needs_input_grad[:5])
    prod = torch.empty_like(x0) if need_grad else None
    t = torch.empty((B, p), dtype=torch.float32, device=x0.device)
    ws = workspace(lib().tfrs_cross_lowrank_tc_workspace_bytes(B, D, p), x0.device, "cross_lowrank")
    check(lib().tfrs_cross_lowrank_tc_fwd_f32(ptr(x0), ptr(x), ptr(U), ptr(V), ptr(b), B, D, p, D, c_f(diag_scale), ptr(out), ptr(prod),
                                              ptr(t), ptr(ws), ws.numel(), stream()), "cross_lowrank_tc_fwd")
    if need_grad:
      ctx.save_for_backward(x0, x, U, V, t, prod)
    ctx.diag = diag_scale
    ctx.has_bias = b is not None
    return out

  @staticmethod
  def backward(ctx, g):
    x0, x, U, V, t, prod = ctx.saved_tensors
    g = f32c(g, "grad")
    B, D = x0.shape; p = U.shape[1]
    n0, n1, n2, n3, n4 = ctx.needs_input_grad[:5]
    dx0 = torch.empty_like(x0) if n0 else None
    dx = torch.empty_like(x) if n1 else None
    dU = torch.empty_like(U) if n2 else None
    dV = torch.empty_like(V) if n3 else None
    db = torch.empty((D,), dtype=torch.float32, device=x0.device) if (n4 and ctx.has_bias) else None
    ws = workspace(lib().tfrs_cross_lowrank_tc_bwd_workspace_bytes(B, D, p), x0.device, "cross_lowrank_bwd")
    check(lib().tfrs_cross_lowrank_tc_bwd_f32(ptr(x0), ptr(x), ptr(U), ptr(V), ptr(t), ptr(prod), ptr(g), B, D, p, D, c_f(ctx.diag),
                                              ptr(dx0), ptr(dx), ptr(dU), ptr(dV), ptr(db), ptr(ws), ws.numel(), stream()),
          "cross_lowrank_tc_bwd")
    return dx0, dx, dU, dV, db, None


def cross_lowrank_supported(B: int, D: int, p: int) -> bool:
  return B >= CROSS_TC_MIN_B and D >= CROSS_TC_MIN_D and D <= 1024 and 1 <= p <= 1024


def cross_lowrank(x0: torch.Tensor, x: torch.Tensor, U: torch.Tensor, V: torch.Tensor, bias: Optional[torch.Tensor],
                  diag_scale: float = 0.0) -> torch.Tensor:
  """x0 * ((x @ U) @ V + bias + diag_scale * x) + x  (dcn.py:131-148,178-186; multi_layer_dcn.py:146-148)."""
  return _CrossLowRank.apply(x0, x, U, V, bias, float(diag_scale))


# ------------------------------------------------------------------------------------------------
# DotInteraction (DLRM pairwise feature dots)
# ------------------------------------------------------------------------------------------------
class _DotInteraction(torch.autograd.Function):

  @staticmethod
  def forward(ctx, feats, self_interaction, skip_gather):
    feats = f32c(feats, "features")
    B, F, d = feats.shape
    od = lib().tfrs_dot_interaction_out_dim(F, int(self_interaction), int(skip_gather))
    out = torch.empty((B, od), dtype=torch.float32, device=feats.device)
    check(lib().tfrs_dot_interaction_fwd_f32(ptr(feats), B, F, d, int(self_interaction), int(skip_gather), ptr(out), stream()),
          "dot_interaction_fwd")
    ctx.save_for_backward(feats)
    ctx.cfg = (int(self_interaction), int(skip_gather))
    return out

  @staticmethod
  def backward(ctx, g):
    (feats,) = ctx.saved_tensors
    g = f32c(g, "grad")
    B, F, d = feats.shape
    df = torch.empty_like(feats)
    check(lib().tfrs_dot_interaction_bwd_f32(ptr(feats), ptr(g), B, F, d, ctx.cfg[0], ctx.cfg[1], ptr(df), stream()),
          "dot_interaction_bwd")
    return df, None, None


def dot_interaction(feats: torch.Tensor, self_interaction: bool = False, skip_gather: bool = False) -> torch.Tensor:
  """feats [B, F, d] -> pairwise dots, lower triangle (dot_interaction.py:72-104)."""
  return _DotInteraction.apply(feats, bool(self_interaction), bool(skip_gather))


def launch_count() -> int:
  return int(lib().tfrs_launch_count())
