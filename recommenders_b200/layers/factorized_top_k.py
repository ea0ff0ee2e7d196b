"""Top-K retrieval layers: the B200 mirror of tensorflow_recommenders/layers/factorized_top_k.py.

Same classes, constructor arguments, method names and error behaviour as the reference
(`TopK` :140-333, `Streaming` :336-512, `BruteForce` :515-610, `ScaNN` stub :613-796); tensors are CUDA
`torch.Tensor`s and the arithmetic runs in libtfrs_b200.so (exact fp32 scan or tcgen05 screening +
exact rescoring).  Results follow tf.math.top_k's contract: scores descending, ties -> lower index.
"""
from __future__ import annotations

import abc
from typing import Dict, Optional, Text, Tuple, Union

import numpy as np
import torch

from .. import ops
from ..data import Dataset, as_dataset

Tensor = torch.Tensor
Identifiers = Union[torch.Tensor, np.ndarray]


def _wrap_batch_too_small_error(k: int) -> ValueError:
  """factorized_top_k.py:34-54 -- the helpful message for chunks smaller than k."""
  return ValueError(
      "Tried to retrieve k={k} top items, but the candidate "
      "dataset batch size is too small. This may be because "
      "your candidate batch size is too small or the last "
      "batch of your dataset is too small. "
      "To resolve this, increase your batch size, set the "
      "drop_remainder argument to True when batching your "
      "candidates, or set the handle_incomplete_batches "
      "argument to True in the constructor. ".format(k=k))


def _take_along_axis(arr, indices: Tensor):
  """factorized_top_k.py:57-80 -- arr[i, indices[i, j]] for torch tensors or NumPy arrays."""
  if isinstance(arr, np.ndarray):
    return np.take_along_axis(arr, indices.cpu().numpy(), 1)
  return torch.gather(arr, 1, indices)


def _gather_identifiers(identifiers: Identifiers, idx: Tensor):
  """tf.gather(identifiers, indices) (:607, :438): torch ids stay on device, others go through NumPy."""
  if isinstance(identifiers, torch.Tensor):
    return identifiers[idx]
  return np.asarray(identifiers)[idx.cpu().numpy()]


def _exclude(scores: Tensor, identifiers, exclude, k: int):
  """Removes a subset of candidates from top K candidates (factorized_top_k.py:83-115).

  Scores of excluded identifiers are lowered by 1e5, the top min(k, cols) of the adjusted scores are
  taken, and the ORIGINAL scores / identifiers at those positions are returned."""
  if isinstance(identifiers, torch.Tensor):
    exclude_t = exclude if isinstance(exclude, torch.Tensor) else torch.as_tensor(np.asarray(exclude))
    isin = (identifiers.unsqueeze(-1) == exclude_t.to(identifiers.device).unsqueeze(1)).any(-1)
  else:
    ex = exclude.cpu().numpy() if isinstance(exclude, torch.Tensor) else np.asarray(exclude)
    isin = torch.from_numpy((np.asarray(identifiers)[:, :, None] == ex[:, None, :]).any(-1)).to(scores.device)
  adjusted = scores - isin.to(torch.float32) * 1.0e5
  k = min(k, scores.shape[1])
  pos = torch.arange(scores.shape[1], device=scores.device, dtype=torch.int64).expand_as(scores)
  _, indices = ops.topk_merge(adjusted.unsqueeze(0), pos.unsqueeze(0), k)  # top_k(adjusted): ties -> lower index
  return _take_along_axis(scores, indices), _take_along_axis(identifiers, indices)


def _check_candidates_with_identifiers(element) -> None:
  """factorized_top_k.py:118-137 -- dataset elements are embeddings or (identifiers, embeddings)."""
  if isinstance(element, tuple):
    if len(element) != 2:
      raise ValueError("The dataset must yield candidate embeddings or "
                       "tuples of (candidate identifiers, candidate embeddings). "
                       f"Got a tuple of length {len(element)} instead.")
    ids, emb = element
    if emb.shape[0] != ids.shape[0]:
      raise ValueError("Candidates and identifiers have to have the same batch dimension. "
                       f"Got {emb.shape[0]} and {ids.shape[0]}.")


def _concat_ids(chunks):
  if isinstance(chunks[0], torch.Tensor):
    return torch.cat(chunks, 0)
  return np.concatenate([np.asarray(c) for c in chunks], 0)


def shard_bounds(num_rows: int, rank: int, world: int) -> Tuple[int, int]:
  """Contiguous row block [lo, hi) of shard `rank`: global index order == (shard, local index) order, so
  the lowest-index tie rule survives the merge."""
  per = (num_rows + world - 1) // world
  lo = min(rank * per, num_rows)
  return lo, min(lo + per, num_rows)


def allgather_topk(scores: Tensor, idx: Tensor, k: int, group=None) -> Tuple[Tensor, Tensor]:
  """The single collective of the sharded scan: one all-gather of every rank's packed [Q,k] (score, index)
  list -> ([world,Q,k] f32, [world,Q,k] i64).  Short shards are padded with (-inf, INT64_MAX)."""
  import torch.distributed as dist
  world = dist.get_world_size(group)
  if scores.shape[1] < k:
    pad = k - scores.shape[1]
    scores = torch.cat([scores, torch.full((scores.shape[0], pad), float("-inf"), device=scores.device)], 1)
    idx = torch.cat([idx, torch.full((idx.shape[0], pad), torch.iinfo(torch.int64).max, device=idx.device,
                                     dtype=torch.int64)], 1)
  # scores ride as raw bits in an int64 lane next to the indices: one buffer, one collective
  packed = torch.stack([scores.contiguous().view(torch.int32).to(torch.int64), idx.to(torch.int64)], 0).contiguous()
  flat = torch.empty((world * 2,) + tuple(packed.shape[1:]), dtype=torch.int64, device=packed.device)
  dist.all_gather_into_tensor(flat, packed, group=group)
  gathered = flat.view((world, 2) + tuple(packed.shape[1:]))
  all_s = gathered[:, 0].to(torch.int32).view(torch.float32).contiguous()
  all_i = gathered[:, 1].contiguous()
  return all_s, all_i


class TopK(torch.nn.Module, abc.ABC):
  """Interface for top K layers (factorized_top_k.py:140-333)."""

  def __init__(self, k: int, **kwargs) -> None:
    name = kwargs.pop("name", None)
    super().__init__()
    self._k = k
    self.name = name

  @abc.abstractmethod
  def index(self, candidates: Tensor, identifiers: Optional[Identifiers] = None) -> "TopK":
    raise NotImplementedError()

  def index_from_dataset(self, candidates) -> "TopK":
    """Builds the retrieval index from a dataset of embeddings or (identifier, embedding) batches (:179-215)."""
    ds = as_dataset(candidates)
    elements = list(ds)
    for el in elements:
      _check_candidates_with_identifiers(el)
    if elements and isinstance(elements[0], tuple):
      cands = torch.cat([emb for _, emb in elements], 0)
      identifiers = _concat_ids([ids for ids, _ in elements])
    else:
      cands = torch.cat(elements, 0)
      identifiers = None
    return self.index(cands, identifiers)

  @abc.abstractmethod
  def call(self, queries: Union[Tensor, Dict[Text, Tensor]], k: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    raise NotImplementedError()

  def forward(self, queries, k: Optional[int] = None):
    return self.call(queries, k=k)

  def query_with_exclusions(self, queries, exclusions, k: Optional[int] = None):
    """Query the index, excluding per-query identifiers (:242-288): over-fetch k+E, then `_exclude`."""
    k = k if k is not None else self._k
    adjusted_k = k + exclusions.shape[1]
    x, y = self(queries=queries, k=adjusted_k)
    return _exclude(x, y, exclude=exclusions, k=k)

  @abc.abstractmethod
  def is_exact(self) -> bool:
    raise NotImplementedError()

  def _reset_tf_function_cache(self):
    """No tf.function cache exists here (:303-318); kept so subclasses written for the reference still run."""

  def _compute_score(self, queries: Tensor, candidates: Tensor) -> Tensor:
    """The standard dot product score matmul(q, c^T) (:320-333) -- exact fp32."""
    return ops.scores(queries, candidates)


class Streaming(TopK):
  """Retrieves K highest scoring items and their ids from a large dataset (factorized_top_k.py:336-512).

  Each dataset batch is scanned on the GPU and merged into the carried [Q,k] state by the same kernel
  (state entries compete with their own indices), so the result equals BruteForce's."""

  def __init__(self, query_model: Optional[torch.nn.Module] = None, k: int = 10,
               handle_incomplete_batches: bool = True, num_parallel_calls: Optional[int] = None,
               sorted_order: bool = True) -> None:
    super().__init__(k=k)
    self.query_model = query_model
    self._candidates = None
    self._handle_incomplete_batches = handle_incomplete_batches
    self._num_parallel_calls = num_parallel_calls
    self._sorted = sorted_order
    self._coalesce_rows = 65536
    self.register_buffer("_counter", torch.zeros((), dtype=torch.int32), persistent=False)

  def index_from_dataset(self, candidates) -> "TopK":
    ds = as_dataset(candidates)
    self._candidates = ds
    return self

  def index(self, candidates, identifiers=None) -> "Streaming":
    """Not implemented. Please call `index_from_dataset` instead (:392-402)."""
    raise NotImplementedError("The streaming top k class only accepts datasets. "
                              "Please call `index_from_dataset` instead.")

  def call(self, queries, k: Optional[int] = None):
    k = k if k is not None else self._k
    if self._candidates is None:
      raise ValueError("The `index` method must be called first to create the retrieval index.")
    if self.query_model is not None:
      queries = self.query_model(queries)
    Q = queries.shape[0]
    state = (torch.zeros((Q, 0), dtype=torch.float32, device=queries.device),
             torch.zeros((Q, 0), dtype=torch.int64, device=queries.device))
    counter = 0
    id_chunks = []
    has_ids = False
    pending, pending_rows = [], 0

    def flush():
      nonlocal state, counter, pending, pending_rows
      if not pending:
        return
      emb = pending[0] if len(pending) == 1 else torch.cat(pending, 0)
      # the scan kernel takes the carried state and numbers the rows with the running counter
      # (enumerate_rows, :474-485); ties resolve to the lower running index == state first (:462-463).
      state = ops.topk_scan(queries, emb, k, index_offset=counter, state=state)
      counter += int(emb.shape[0])
      pending, pending_rows = [], 0

    for el in self._candidates:
      _check_candidates_with_identifiers(el)
      if isinstance(el, tuple):
        ids, emb = el
        has_ids = True
        id_chunks.append(ids)
      else:
        emb = el
      if not self._handle_incomplete_batches and emb.shape[0] < k:
        raise _wrap_batch_too_small_error(k)
      # Dataset batches are tiny (README uses 128): coalesce them into >= 64K-row scans.  The result is the
      # same as merging per batch -- indices are the running row numbers either way.
      pending.append(emb); pending_rows += int(emb.shape[0])
      if pending_rows >= self._coalesce_rows:
        flush()
    flush()
    self._counter.fill_(counter)
    scores, idx = state
    if has_ids:
      return scores, _gather_identifiers(_concat_ids(id_chunks), idx)
    return scores, idx.to(torch.int32)

  def is_exact(self) -> bool:
    return True


class BruteForce(TopK):
  """Brute force retrieval (factorized_top_k.py:515-610).

  `index` keeps the fp32 corpus and, for large corpora, builds the bf16 tensor-core screening image;
  `call` returns exactly the top-k of the fp32 scores either way.  `index_shard` adds the row-sharded
  multi-GPU mode: every rank scans its contiguous shard and one all-gather of the per-shard (score,
  index) lists is merged on every rank."""

  def __init__(self, query_model: Optional[torch.nn.Module] = None, k: int = 10, name: Optional[Text] = None):
    super().__init__(k=k, name=name)
    self.query_model = query_model
    self._candidates = None
    self._identifiers = None
    self._tc_index = None
    self._shard = None  # (global_offset, group)
    self.use_tensor_cores = True

  def index(self, candidates: Tensor, identifiers: Optional[Identifiers] = None) -> "BruteForce":
    if identifiers is None:
      identifiers_ = None
      n_ids = candidates.shape[0]
    else:
      identifiers_ = identifiers
      n_ids = identifiers.shape[0]
    if candidates.dim() != 2:
      raise ValueError(f"The candidates tensor must be 2D (got {tuple(candidates.shape)}).")
    if candidates.shape[0] != n_ids:
      raise ValueError("The candidates and identifiers tensors must have the same number of"
                       f" rows (got {candidates.shape[0]} candidates rows and"
                       f" {n_ids} identifier rows). ")
    self._set_index(ops.f32c(candidates, "candidates").detach(), identifiers_)
    self._shard = None
    self._reset_tf_function_cache()
    return self

  def index_shard(self, local_candidates: Tensor, global_offset: int, identifiers: Optional[Identifiers] = None,
                  group=None) -> "BruteForce":
    """Row-sharded index: this rank owns corpus rows [global_offset, global_offset + len(local_candidates)).
    `identifiers`, when given, covers the WHOLE corpus (it is only used to map the merged indices)."""
    if local_candidates.dim() != 2:
      raise ValueError(f"The candidates tensor must be 2D (got {tuple(local_candidates.shape)}).")
    self._set_index(ops.f32c(local_candidates, "candidates").detach(), identifiers)
    self._shard = (int(global_offset), group)
    return self

  def _set_index(self, cands: Tensor, identifiers) -> None:
    self._candidates = cands
    self._identifiers = identifiers
    self._tc_index = None
    if self.use_tensor_cores and cands.shape[0] >= ops.TC_MIN_N and cands.shape[1] <= 128:
      self._tc_index = ops.index_build(cands)

  def _local_topk(self, queries: Tensor, k: int, offset: int, out=None):
    if self._tc_index is not None and ops.tc_supported(queries.shape[0], self._candidates.shape[0],
                                                       self._candidates.shape[1], k):
      return ops.topk_tc(queries, self._candidates, self._tc_index, k, index_offset=offset, out=out)
    return ops.topk_scan(queries, self._candidates, k, index_offset=offset, out=out)

  def call(self, queries, k: Optional[int] = None):
    k = k if k is not None else self._k
    if self._candidates is None:
      raise ValueError("The `index` method must be called first to create the retrieval index.")
    if self.query_model is not None:
      queries = self.query_model(queries)
    if self._shard is None:
      n_total = self._candidates.shape[0]
      if k > n_total:
        raise ValueError(f"input must have at least k columns. Had {n_total}, needed {k}")
      values, indices = self._local_topk(queries, k, 0)
    else:
      values, indices = self._sharded_topk(queries, k)
    if self._identifiers is None:
      return values, indices.to(torch.int32)  # default identifiers = range(N) int32 (:544-545)
    return values, _gather_identifiers(self._identifiers, indices)

  def _sharded_topk(self, queries: Tensor, k: int):
    """Local scan -> ONE all-gather of every rank's packed [scores | indices] block -> merge kernel reading the
    receive buffer in place.  The local scan writes straight into the send block (no packing kernels)."""
    import torch.distributed as dist
    offset, group = self._shard
    world = dist.get_world_size(group)
    Q = queries.shape[0]
    k_local = min(k, self._candidates.shape[0])
    if k_local < k:  # a shard smaller than k: rectangular lists via the generic (padding) path
      s, i = self._local_topk(queries, k_local, offset)
      all_s, all_i = allgather_topk(s, i, k, group)
      return ops.topk_merge(all_s, all_i, k)
    idx_off = (Q * k * 4 + 7) // 8 * 8
    block = idx_off + Q * k * 8
    send = torch.empty(block, dtype=torch.uint8, device=queries.device)
    out_s = send[:Q * k * 4].view(torch.float32).view(Q, k)
    out_i = send[idx_off:].view(torch.int64).view(Q, k)
    self._local_topk(queries, k, offset, out=(out_s, out_i))
    recv = torch.empty(world * block, dtype=torch.uint8, device=queries.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return ops.topk_merge_packed(recv, world, Q, k, k, idx_off, block)

  def is_exact(self) -> bool:
    return True

  # -- checkpointing: the index is model state (reference keeps it as non-trainable weights, :562-580)
  def get_extra_state(self):
    ids = self._identifiers
    return {"candidates": None if self._candidates is None else self._candidates.cpu(),
            "identifiers": ids.cpu() if isinstance(ids, torch.Tensor) else ids}

  def set_extra_state(self, state):
    if state and state.get("candidates") is not None:
      dev = torch.device("cuda", torch.cuda.current_device())
      ids = state.get("identifiers")
      if isinstance(ids, torch.Tensor):
        ids = ids.to(dev)
      self.index(state["candidates"].to(dev), ids)


class ScaNN(TopK):
  """ScaNN approximate retrieval lives in the un-vendored `scann` pip package (factorized_top_k.py:25-31);
  as in the reference without that package, constructing it raises ImportError (:675-679)."""

  def __init__(self, *args, **kwargs):
    raise ImportError("The scann library is not present. Please install it using `pip install scann` to use "
                      "the ScaNN layer.")

  def index(self, candidates, identifiers=None):  # pragma: no cover
    raise NotImplementedError()

  def call(self, queries, k=None):  # pragma: no cover
    raise NotImplementedError()

  def is_exact(self) -> bool:  # pragma: no cover
    return False
