"""Factorized top-K metrics: mirror of tensorflow_recommenders/metrics/factorized_top_k.py."""
from __future__ import annotations

import abc
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import ops
from .data import Dataset
from .layers import factorized_top_k as ftk


class Mean:
  """`tf.keras.metrics.Mean`: weighted running mean."""

  def __init__(self, name: str = "mean"):
    self.name = name
    self.reset_states()

  def reset_states(self) -> None:
    self._total = 0.0
    self._count = 0.0

  reset_state = reset_states

  def update_state(self, values, sample_weight=None) -> None:
    v = values.to(torch.float32) if isinstance(values, torch.Tensor) else torch.as_tensor(values, dtype=torch.float32)
    if sample_weight is None:
      self._total += float(v.sum()); self._count += float(v.numel())
    else:
      w = sample_weight.to(torch.float32).to(v.device) if isinstance(sample_weight, torch.Tensor) \
          else torch.as_tensor(sample_weight, dtype=torch.float32, device=v.device)
      if w.numel() == v.numel():
        w = w.reshape(v.shape)
      elif w.dim() <= v.dim():
        w = torch.broadcast_to(w.reshape(w.shape + (1,) * (v.dim() - w.dim())), v.shape)
      self._total += float((v * w).sum()); self._count += float(w.sum())

  def result(self) -> float:
    return self._total / self._count if self._count else 0.0


class TopKCategoricalAccuracy(Mean):
  """`tf.keras.metrics.TopKCategoricalAccuracy` (used as a batch metric in tasks/retrieval_test.py:44-47)."""

  def __init__(self, k: int = 5, name: str = "top_k_categorical_accuracy"):
    super().__init__(name)
    self.k = k

  def update_state(self, y_true, y_pred, sample_weight=None) -> None:
    target = y_true.argmax(dim=1)
    t = y_pred.gather(1, target[:, None])
    hit = ((y_pred > t).sum(1) < self.k).to(torch.float32)
    super().update_state(hit, sample_weight)


def in_top_k(targets: torch.Tensor, predictions: torch.Tensor, k: int) -> torch.Tensor:
  """tf.math.in_top_k: target's prediction is finite and fewer than k predictions are strictly larger."""
  t = predictions.gather(1, targets.to(torch.int64)[:, None])
  return torch.isfinite(t[:, 0]) & ((predictions > t).sum(1) < k)


class Factorized(torch.nn.Module, abc.ABC):
  """Computes metrics across top K candidates surfaced by a retrieval model (:27-49)."""

  @abc.abstractmethod
  def update_state(self, query_embeddings, true_candidate_embeddings, true_candidate_ids=None):
    raise NotImplementedError()

  @property
  def metrics(self) -> List[Mean]:
    return []

  def reset_states(self) -> None:
    for metric in self.metrics:
      metric.reset_states()

  def result(self) -> List[float]:
    return [metric.result() for metric in self.metrics]


class FactorizedTopK(Factorized):
  """Top-K categorical accuracy across the candidates surfaced by a retrieval layer (:52-194)."""

  def __init__(self, candidates: Union[ftk.TopK, Dataset, list], ks: Sequence[int] = (1, 5, 10, 50, 100),
               name: str = "factorized_top_k") -> None:
    super().__init__()
    self.name = name
    if not isinstance(candidates, ftk.TopK):
      candidates = ftk.Streaming(k=max(ks)).index_from_dataset(candidates)  # :77-81
    self._ks = ks
    self._candidates = candidates
    self._top_k_metrics = [Mean(name=f"{self.name}/top_{x}_categorical_accuracy") for x in ks]

  @property
  def metrics(self) -> List[Mean]:
    return self._top_k_metrics

  @torch.no_grad()
  def update_state(self, query_embeddings: torch.Tensor, true_candidate_embeddings: torch.Tensor,
                   true_candidate_ids=None, sample_weight=None) -> None:
    if true_candidate_ids is None and not self._candidates.is_exact():
      raise ValueError(f"The candidate generation layer ({self._candidates}) does not return "
                       "exact results. To perform evaluation using that layer, you must "
                       "supply `true_candidate_ids`, which will be checked against "
                       "the candidate ids returned from the candidate generation layer.")
    # positive score with the same canonical chain as the retrieved scores (:133-134)
    positive_scores = ops.rowwise_dot(query_embeddings, true_candidate_embeddings)[:, None]
    top_k_predictions, retrieved_ids = self._candidates(query_embeddings, k=max(self._ks))

    if true_candidate_ids is not None:
      nan_padding = torch.isnan(top_k_predictions)
      top_k_predictions = torch.where(nan_padding, torch.full_like(top_k_predictions, torch.finfo(torch.float32).min),
                                      top_k_predictions)
      if top_k_predictions.shape[1] > 1 and bool(((top_k_predictions[:, :-1] - top_k_predictions[:, 1:]) < 0).any()):
        raise AssertionError("Top-K predictions must be sorted.")
      if isinstance(retrieved_ids, torch.Tensor):
        tid = true_candidate_ids if isinstance(true_candidate_ids, torch.Tensor) else \
            torch.as_tensor(np.asarray(true_candidate_ids))
        tid = tid.to(retrieved_ids.device)
        if tid.dim() == 1:
          tid = tid[:, None]
        eq = (tid == retrieved_ids)
      else:
        tid = true_candidate_ids.cpu().numpy() if isinstance(true_candidate_ids, torch.Tensor) \
            else np.asarray(true_candidate_ids)
        if tid.ndim == 1:
          tid = tid[:, None]
        eq = torch.from_numpy(np.asarray(tid == np.asarray(retrieved_ids))).to(top_k_predictions.device)
      ids_match = (eq & ~nan_padding).to(torch.float32)
      for k, metric in zip(self._ks, self._top_k_metrics):
        match_found = torch.clamp(ids_match[:, :k].sum(1, keepdim=True), 0.0, 1.0)
        metric.update_state(match_found, sample_weight)
    else:
      y_pred = torch.cat([positive_scores, top_k_predictions], dim=1)
      targets = torch.zeros(y_pred.shape[0], dtype=torch.int64, device=y_pred.device)
      for k, metric in zip(self._ks, self._top_k_metrics):
        metric.update_state(in_top_k(targets, y_pred, k).to(torch.float32),
                            None if sample_weight is None else _flat(sample_weight))


def _flat(w):
  return w.reshape(-1) if isinstance(w, torch.Tensor) else np.asarray(w).reshape(-1)
