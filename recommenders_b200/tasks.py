"""Tasks: mirror of tensorflow_recommenders/tasks/{base,retrieval}.py."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Text, Union

import torch

from . import metrics as tfrs_metrics
from . import ops
from .layers import loss as loss_layers

MIN_FLOAT = loss_layers.MIN_FLOAT  # tasks/retrieval.py:25


class Task:
  """Task marker class (tasks/base.py:23-30)."""


def _categorical_crossentropy_sum(y_true: torch.Tensor, y_pred: torch.Tensor, sample_weight=None) -> torch.Tensor:
  """CategoricalCrossentropy(from_logits=True, reduction=SUM) -- the default loss, retrieval.py:86-87."""
  per = -(y_true * torch.log_softmax(y_pred, dim=1)).sum(1)
  if sample_weight is not None:
    per = per * torch.as_tensor(sample_weight, dtype=per.dtype, device=per.device).reshape(-1)
  return per.sum()


class Retrieval(torch.nn.Module, Task):
  """A factorized retrieval task (tasks/retrieval.py:29-235).

  2-D queries with the default loss run fused and never materialise the [B, C] logits or the eye() labels:
  `temperature`, the sampling-probability correction, accidental-hit removal and `score_mask` are folded into the
  tensor-core loss kernels, hard-negative mining runs on the top-K scan.  A custom loss object, multi-head (3-D)
  queries, batch metrics, or shapes outside the tensor-core range use the exact score matrix + the reference's
  op sequence."""

  def __init__(self, loss: Optional[Callable] = None,
               metrics: Optional[Union[Sequence[tfrs_metrics.Factorized], tfrs_metrics.Factorized]] = None,
               batch_metrics: Optional[List] = None, loss_metrics: Optional[List] = None,
               temperature: Optional[float] = None, num_hard_negatives: Optional[int] = None,
               remove_accidental_hits: bool = False, name: Optional[Text] = None) -> None:
    super().__init__()
    self.name = name
    self._loss = loss
    if metrics is None:
      metrics = []
    if not isinstance(metrics, Sequence):
      metrics = [metrics]
    self._factorized_metrics = list(metrics)
    self._batch_metrics = batch_metrics or []
    self._loss_metrics = loss_metrics or []
    self._temperature = temperature
    self._num_hard_negatives = num_hard_negatives
    self._remove_accidental_hits = remove_accidental_hits

  @property
  def factorized_metrics(self):
    """The metrics object used to compute retrieval metrics (:101-106)."""
    return self._factorized_metrics

  @factorized_metrics.setter
  def factorized_metrics(self, value) -> None:
    if not isinstance(value, Sequence):
      value = []
    self._factorized_metrics = list(value)

  @property
  def metrics(self):
    """Flat list of metric objects, Keras `layer.metrics` order: factorized, batch, loss metrics."""
    out = []
    for m in self._factorized_metrics:
      out.extend(m.metrics)
    return out + list(self._batch_metrics) + list(self._loss_metrics)

  def call(self, query_embeddings: torch.Tensor, candidate_embeddings: torch.Tensor,
           sample_weight: Optional[torch.Tensor] = None, candidate_sampling_probability: Optional[torch.Tensor] = None,
           candidate_ids=None, compute_metrics: bool = True, compute_batch_metrics: bool = True,
           score_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    three_d = query_embeddings.dim() == 3
    if self._remove_accidental_hits and candidate_ids is None:
      raise ValueError("When accidental hit removal is enabled, candidate ids must be supplied.")
    wants_batch_scores = compute_batch_metrics and len(self._batch_metrics) > 0
    options = (candidate_sampling_probability is not None or self._remove_accidental_hits or score_mask is not None)
    # Everything except a custom loss object, 3-D (multi-head) queries and batch metrics (which need the logits) runs
    # fused: temperature, sampling-probability correction (a per-candidate bias), accidental-hit removal (candidate ids
    # compared in the epilogue) and score_mask (keep-bits) inside the tensor-core loss; hard-negative mining through the
    # top-K scan + a sparse loss (`ops.hard_negative_softmax_loss`).  Nothing of size [B, C] is materialised there.
    fusable = not three_d and self._loss is None and not wants_batch_scores
    B_, C_, d_ = query_embeddings.shape[0], candidate_embeddings.shape[0], query_embeddings.shape[-1]
    fused_hard = (fusable and self._num_hard_negatives is not None and not options and
                  (self._temperature is None or self._temperature > 0) and
                  ops.hard_negative_supported(B_, C_, d_, self._num_hard_negatives))
    fused_opts = (fusable and self._num_hard_negatives is None and options and
                  ops.inbatch_softmax_bias_supported(B_, C_, d_))
    plain = not three_d and self._loss is None and self._num_hard_negatives is None and not options
    # multi-head queries (maxsim, :172-176) with the default loss: the head maximum is folded inside the blocked loss kernels
    fused_maxsim = (three_d and self._loss is None and self._num_hard_negatives is None and not options and
                    (self._temperature is None or self._temperature > 0))
    need_scores = wants_batch_scores or not (fused_hard or fused_opts or plain or fused_maxsim)

    scores = labels = None
    if need_scores:
      if three_d:  # maxsim over query heads, :172-176
        nq, nh, e = query_embeddings.shape
        s = ops.scores(query_embeddings.reshape(nq * nh, e), candidate_embeddings)
        scores = s.reshape(nq, nh, -1).max(dim=1).values
      else:
        scores = ops.scores(query_embeddings, candidate_embeddings)  # :178-180
      num_queries, num_candidates = scores.shape
      labels = torch.eye(num_queries, num_candidates, device=scores.device)  # :185
      if self._temperature is not None:
        scores = scores / self._temperature
      if candidate_sampling_probability is not None:
        scores = loss_layers.SamplingProbablityCorrection()(scores, candidate_sampling_probability)
      if self._remove_accidental_hits:
        scores = loss_layers.RemoveAccidentalHits()(labels, scores, candidate_ids)
      if score_mask is not None:
        scores = torch.where(score_mask.to(scores.device).bool(), scores, torch.full_like(scores, MIN_FLOAT))
      if self._num_hard_negatives is not None:
        scores, labels = loss_layers.HardNegativeMining(self._num_hard_negatives)(scores, labels)

    if fused_opts:
      bias = None
      if candidate_sampling_probability is not None:
        # logits - log(clip(p, 1e-6, 1))  (layers/loss.py:150-158) as a bias vector
        p_c = torch.as_tensor(candidate_sampling_probability, dtype=torch.float32, device=candidate_embeddings.device).reshape(-1)
        bias = -torch.log(torch.clamp(p_c, 1e-6, 1.0))
      loss = ops.inbatch_softmax_loss(query_embeddings, candidate_embeddings, sample_weight, self._temperature, bias,
                                      candidate_ids if self._remove_accidental_hits else None,
                                      None if score_mask is None else score_mask.to(candidate_embeddings.device))
    elif fused_hard:
      loss = ops.hard_negative_softmax_loss(query_embeddings, candidate_embeddings, self._num_hard_negatives, sample_weight,
                                            self._temperature)
    elif fused_maxsim:
      loss = ops.inbatch_softmax_maxsim_loss(query_embeddings, candidate_embeddings, sample_weight, self._temperature)
    elif plain:
      loss = ops.inbatch_softmax_loss(query_embeddings, candidate_embeddings, sample_weight, self._temperature)
    elif self._loss is not None:
      loss = self._loss(labels, scores, sample_weight) if sample_weight is not None else self._loss(labels, scores)
    else:
      loss = _categorical_crossentropy_sum(labels, scores, sample_weight)

    with torch.no_grad():
      for metric in self._loss_metrics:
        metric.update_state(loss.detach())
      if compute_metrics and not three_d:
        for metric in self._factorized_metrics:
          metric.update_state(query_embeddings.detach(),
                              candidate_embeddings[:query_embeddings.shape[0]].detach(),  # :221-223
                              true_candidate_ids=candidate_ids, sample_weight=sample_weight)
      if compute_batch_metrics:
        for metric in self._batch_metrics:
          metric.update_state(labels, scores.detach(), sample_weight=sample_weight)
    return loss

  def forward(self, *args, **kwargs):
    return self.call(*args, **kwargs)
