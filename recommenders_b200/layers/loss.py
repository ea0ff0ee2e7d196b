"""Loss-transform layers: mirror of tensorflow_recommenders/layers/loss.py (API kept; these are the
optional arguments of tfrs.tasks.Retrieval and run as plain tensor ops on the materialised logits)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

MAX_FLOAT = float(np.finfo(np.float32).max / 100.0)  # loss.py:22
MIN_FLOAT = float(np.finfo(np.float32).min / 100.0)  # loss.py:23


def _gather_elements_along_row(data: torch.Tensor, column_indices: torch.Tensor) -> torch.Tensor:
  """loss.py:26-58."""
  if data.shape[0] != column_indices.shape[0]:
    raise ValueError("The first dimensions of data and column_indices must match.")
  return torch.gather(data, 1, column_indices)


class HardNegativeMining(torch.nn.Module):
  """Transforms logits and labels to return hard negatives (loss.py:61-111)."""

  def __init__(self, num_hard_negatives: int) -> None:
    super().__init__()
    self._num_hard_negatives = num_hard_negatives

  def forward(self, logits: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    num_sampled = min(self._num_hard_negatives + 1, logits.shape[1])
    _, col_indices = torch.topk(logits + labels * MAX_FLOAT, k=num_sampled, dim=1, sorted=False)
    return _gather_elements_along_row(logits, col_indices), _gather_elements_along_row(labels, col_indices)


class RemoveAccidentalHits(torch.nn.Module):
  """Zeroes the logits of accidental negatives (loss.py:114-147)."""

  def forward(self, labels: torch.Tensor, logits: torch.Tensor, candidate_ids) -> torch.Tensor:
    if not isinstance(candidate_ids, torch.Tensor):
      candidate_ids = torch.as_tensor(np.asarray(candidate_ids), device=logits.device)
    candidate_ids = candidate_ids.to(logits.device).reshape(-1, 1)
    positive_indices = torch.argmax(labels, dim=1)
    positive_candidate_ids = candidate_ids[positive_indices]
    duplicate = (positive_candidate_ids == candidate_ids.t()).to(labels.dtype)
    duplicate = duplicate - labels
    return logits + duplicate * MIN_FLOAT


class SamplingProbablityCorrection(torch.nn.Module):
  """Sampling probability correction (loss.py:150-158)."""

  def forward(self, logits: torch.Tensor, candidate_sampling_probability: torch.Tensor) -> torch.Tensor:
    p = torch.as_tensor(candidate_sampling_probability, device=logits.device, dtype=logits.dtype)
    return logits - torch.log(torch.clamp(p, 1e-6, 1.0))
