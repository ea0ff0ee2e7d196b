"""DLRM dot interaction: mirror of tensorflow_recommenders/layers/feature_interaction/dot_interaction.py."""
from __future__ import annotations

from typing import List, Optional

import torch

from ... import ops


class DotInteraction(torch.nn.Module):
  """Dot interaction layer (dot_interaction.py:23-104).

  Applied to a list of tensors [e1, ..., ek] of the same shape [batch, dim]; the output holds all distinct
  pairwise dot products dot(e_i, e_j), i <= j if `self_interaction` else i < j, in the order of the lower
  triangle of the interaction matrix; with `skip_gather` the full [num_features * num_features] matrix is
  returned with the upper triangle zeroed."""

  def __init__(self, self_interaction: bool = False, skip_gather: bool = False, name: Optional[str] = None, **kwargs):
    super().__init__()
    self._self_interaction = self_interaction
    self._skip_gather = skip_gather
    self.name = name

  def call(self, inputs: List[torch.Tensor]) -> torch.Tensor:
    try:
      dims = {int(t.shape[1]) for t in inputs}
      batches = {int(t.shape[0]) for t in inputs}
      if len(dims) != 1 or len(batches) != 1 or any(t.dim() != 2 for t in inputs):
        raise ValueError(f"got shapes {[tuple(t.shape) for t in inputs]}")
      feats = torch.stack([t.to(torch.float32) for t in inputs], dim=1)  # [batch, num_features, dim] (:75-77)
    except (ValueError, RuntimeError, IndexError) as e:
      raise ValueError(f"Input tensors` dimensions must be equal, original"
                       f"error message: {e}")
    return ops.dot_interaction(feats, self._self_interaction, self._skip_gather)

  def forward(self, inputs):
    return self.call(inputs)

  def get_config(self):
    return {"self_interaction": self._self_interaction, "skip_gather": self._skip_gather, "name": self.name}

  @classmethod
  def from_config(cls, config):
    return cls(**config)
