"""Stacked low-rank cross layers: mirror of tensorflow_recommenders/layers/feature_interaction/multi_layer_dcn.py."""
from __future__ import annotations

from typing import Optional

import torch

from ... import ops
from .dcn import _init


class MultiLayerDCN(torch.nn.Module):
  """`num_layers` stacked low-rank cross layers sharing x0 (multi_layer_dcn.py:29-153):

      x_{l+1} = x0 .* ((x_l . U_l) . V_l + bias_l) + x_l          U_l [D, p], V_l [p, D]

  The two projections are the exact fp32 SGEMM (ops.matmul); kernels are stored [in, out] like Keras' Dense
  (:116-133)."""

  def __init__(self, projection_dim: Optional[int] = 1, num_layers: Optional[int] = 3, use_bias: bool = True,
               kernel_initializer="truncated_normal", bias_initializer="zeros", kernel_regularizer=None,
               bias_regularizer=None, **kwargs):
    super().__init__()
    self._projection_dim = projection_dim
    self._num_layers = num_layers
    self._use_bias = use_bias
    self._kernel_initializer = kernel_initializer
    self._bias_initializer = bias_initializer
    self._kernel_regularizer = kernel_regularizer
    self._bias_regularizer = bias_regularizer
    self._input_dim = None
    self.built = False
    self.name = kwargs.get("name")

  def build(self, input_shape, device=None):
    last_dim = int(input_shape[-1])
    device = device or torch.device("cuda", torch.cuda.current_device())
    self._input_dim = last_dim
    self.u_kernels = torch.nn.ParameterList(
        [torch.nn.Parameter(_init(self._kernel_initializer, (last_dim, self._projection_dim), device)) for _ in range(self._num_layers)])
    self.v_kernels = torch.nn.ParameterList(
        [torch.nn.Parameter(_init(self._kernel_initializer, (self._projection_dim, last_dim), device)) for _ in range(self._num_layers)])
    self.biases = torch.nn.ParameterList(
        [torch.nn.Parameter(_init(self._bias_initializer, (last_dim,), device)) for _ in range(self._num_layers)]) if self._use_bias else None
    self.built = True

  @property
  def losses(self):
    out = []
    if self.built and self._kernel_regularizer is not None:
      out += [self._kernel_regularizer(k) for k in list(self.u_kernels) + list(self.v_kernels)]
    if self.built and self._bias_regularizer is not None and self.biases is not None:
      out += [self._bias_regularizer(b) for b in self.biases]
    return out

  def call(self, x0: torch.Tensor) -> torch.Tensor:
    if not self.built:
      self.build(x0.shape, x0.device if isinstance(x0, torch.Tensor) else None)
    lead = x0.shape[:-1]
    x0f = x0.reshape(-1, x0.shape[-1]).to(torch.float32)
    xl = x0f
    fused = ops.cross_lowrank_supported(x0f.shape[0], x0f.shape[1], self._projection_dim)
    for i in range(self._num_layers):
      if fused:   # x0 * ((xl @ U) @ V + b) + xl: two tensor-core GEMMs per layer, formula in the second one's epilogue
        xl = ops.cross_lowrank(x0f, xl, self.u_kernels[i], self.v_kernels[i], None if self.biases is None else self.biases[i], 0.0)
        continue
      prod = ops.matmul(ops.matmul(xl, self.u_kernels[i]), self.v_kernels[i])   # (:146-147)
      if self.biases is not None:
        prod = prod + self.biases[i]
      xl = x0f * prod + xl                                                        # (:148)
    return xl.reshape(*lead, -1)

  def forward(self, x0):
    return self.call(x0)

  def get_config(self):
    return {
        "projection_dim": self._projection_dim, "num_layers": self._num_layers, "use_bias": self._use_bias,
        "kernel_initializer": self._kernel_initializer if isinstance(self._kernel_initializer, str) else "custom",
        "bias_initializer": self._bias_initializer if isinstance(self._bias_initializer, str) else "custom",
        "kernel_regularizer": None if self._kernel_regularizer is None else getattr(self._kernel_regularizer, "__name__", "custom"),
        "bias_regularizer": None if self._bias_regularizer is None else getattr(self._bias_regularizer, "__name__", "custom"),
        "name": self.name,
    }

  @classmethod
  def from_config(cls, config):
    cfg = dict(config)
    cfg.pop("kernel_regularizer", None); cfg.pop("bias_regularizer", None)
    return cls(**cfg)
