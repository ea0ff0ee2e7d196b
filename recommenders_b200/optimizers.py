"""Adagrad with the sparse row update of the retrieval hot path (the optimizer the reference's README
passes to `model.compile`, README.md:84: `tf.keras.optimizers.Adagrad(0.5)`).

Keras defaults: initial_accumulator_value=0.1, epsilon=1e-7.  `eps_inside_sqrt=True` is the Keras-3 /
tf-keras `optimizers.Adagrad` rule  var -= lr*g/sqrt(acc+eps); False is the legacy
`optimizers.legacy.Adagrad` rule  var -= lr*g/(sqrt(acc)+eps)  (SURVEY.md A10 -- third-party, unpinned)."""
from __future__ import annotations

from typing import Iterable, List

import torch

from . import ops
from .layers.embedding import Embedding


class Adagrad:

  def __init__(self, learning_rate: float = 0.001, initial_accumulator_value: float = 0.1, epsilon: float = 1e-7,
               eps_inside_sqrt: bool = True):
    self.learning_rate = learning_rate
    self.initial_accumulator_value = initial_accumulator_value
    self.epsilon = epsilon
    self.eps_inside_sqrt = eps_inside_sqrt
    self._module = None
    self._dense: List[torch.nn.Parameter] = []
    self._tables: List[Embedding] = []
    self._acc = {}

  def bind(self, module: torch.nn.Module) -> "Adagrad":
    """Attach to a model.  The variable lists are re-read on every step (`_refresh`), like the reference's
    `self.trainable_variables` at models/base.py:77: layers that create their weights lazily on the first
    forward (Cross, MultiLayerDCN) are picked up even when compile() ran before the first batch."""
    self._module = module
    self._refresh()
    return self

  def _refresh(self) -> None:
    if self._module is None:
      return
    self._tables = [m for m in self._module.modules() if isinstance(m, Embedding)]
    anchors = {id(t._anchor) for t in self._tables}
    self._dense = [p for p in self._module.parameters() if p.requires_grad and id(p) not in anchors]

  def _accum(self, owner, like: torch.Tensor) -> torch.Tensor:
    """Accumulator slot of a variable, stored ON its owner object (an Embedding module or a Parameter) so it
    lives and dies with it -- an id()-keyed dict would hand a recycled id the previous owner's state."""
    a = getattr(owner, "_tfrs_adagrad_acc", None)
    if a is None or a.shape != like.shape or a.device != like.device:
      a = torch.full_like(like, self.initial_accumulator_value)
      owner._tfrs_adagrad_acc = a
      self._acc[id(owner)] = a
    return a

  def zero_grad(self):
    self._refresh()
    for p in self._dense:
      p.grad = None
    for t in self._tables:
      t.pop_sparse_grads()
      t._anchor.grad = None

  @torch.no_grad()
  def apply_gradients(self):
    """optimizer.apply_gradients(zip(grads, vars)) -- models/base.py:78."""
    self._refresh()
    for t in self._tables:
      grads = t.pop_sparse_grads()
      if not grads:
        continue
      ids = torch.cat([i.reshape(-1) for i, _ in grads], 0)
      rows = torch.cat([g.reshape(-1, t.output_dim) for _, g in grads], 0)
      ops.sparse_adagrad_(t.weight, self._accum(t, t.weight), ids, rows, self.learning_rate, self.epsilon,
                          self.eps_inside_sqrt)
    for p in self._dense:
      if p.grad is None:
        continue
      a = self._accum(p, p)
      g = p.grad
      a.addcmul_(g, g)
      den = (a + self.epsilon).sqrt_() if self.eps_inside_sqrt else a.sqrt().add_(self.epsilon)
      p.addcdiv_(g, den, value=-self.learning_rate)

  step = apply_gradients

  def state_dict(self):
    return {"acc": {k: v.clone() for k, v in self._acc.items()}}
