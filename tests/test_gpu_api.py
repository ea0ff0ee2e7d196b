"""GPU tests of the reference-facing API (recommenders_b200 as tfrs): the reference's own tests restated
with torch tensors (file:line per test).  Run with -m gpu."""
import itertools

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def cu(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def tfrs():
  import recommenders_b200 as t
  return t


def _cases():
  return list(itertools.product((5, 10), (3, 16), (3, 15, 16), (1024, 128), (str, None), (True, False)))


@pytest.mark.parametrize("layer_name", ["Streaming", "BruteForce"])
@pytest.mark.parametrize("k,batch_size,num_queries,num_candidates,indices_dtype,use_exclusions", _cases())
def test_top_k_layers(tfrs, layer_name, k, batch_size, num_queries, num_candidates, indices_dtype, use_exclusions):
  """layers/factorized_top_k_test.py:85-147 (+ save/restore :152-165 for BruteForce)."""
  Dataset = tfrs.data.Dataset
  layer = getattr(tfrs.layers.factorized_top_k, layer_name)(k=k)
  rng = np.random.RandomState(42)
  candidates = rng.normal(size=(num_candidates, 4)).astype(np.float32)
  query = rng.normal(size=(num_queries, 4)).astype(np.float32)
  candidate_indices = np.arange(num_candidates).astype(indices_dtype if indices_dtype is not None else np.int32)
  exclude = rng.randint(0, num_candidates, size=(num_queries, 5))
  scores = np.dot(query, candidates.T)
  adjusted = scores.copy()
  exclude_identifiers = None
  if use_exclusions:
    exclude_identifiers = candidate_indices[exclude]
    for r, row in enumerate(exclude):
      for c in set(row):
        adjusted[r, c] -= 1000.0
  indices = np.argsort(-adjusted, axis=1)[:, :k]
  expected_scores = np.take_along_axis(scores, indices, 1)
  expected_ids = candidate_indices[indices]

  ds = Dataset.from_tensor_slices(cu(candidates)).batch(batch_size)
  if indices_dtype is not None:
    ds = Dataset.zip((Dataset.from_tensor_slices(candidate_indices).batch(batch_size), ds))
  q = cu(query)
  excl = exclude_identifiers if (exclude_identifiers is None or indices_dtype is not None) else cu(exclude_identifiers)
  for _ in range(2):
    layer.index_from_dataset(ds)
    if use_exclusions:
      top_s, top_i = layer.query_with_exclusions(q, excl)
    else:
      top_s, top_i = layer(q)
  top_i = top_i.cpu().numpy() if isinstance(top_i, torch.Tensor) else top_i
  assert top_s.shape == expected_scores.shape and top_i.shape == expected_ids.shape
  np.testing.assert_allclose(top_s.cpu().numpy(), expected_scores, atol=1e-4)
  np.testing.assert_array_equal(top_i.astype(expected_ids.dtype), expected_ids)

  if layer_name == "BruteForce":
    restored = tfrs.layers.factorized_top_k.BruteForce(k=k)
    restored.load_state_dict(layer.state_dict())
    if use_exclusions:
      _, ri = restored.query_with_exclusions(q, excl)
    else:
      _, ri = restored(q)
    ri = ri.cpu().numpy() if isinstance(ri, torch.Tensor) else ri
    np.testing.assert_array_equal(ri.astype(expected_ids.dtype), expected_ids)


def test_layer_errors(tfrs):
  """factorized_top_k_test.py:229-243 and the error contract of SURVEY.md 8b."""
  ftk = tfrs.layers.factorized_top_k
  Dataset = tfrs.data.Dataset
  cands = cu(np.random.normal(size=(100, 4)).astype(np.float32))
  with pytest.raises(ValueError):
    ftk.BruteForce().index_from_dataset(Dataset.zip((Dataset.from_tensor_slices(np.arange(99)).batch(20),
                                                     Dataset.from_tensor_slices(cands).batch(100))))
  with pytest.raises(ValueError):
    ftk.BruteForce().index(cands, np.arange(99))
  with pytest.raises(ValueError):
    ftk.BruteForce().index(cands.reshape(-1))
  with pytest.raises(ValueError):
    ftk.BruteForce()(cands[:2])
  with pytest.raises(ValueError):
    ftk.Streaming()(cands[:2])
  with pytest.raises(NotImplementedError):
    ftk.Streaming().index(cands)
  with pytest.raises(ValueError, match="batch size is too small"):
    ftk.Streaming(k=10, handle_incomplete_batches=False).index_from_dataset(
        Dataset.from_tensor_slices(cands).batch(8))(cands[:2])
  with pytest.raises(ImportError):
    ftk.ScaNN()


@pytest.mark.parametrize("top_k_layer", ["Streaming", "BruteForce", None])
@pytest.mark.parametrize("use_candidate_ids", [True, False])
def test_factorized_top_k_metric(tfrs, top_k_layer, use_candidate_ids):
  """metrics/factorized_top_k_test.py:39-86."""
  rng = np.random.RandomState(42)
  N, Q, d = 100, 10, 4
  candidate_ids = np.arange(0, N).astype(str)
  candidates = rng.normal(size=(N, d)).astype(np.float32)
  query = rng.normal(size=(Q, d)).astype(np.float32)
  sample_weight = rng.uniform(size=(Q, 1)).astype(np.float32)
  true_idx = rng.randint(0, N, size=Q)
  cs = query @ candidates.T
  ks = [1, 5, 10, 50]
  ds = tfrs.data.Dataset.from_tensor_slices((candidate_ids, cu(candidates))).batch(32)
  cand = ds if top_k_layer is None else getattr(tfrs.layers.factorized_top_k, top_k_layer)().index_from_dataset(ds)
  metric = tfrs.metrics.FactorizedTopK(candidates=cand, ks=ks)
  metric.update_state(query_embeddings=cu(query), true_candidate_embeddings=cu(candidates[true_idx]),
                      true_candidate_ids=candidate_ids[true_idx] if use_candidate_ids else None,
                      sample_weight=cu(sample_weight))
  for k, val in zip(ks, metric.result()):
    exp = np.average(orc.in_top_k(true_idx, cs, k).astype(np.float32), weights=sample_weight[:, 0])
    np.testing.assert_allclose(val, exp, rtol=1e-5)


@pytest.mark.parametrize("layer", ["Streaming", "BruteForce"])
def test_id_based_evaluation(tfrs, layer):
  """metrics/factorized_top_k_test.py:93-131."""
  rng = np.random.default_rng(42)
  k, N, Q, d = 100, 1280, 128, 128
  cand = rng.normal(size=(N, d)).astype(np.float32)
  qs = rng.normal(size=(Q, d)).astype(np.float32)
  true_idx = rng.integers(0, N, size=Q).astype(np.int32)
  index = getattr(tfrs.layers.factorized_top_k, layer)(k=k).index_from_dataset(
      tfrs.data.Dataset.from_tensor_slices(cu(cand)).batch(32))
  metric = tfrs.metrics.FactorizedTopK(candidates=index, ks=[k])
  hits = 0
  tq, tc = cu(qs), cu(cand)
  for i in range(Q):
    metric.update_state(tq[i:i + 1], tc[int(true_idx[i])].reshape(1, -1), cu(true_idx[i:i + 1]))
    _, ti = index(tq[i:i + 1])
    hits += int(int(true_idx[i]) in ti[0].cpu().tolist())
  assert metric.result()[0] == hits / Q
  # all 128 queries at once agree with the oracle bit-for-bit
  _, ti = index(tq)
  _, ei = orc.topk_scan(qs, cand, k)
  np.testing.assert_array_equal(ti.cpu().numpy(), ei)


def _sigmoid(x):
  return 1.0 / (1 + np.exp(-x))


def test_retrieval_task(tfrs):
  """tasks/retrieval_test.py:31-137."""
  query = cu(np.array([[1, 2, 3], [2, 3, 4]], np.float32))
  candidate = cu(np.array([[1, 1, 1], [1, 1, 0]], np.float32))
  ds = tfrs.data.Dataset.from_tensor_slices(cu(np.zeros((20, 3), np.float32))).batch(16)
  task = tfrs.tasks.Retrieval(
      metrics=tfrs.metrics.FactorizedTopK(candidates=ds, ks=[5]),
      batch_metrics=[tfrs.metrics.TopKCategoricalAccuracy(k=1, name="batch_categorical_accuracy_at_1")],
      loss_metrics=[tfrs.metrics.Mean(name="batch_loss")])
  expected_loss = -np.log(_sigmoid(3.0)) - np.log(1 - _sigmoid(4.0))

  def run(**kw):
    for m in task.metrics:
      m.reset_states()
    loss = task(query_embeddings=query, candidate_embeddings=candidate, **kw)
    return float(loss), {m.name: m.result() for m in task.metrics}

  loss, m = run()
  np.testing.assert_allclose(loss, expected_loss, rtol=1e-5)
  np.testing.assert_allclose([m["factorized_top_k/top_5_categorical_accuracy"], m["batch_categorical_accuracy_at_1"],
                              m["batch_loss"]], [1.0, 0.5, expected_loss], rtol=1e-5)
  loss, m = run(compute_metrics=False)
  np.testing.assert_allclose([loss, m["factorized_top_k/top_5_categorical_accuracy"], m["batch_categorical_accuracy_at_1"]],
                             [expected_loss, 0.0, 0.5], rtol=1e-5)
  loss, m = run(compute_batch_metrics=False)
  np.testing.assert_allclose([loss, m["factorized_top_k/top_5_categorical_accuracy"], m["batch_categorical_accuracy_at_1"]],
                             [expected_loss, 1.0, 0.0], rtol=1e-5)
  expected3 = -0.7 * np.log(_sigmoid(3.0)) - 0.3 * np.log(1 - _sigmoid(4.0))
  loss, m = run(sample_weight=cu(np.array([0.7, 0.3], np.float32)))
  np.testing.assert_allclose([loss, m["factorized_top_k/top_5_categorical_accuracy"], m["batch_categorical_accuracy_at_1"],
                              m["batch_loss"]], [expected3, 1.0, 0.7, expected3], rtol=1e-5)


def test_retrieval_extra_negatives_and_multipoint(tfrs):
  """tasks/retrieval_test.py:179-213 and :255-298."""
  c = cu(np.array([[0, 1, 0], [0, 1, 1], [1, 1, 0]], np.float32))
  ds = tfrs.data.Dataset.from_tensor_slices(cu(np.zeros((20, 3), np.float32))).batch(16)
  mk = lambda: tfrs.tasks.Retrieval(metrics=tfrs.metrics.FactorizedTopK(candidates=ds, ks=[5]),
                                    batch_metrics=[tfrs.metrics.TopKCategoricalAccuracy(k=1, name="acc1")])
  task = mk()
  q = cu(np.array([[3, 2, 1], [2, 3, 4]], np.float32))
  exp = (-np.log(1 / (1 + np.exp(1) + np.exp(3))) - np.log(np.exp(4) / (1 + np.exp(4) + np.exp(2))))
  np.testing.assert_allclose(float(task(q, c)), exp, rtol=1e-5)
  m = {x.name: x.result() for x in task.metrics}
  assert m["factorized_top_k/top_5_categorical_accuracy"] == 1.0 and m["acc1"] == 0.5
  task = mk()
  q3 = cu(np.array([[[3, 2, 1], [1, 2, 3]], [[2, 3, 4], [4, 3, 2]]], np.float32))
  exp = -np.log(1 / (1 + np.exp(3) + np.exp(3))) - np.log(np.exp(5) / (np.exp(1) + np.exp(5) + np.exp(5)))
  np.testing.assert_allclose(float(task(q3, c)), exp, rtol=1e-5)
  m = {x.name: x.result() for x in task.metrics}
  assert m["factorized_top_k/top_5_categorical_accuracy"] == 0.0 and m["acc1"] == 0.5


def test_retrieval_options_match_oracle(tfrs):
  """temperature / sampling probability / accidental hits / score mask / hard negatives (retrieval.py:187-208)."""
  rng = np.random.RandomState(0)
  q = rng.normal(size=(12, 8)).astype(np.float32); c = rng.normal(size=(20, 8)).astype(np.float32)
  ids = rng.randint(0, 6, size=20); prob = rng.uniform(0.01, 1, size=20).astype(np.float32)
  mask = rng.uniform(size=(12, 20)) > 0.2
  mask[np.arange(12), np.arange(12)] = True
  w = rng.uniform(size=12).astype(np.float32)
  for kw_t, kw_o in [
      (dict(temperature=0.3), dict(temperature=0.3)),
      (dict(temperature=2.0, remove_accidental_hits=True), dict(temperature=2.0, remove_accidental_hits_=True, candidate_ids=ids)),
      (dict(num_hard_negatives=4), dict(num_hard_negatives=4)),
  ]:
    task = tfrs.tasks.Retrieval(**kw_t)
    got = float(task(cu(q), cu(c), sample_weight=cu(w), candidate_ids=ids if "remove_accidental_hits" in kw_t else None,
                     compute_metrics=False))
    exp = orc.retrieval_loss(q, c, sample_weight=w, **kw_o)
    np.testing.assert_allclose(got, exp, rtol=2e-5)
  task = tfrs.tasks.Retrieval()
  got = float(task(cu(q), cu(c), candidate_sampling_probability=cu(prob), score_mask=cu(mask), compute_metrics=False))
  exp = orc.retrieval_loss(q, c, candidate_sampling_probability=prob, score_mask=mask)
  np.testing.assert_allclose(got, exp, rtol=2e-5)
  with pytest.raises(ValueError):
    tfrs.tasks.Retrieval(remove_accidental_hits=True)(cu(q), cu(c))


def test_cross_known_answers(tfrs):
  """layers/feature_interaction/dcn_test.py:29-101."""
  Cross = tfrs.layers.dcn.Cross
  x0 = cu(np.array([[0.1, 0.2, 0.3]], np.float32)); x = cu(np.array([[0.4, 0.5, 0.6]], np.float32))
  close = lambda a, b: np.testing.assert_allclose(a.detach().cpu().numpy(), b, rtol=1e-5)
  close(Cross(projection_dim=None, kernel_initializer="ones")(x0, x), [[0.55, 0.8, 1.05]])
  close(Cross(projection_dim=1, kernel_initializer="ones")(x0, x), [[0.55, 0.8, 1.05]])
  close(Cross(projection_dim=None, kernel_initializer="ones")(x0), [[0.16, 0.32, 0.48]])
  close(Cross(projection_dim=None, kernel_initializer="ones", bias_initializer="ones")(x0, x), [[0.65, 1.0, 1.35]])
  close(Cross(projection_dim=None, diag_scale=1.0, kernel_initializer="ones")(x0, x), [[0.59, 0.9, 1.23]])
  close(Cross(projection_dim=None, preactivation=torch.zeros_like)(x0, x), x.cpu().numpy())
  with pytest.raises(ValueError, match="dimension mismatch"):
    Cross()(cu(np.random.random((12, 5)).astype(np.float32)), cu(np.random.random((12, 7)).astype(np.float32)))
  with pytest.raises(ValueError, match="`diag_scale` should be non-negative"):
    Cross(diag_scale=-1.0)
  layer = Cross(projection_dim=None, preactivation="swish")
  assert Cross.from_config(layer.get_config()).get_config() == layer.get_config()
  # state_dict round trip of a 2-layer stack (dcn_test.py:103-126)
  a, b = Cross(), Cross()
  xin = cu(np.random.uniform(size=(10, 13)).astype(np.float32))
  ref = b(xin, a(xin, xin))
  a2, b2 = Cross(), Cross()
  a2.build(xin.shape); b2.build(xin.shape)
  a2.load_state_dict(a.state_dict()); b2.load_state_dict(b.state_dict())
  assert torch.equal(b2(xin, a2(xin, xin)), ref)


def test_two_tower_model_trains(tfrs):
  """README.md:44-98 shaped end-to-end: Embedding towers -> Retrieval -> sparse Adagrad; the CUDA step must
  track the oracle step by step (config 1: MovieLens-100K-shaped 2k x 2k x 64)."""
  torch.manual_seed(0)
  rng = np.random.RandomState(42)
  U, I, d, B = 2000, 2000, 64, 4096

  class TwoTower(tfrs.Model):

    def __init__(self):
      super().__init__()
      self.user_model = tfrs.layers.embedding.Embedding(U, d)
      self.item_model = tfrs.layers.embedding.Embedding(I, d)
      self.task = tfrs.tasks.Retrieval()

    def compute_loss(self, features, training=False):
      return self.task(self.user_model(features["user_id"]), self.item_model(features["movie_id"]),
                       compute_metrics=not training)

  model = TwoTower()
  model.compile(optimizer=tfrs.optimizers.Adagrad(0.5))
  ut = model.user_model.weight.cpu().numpy().copy(); it = model.item_model.weight.cpu().numpy().copy()
  ua = np.full_like(ut, 0.1); ia = np.full_like(it, 0.1)
  losses = []
  uid = rng.randint(0, U, size=B).astype(np.int64); iid = rng.randint(0, I, size=B).astype(np.int64)
  for step in range(3):  # the same batch three times: the loss must fall
    out = model.train_step({"user_id": cu(uid), "movie_id": cu(iid)})
    assert set(out) >= {"loss", "regularization_loss", "total_loss"}
    losses.append(float(out["loss"]))
    # oracle step
    qe, ce = orc.gather(ut, uid), orc.gather(it, iid)
    exp_loss = orc.retrieval_loss(qe, ce)
    dq, dc = orc.retrieval_loss_grads(qe, ce)
    np.testing.assert_allclose(losses[-1], exp_loss, rtol=1e-5)
    ut, ua = orc.sparse_adagrad(ut, ua, uid, dq.astype(np.float32), 0.5)
    it, ia = orc.sparse_adagrad(it, ia, iid, dc.astype(np.float32), 0.5)
    np.testing.assert_allclose(model.user_model.weight.cpu().numpy(), ut, rtol=1e-4, atol=2e-5)
  assert losses[-1] < losses[0]
  ev = model.test_step({"user_id": cu(uid), "movie_id": cu(iid)})
  assert "loss" in ev
  with pytest.raises(NotImplementedError):
    tfrs.Model().compute_loss(None)


def test_streaming_carried_state_across_scans(tfrs):
  """Streaming.call (factorized_top_k.py:404-509): many small batches, several coalesced scans, carried [Q,k] state;
  the result must equal one brute-force scan, with and without identifiers."""
  rng = np.random.RandomState(3)
  cand = rng.normal(size=(5000, 16)).astype(np.float32); q = rng.normal(size=(33, 16)).astype(np.float32)
  es, ei = orc.topk_scan(q, cand, 50)
  ds = tfrs.data.Dataset.from_tensor_slices(cu(cand)).batch(37)
  layer = tfrs.layers.factorized_top_k.Streaming(k=50).index_from_dataset(ds)
  layer._coalesce_rows = 300   # ~17 scans with carried state
  s, i = layer(cu(q))
  np.testing.assert_array_equal(i.cpu().numpy(), ei.astype(np.int32)); np.testing.assert_array_equal(s.cpu().numpy(), es)
  ids = (np.arange(5000) * 7 + 3).astype(np.int64)
  ds2 = tfrs.data.Dataset.from_tensor_slices((cu(ids), cu(cand))).batch(41)
  layer2 = tfrs.layers.factorized_top_k.Streaming(k=50).index_from_dataset(ds2)
  layer2._coalesce_rows = 1
  s2, i2 = layer2(cu(q))
  np.testing.assert_array_equal(i2.cpu().numpy(), ids[ei]); np.testing.assert_array_equal(s2.cpu().numpy(), es)


def test_dot_interaction_layer_known_answers_and_parity():
  """dot_interaction_test.py:27-72 through the layer, then bit-exact parity (canonical fmaf chain) and gradients vs the
  oracle on random data, all four (self_interaction, skip_gather) modes, cfg5-like shape."""
  import recommenders_b200 as tfrs
  DotInteraction = tfrs.layers.feature_interaction.DotInteraction
  f1 = np.asarray([[0.1, -4.3, 0.2, 1.1, 0.3]], np.float32)
  f2 = np.asarray([[2.0, 3.2, -1.0, 0.0, 1.0]], np.float32)
  f3 = np.asarray([[0.0, 1.0, -3.0, -2.2, -0.2]], np.float32)
  for si in (True, False):
    for sg in (True, False):
      exp = orc.dot_interaction([f1, f2, f3], si, sg)
      got = DotInteraction(self_interaction=si, skip_gather=sg)([cu(f1), cu(f2), cu(f3)])
      np.testing.assert_array_equal(got.cpu().numpy(), exp)
  with pytest.raises(ValueError, match="dimensions must be equal"):
    DotInteraction()([cu(np.zeros((1, 3), np.float32)), cu(np.zeros((1, 3), np.float32)), cu(np.zeros((1, 2), np.float32))])
  rng = np.random.RandomState(3)
  for (B, F, d) in ((37, 27, 32), (5, 1, 8), (130, 8, 17)):
    feats = [rng.normal(size=(B, d)).astype(np.float32) for _ in range(F)]
    for si in (True, False):
      for sg in (True, False):
        exp = orc.dot_interaction(feats, si, sg)
        ts = [cu(f).requires_grad_(True) for f in feats]
        got = DotInteraction(self_interaction=si, skip_gather=sg)(ts)
        np.testing.assert_array_equal(got.detach().cpu().numpy(), exp)
        if exp.shape[1] == 0:
          continue
        g = rng.normal(size=exp.shape).astype(np.float32)
        got.backward(cu(g))
        edf = orc.dot_interaction_grads(feats, g, si, sg)
        for f in range(F):
          err = np.abs(ts[f].grad.cpu().numpy().astype(np.float64) - edf[:, f, :]).max()
          assert err <= 1e-5 * np.abs(edf).max()


def test_multi_layer_dcn_known_answers_and_parity():
  """multi_layer_dcn_test.py:28-59, then parity with the float64 oracle on random weights (1e-5)."""
  import recommenders_b200 as tfrs
  MultiLayerDCN = tfrs.layers.feature_interaction.MultiLayerDCN
  x0 = cu(np.asarray([[0.1, 0.2, 0.3]], np.float32))
  out = MultiLayerDCN(projection_dim=3, num_layers=1, use_bias=False, kernel_initializer="ones")(x0)
  np.testing.assert_allclose(out.detach().cpu().numpy(), [[0.28, 0.56, 0.84]], rtol=1e-6)
  out = MultiLayerDCN(projection_dim=1, num_layers=1, use_bias=False, kernel_initializer="ones")(x0)
  np.testing.assert_allclose(out.detach().cpu().numpy(), [[0.16, 0.32, 0.48]], rtol=1e-6)
  out = MultiLayerDCN(projection_dim=1, kernel_initializer="ones", bias_initializer="ones")(x0)
  np.testing.assert_allclose(out.detach().cpu().numpy(), [[0.9256, 1.8512, 2.7768]], rtol=1e-5)
  layer = MultiLayerDCN(projection_dim=1)
  assert MultiLayerDCN.from_config(layer.get_config()).get_config() == layer.get_config()
  rng = np.random.RandomState(9)
  x = rng.uniform(size=(300, 40)).astype(np.float32)
  layer = MultiLayerDCN(projection_dim=10, num_layers=3)
  y = layer(cu(x))
  exp = orc.multi_layer_dcn(x, [u.detach().cpu().numpy() for u in layer.u_kernels], [v.detach().cpu().numpy() for v in layer.v_kernels],
                            [b.detach().cpu().numpy() for b in layer.biases])
  np.testing.assert_allclose(y.detach().cpu().numpy(), exp, rtol=1e-5, atol=1e-5 * np.abs(exp).max())
  y.sum().backward()
  assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in layer.parameters())
