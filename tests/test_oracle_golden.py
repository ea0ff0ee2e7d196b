"""Pins the CPU oracle against the reference's own known-answer tests.

Each test restates one reference test (file:line in the docstring); TensorFlow is replaced by the
oracle, the expectation is the reference's NumPy expectation verbatim."""
import itertools

import numpy as np
import pytest

from oracle import oracle as orc


def _cases():
  # layers/factorized_top_k_test.py:31-66 -- the 96-way product
  return list(itertools.product((5, 10), (3, 16), (3, 15, 16), (1024, 128), (str, None), (True, False)))


def _batches(arr, b):
  return [arr[i:i + b] for i in range(0, len(arr), b)]


@pytest.mark.parametrize("layer", ["streaming", "brute_force"])
@pytest.mark.parametrize("k,batch_size,num_queries,num_candidates,indices_dtype,use_exclusions", _cases())
def test_top_k_layers(layer, k, batch_size, num_queries, num_candidates, indices_dtype, use_exclusions):
  """layers/factorized_top_k_test.py:85-147 (run_top_k_test), Streaming :168-173, BruteForce :176-180."""
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

  def fn(q, kk):
    if layer == "brute_force":
      ids = candidate_indices if indices_dtype is not None else None
      return orc.brute_force(q, candidates, ids, kk)
    if indices_dtype is not None:
      chunks = list(zip(_batches(candidate_indices, batch_size), _batches(candidates, batch_size)))
    else:
      chunks = _batches(candidates, batch_size)
    return orc.streaming(q, chunks, kk)

  for _ in range(2):  # repeatability, :132-140
    if use_exclusions:
      top_s, top_i = orc.query_with_exclusions(fn, query, exclude_identifiers, k)
    else:
      top_s, top_i = fn(query, k)
  assert top_s.shape == expected_scores.shape and top_i.shape == expected_ids.shape
  np.testing.assert_allclose(top_s, expected_scores, atol=1e-4)
  np.testing.assert_array_equal(top_i.astype(expected_ids.dtype), expected_ids)


def test_raise_on_incorrect_input_shape():
  """layers/factorized_top_k_test.py:229-243 -- 100 candidates vs 99 identifiers."""
  c = np.random.normal(size=(100, 4)).astype(np.float32)
  with pytest.raises(ValueError):
    orc.brute_force(c[:2], c, np.arange(99), 5)


@pytest.mark.parametrize("layer", ["streaming", "brute_force"])
@pytest.mark.parametrize("use_ids", [True, False])
def test_factorized_top_k_metric(layer, use_ids):
  """metrics/factorized_top_k_test.py:39-86."""
  rng = np.random.RandomState(42)
  N, Q, d = 100, 10, 4
  candidate_ids = np.arange(0, N).astype(str)
  candidates = rng.normal(size=(N, d)).astype(np.float32)
  query = rng.normal(size=(Q, d)).astype(np.float32)
  w = rng.uniform(size=(Q, 1)).astype(np.float32)
  true_idx = rng.randint(0, N, size=Q)
  ks = [1, 5, 10, 50]
  cs = query @ candidates.T

  def fn(q, kk):
    if layer == "brute_force":
      return orc.brute_force(q, candidates, candidate_ids, kk)
    return orc.streaming(q, list(zip(_batches(candidate_ids, 32), _batches(candidates, 32))), kk)

  res = orc.factorized_top_k_update(query, candidates[true_idx], fn, ks,
                                    true_ids=candidate_ids[true_idx] if use_ids else None, sample_weight=w)
  for k, (num, den) in zip(ks, res):
    expected = np.average(orc.in_top_k(true_idx, cs, k).astype(np.float32), weights=w[:, 0])
    np.testing.assert_allclose(num / den, expected, rtol=1e-6)


def test_id_based_evaluation():
  """metrics/factorized_top_k_test.py:93-131 -- N=1280, d=128, K=100, chunks of 32."""
  rng = np.random.default_rng(42)
  k, N, Q, d = 100, 1280, 128, 128
  cand = rng.normal(size=(N, d)).astype(np.float32)
  qs = rng.normal(size=(Q, d)).astype(np.float32)
  true_idx = rng.integers(0, N, size=Q).astype(np.int32)
  fn = lambda q, kk: orc.streaming(q, _batches(cand, 32), kk)
  num = den = 0.0
  hits = 0
  for q, t in zip(qs, true_idx):
    (n, d_), = orc.factorized_top_k_update(q.reshape(1, -1), cand[t].reshape(1, -1), fn, [k], true_ids=np.array([t]))
    num += n; den += d_
    _, ti = fn(q.reshape(1, -1), k)
    hits += int(t in ti[0].tolist())
  assert num / den == hits / Q
  # second brute-force oracle (examples/movielens.py:74-93): argsort(-scores)[:k]
  ref = np.argsort(-(qs @ cand.T), axis=1, kind="stable")[:, :k]
  _, got = orc.brute_force(qs, cand, None, k)
  assert (got == ref).mean() > 0.999  # float32-vs-BLAS order may flip sub-ulp near-ties only


def _sigmoid(x):
  return 1.0 / (1 + np.exp(-x))


def test_retrieval_loss_2x2():
  """tasks/retrieval_test.py:31-137."""
  q = np.array([[1, 2, 3], [2, 3, 4]], np.float32)
  c = np.array([[1, 1, 1], [1, 1, 0]], np.float32)
  s, y = orc.retrieval_scores(q, c)
  np.testing.assert_array_equal(s, [[6, 3], [9, 5]])
  expected = -np.log(_sigmoid(3.0)) - np.log(1 - _sigmoid(4.0))
  np.testing.assert_allclose(orc.retrieval_loss(q, c), expected, rtol=1e-6)
  expected3 = -0.7 * np.log(_sigmoid(3.0)) - 0.3 * np.log(1 - _sigmoid(4.0))
  np.testing.assert_allclose(orc.retrieval_loss(q, c, sample_weight=[0.7, 0.3]), expected3, rtol=1e-6)
  # corpus of 20 zero rows in chunks of 16, ks=[5] -> accuracy 1.0 (:36-37,60)
  corpus = np.zeros((20, 3), np.float32)
  fn = lambda qq, kk: orc.streaming(qq, _batches(corpus, 16), kk)
  (n, d), = orc.factorized_top_k_update(q, c, fn, [5])
  assert n / d == 1.0


def test_retrieval_extra_negatives():
  """tasks/retrieval_test.py:179-213."""
  q = np.array([[3, 2, 1], [2, 3, 4]], np.float32)
  c = np.array([[0, 1, 0], [0, 1, 1], [1, 1, 0]], np.float32)
  s, _ = orc.retrieval_scores(q, c)
  np.testing.assert_array_equal(s, [[2, 3, 5], [3, 7, 5]])
  expected = (-np.log(1 / (1 + np.exp(1) + np.exp(3))) - np.log(np.exp(4) / (1 + np.exp(4) + np.exp(2))))
  np.testing.assert_allclose(orc.retrieval_loss(q, c), expected, rtol=1e-6)


def test_retrieval_multipoint():
  """tasks/retrieval_test.py:255-298."""
  q = np.array([[[3, 2, 1], [1, 2, 3]], [[2, 3, 4], [4, 3, 2]]], np.float32)
  c = np.array([[0, 1, 0], [0, 1, 1], [1, 1, 0]], np.float32)
  s, _ = orc.retrieval_scores(q, c)
  np.testing.assert_array_equal(s, [[2, 5, 5], [3, 7, 7]])
  expected = -np.log(1 / (1 + np.exp(3) + np.exp(3))) - np.log(np.exp(5) / (np.exp(1) + np.exp(5) + np.exp(5)))
  np.testing.assert_allclose(orc.retrieval_loss(q, c), expected, rtol=1e-6)


@pytest.mark.parametrize("seed", [42, 123, 8391, 12390, 1230])
def test_loss_layers(seed):
  """layers/loss_test.py:29-130."""
  rng = np.random.RandomState(seed)
  shape = (2, 20)
  logits = rng.uniform(size=shape).astype(np.float32)
  labels = rng.permutation(np.eye(*shape).T).T.astype(np.float32)
  ol, oy = orc.hard_negative_mining(logits, labels, 3)
  assert ol.shape[-1] == 4
  np.testing.assert_allclose((ol * oy).sum(1), (logits * labels).sum(1))
  l2 = logits + labels * 1000.0
  ol, oy = orc.hard_negative_mining(l2, labels, 3)
  np.testing.assert_allclose(np.sort(l2, axis=1)[:, -4:], np.sort(ol))

  rng = np.random.RandomState(seed)
  shape = (2, 4)
  logits = rng.uniform(size=shape).astype(np.float32)
  labels = rng.permutation(np.eye(*shape).T).T.astype(np.float32)
  cid = rng.randint(0, 3, size=shape[-1])
  out = orc.remove_accidental_hits(labels, logits, cid)
  np.testing.assert_allclose((out * labels).sum(1), (logits * labels).sum(1))
  for r in range(shape[0]):
    p = np.argmax(labels[r])
    for col in range(shape[1]):
      if cid[p] == cid[col] and col != p:
        np.testing.assert_allclose(out[r, col], logits[r, col] + orc.MIN_FLOAT)
      else:
        np.testing.assert_allclose(out[r, col], logits[r, col])

  rng = np.random.RandomState(seed)
  shape = (10, 20)
  logits = rng.uniform(size=shape).astype(np.float32)
  probs = rng.uniform(size=shape[1]).astype(np.float32)
  assert (logits < orc.sampling_probability_correction(logits, probs)).all()
  pz = probs * rng.choice([0.0, 1.0], size=probs.shape)
  assert (logits < orc.sampling_probability_correction(logits, pz)).all()


def test_cross_known_answers():
  """layers/feature_interaction/dcn_test.py:29-101."""
  x0 = np.array([[0.1, 0.2, 0.3]], np.float32)
  x = np.array([[0.4, 0.5, 0.6]], np.float32)
  ones = np.ones((3, 3), np.float32)
  np.testing.assert_allclose(orc.cross(x0, x, ones), [[0.55, 0.8, 1.05]], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, x, None, U=np.ones((3, 1)), V=np.ones((1, 3))), [[0.55, 0.8, 1.05]], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, None, ones), [[0.16, 0.32, 0.48]], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, x, ones, bias=np.ones(3)), [[0.65, 1.0, 1.35]], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, x, ones, diag_scale=1.0), [[0.59, 0.9, 1.23]], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, x, ones, preactivation=np.zeros_like), x, rtol=1e-6)
  with pytest.raises(ValueError, match="dimension mismatch"):
    orc.cross(np.random.random((12, 5)), np.random.random((12, 7)), np.ones((7, 7)))


def test_sparse_adagrad_dedupes():
  """Keras sparse Adagrad: duplicate rows are summed before the square (unpinned)."""
  t = np.ones((4, 2), np.float32); a = np.full((4, 2), 0.1, np.float32)
  ids = np.array([1, 3, 1]); g = np.array([[1, 2], [3, 4], [5, 6]], np.float32)
  t2, a2 = orc.sparse_adagrad(t, a, ids, g, lr=0.5, eps=1e-7)
  gs = np.array([[6, 8]], np.float32)
  np.testing.assert_allclose(a2[1], 0.1 + gs[0] ** 2, rtol=1e-6)
  np.testing.assert_allclose(t2[1], 1 - 0.5 * gs[0] / np.sqrt(a2[1] + 1e-7), rtol=1e-6)
  np.testing.assert_array_equal(t2[[0, 2]], t[[0, 2]])


def test_merge_matches_scan():
  rng = np.random.RandomState(0)
  q = rng.normal(size=(7, 8)).astype(np.float32); c = rng.normal(size=(300, 8)).astype(np.float32)
  full_s, full_i = orc.topk_scan(q, c, 10)
  parts = [orc.topk_scan(q, c[o:o + 100], 10, index_offset=o) for o in (0, 100, 200)]
  ms, mi = orc.topk_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 10)
  np.testing.assert_array_equal(ms, full_s); np.testing.assert_array_equal(mi, full_i)


def test_tie_rule_lowest_index_first():
  q = np.ones((1, 4), np.float32); c = np.ones((9, 4), np.float32)
  _, i = orc.topk_scan(q, c, 4)
  np.testing.assert_array_equal(i, [[0, 1, 2, 3]])


def test_multi_layer_dcn_known_answers():
  """multi_layer_dcn_test.py:28-59 (kernel_initializer="ones")."""
  x0 = np.asarray([[0.1, 0.2, 0.3]], np.float32)
  np.testing.assert_allclose(orc.multi_layer_dcn(x0, [np.ones((3, 3))], [np.ones((3, 3))]), [[0.28, 0.56, 0.84]], rtol=1e-6)
  np.testing.assert_allclose(orc.multi_layer_dcn(x0, [np.ones((3, 1))], [np.ones((1, 3))]), [[0.16, 0.32, 0.48]], rtol=1e-6)
  ones_u, ones_v, ones_b = [np.ones((3, 1))] * 3, [np.ones((1, 3))] * 3, [np.ones(3)] * 3
  np.testing.assert_allclose(orc.multi_layer_dcn(x0, ones_u, ones_v, ones_b), [[0.9256, 1.8512, 2.7768]], rtol=1e-5)


def test_dot_interaction_known_answers():
  """dot_interaction_test.py:27-65."""
  f1 = np.asarray([[0.1, -4.3, 0.2, 1.1, 0.3]], np.float32)
  f2 = np.asarray([[2.0, 3.2, -1.0, 0.0, 1.0]], np.float32)
  f3 = np.asarray([[0.0, 1.0, -3.0, -2.2, -0.2]], np.float32)
  d = lambda a, b: np.dot(a[0], b[0])
  f11, f12, f13, f22, f23, f33 = d(f1, f1), d(f1, f2), d(f1, f3), d(f2, f2), d(f2, f3), d(f3, f3)
  np.testing.assert_allclose(orc.dot_interaction([f1, f2, f3], True, False), [[f11, f12, f22, f13, f23, f33]], rtol=1e-6)
  np.testing.assert_allclose(orc.dot_interaction([f1, f2, f3], True, True), [[f11, 0, 0, f12, f22, 0, f13, f23, f33]], rtol=1e-6)
  np.testing.assert_allclose(orc.dot_interaction([f1, f2, f3], False, False), [[f12, f13, f23]], rtol=1e-6)
  np.testing.assert_allclose(orc.dot_interaction([f1, f2, f3], False, True), [[0, 0, 0, f12, 0, 0, f13, f23, 0]], rtol=1e-6)
  with pytest.raises(ValueError, match="dimensions must be equal"):
    orc.dot_interaction([np.zeros((1, 3), np.float32), np.zeros((1, 3), np.float32), np.zeros((1, 2), np.float32)])
