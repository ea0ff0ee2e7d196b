"""The float64 loss/gradient restatement used as the checker of the fused loss kernels (oracle.retrieval_loss_and_grads_general)
is itself pinned here, on the CPU: its loss must equal the reference's op sequence (oracle.retrieval_scores ->
softmax_xent_sum, which the golden tests pin to the reference's known answers), and its gradients must equal central finite
differences of that loss -- for every option of tasks/retrieval.py:187-210."""
import numpy as np
import pytest

from oracle import oracle as orc

CASES = {
    "plain": {},
    "temperature": {"temperature": 0.7},
    "sampling_probability": {"candidate_sampling_probability": "prob"},
    "accidental_hits": {"candidate_ids": "ids", "remove_accidental_hits_": True},
    "score_mask": {"score_mask": "mask"},
    "hard_negatives": {"num_hard_negatives": 3},
    "everything_but_mining": {"temperature": 1.3, "candidate_sampling_probability": "prob", "candidate_ids": "ids",
                              "remove_accidental_hits_": True, "score_mask": "mask"},
    "mining_after_masks": {"candidate_ids": "ids", "remove_accidental_hits_": True, "score_mask": "mask", "num_hard_negatives": 2},
}


def _inputs(seed=0, B=6, C=9, d=4):
  rng = np.random.RandomState(seed)
  q = rng.normal(size=(B, d)).astype(np.float32); c = rng.normal(size=(C, d)).astype(np.float32)
  ids = rng.randint(0, 4, size=C)
  mask = rng.uniform(size=(B, C)) < 0.8
  mask[np.arange(B), np.arange(B)] = True
  prob = rng.uniform(0.05, 0.9, size=C).astype(np.float32)
  w = rng.uniform(0.5, 1.5, size=B).astype(np.float32)
  return q, c, {"ids": ids, "mask": mask, "prob": prob}, w


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("weighted", [False, True])
def test_general_loss_matches_the_reference_sequence_and_finite_differences(name, weighted):
  q, c, aux, w = _inputs()
  kw = {k: (aux[v] if isinstance(v, str) else v) for k, v in CASES[name].items()}
  sw = w if weighted else None
  loss, dq, dc = orc.retrieval_loss_and_grads_general(q, c, sw, **kw)
  ref = orc.retrieval_loss(q, c, sw, **kw)                      # float32 logits through the reference's op sequence
  assert abs(loss - ref) <= 2e-5 * max(1.0, abs(ref)), (loss, ref)

  def f(q_, c_):
    return orc.retrieval_loss_and_grads_general(q_, c_, sw, **kw)[0]
  eps = 1e-6
  q64, c64 = q.astype(np.float64), c.astype(np.float64)
  rng = np.random.RandomState(1)
  for _ in range(6):                                             # a few random coordinates of each gradient
    i, k = rng.randint(q.shape[0]), rng.randint(q.shape[1])
    qp, qm = q64.copy(), q64.copy(); qp[i, k] += eps; qm[i, k] -= eps
    assert abs((f(qp, c64) - f(qm, c64)) / (2 * eps) - dq[i, k]) <= 1e-5 * max(1.0, abs(dq[i, k]))
    j, k = rng.randint(c.shape[0]), rng.randint(c.shape[1])
    cp, cm = c64.copy(), c64.copy(); cp[j, k] += eps; cm[j, k] -= eps
    assert abs((f(q64, cp) - f(q64, cm)) / (2 * eps) - dc[j, k]) <= 1e-5 * max(1.0, abs(dc[j, k]))


def test_fully_masked_row_gives_log_C():
  """TF's fused softmax cross-entropy subtracts the row maximum first: a row whose logits are all MIN_FLOAT costs log(C)."""
  q, c, aux, _ = _inputs(B=4, C=5)
  mask = np.ones((4, 5), bool); mask[2, :] = False
  full, _, _ = orc.retrieval_loss_and_grads_general(q, c, score_mask=mask)
  rest, _, _ = orc.retrieval_loss_and_grads_general(np.delete(q, 2, 0), np.delete(c, 2, 0), score_mask=np.delete(np.delete(mask, 2, 0), 2, 1))
  # not comparable row by row (the candidate set differs), so check the masked row through its own single-row problem
  one, dq, _ = orc.retrieval_loss_and_grads_general(q[2:3], np.roll(c, -2, 0), score_mask=np.zeros((1, 5), bool))
  assert abs(one - np.log(5.0)) <= 1e-12 and np.all(dq == 0.0)
  assert np.isfinite(full) and np.isfinite(rest)
