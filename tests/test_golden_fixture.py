"""The oracle against the committed golden fixture (tests/golden/known_answers.json)."""
import json
import os

import numpy as np

from oracle import oracle as orc

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "known_answers.json")))


def test_retrieval_fixtures():
  g = G["retrieval_2x2"]
  q, c = np.array(g["query"], np.float32), np.array(g["candidate"], np.float32)
  np.testing.assert_array_equal(orc.retrieval_scores(q, c)[0], g["scores"])
  np.testing.assert_allclose(orc.retrieval_loss(q, c), g["loss"], rtol=1e-9)
  np.testing.assert_allclose(orc.retrieval_loss(q, c, sample_weight=[0.7, 0.3]), g["loss_weighted_0.7_0.3"], rtol=1e-9)
  g = G["retrieval_extra_negatives"]
  q, c = np.array(g["query"], np.float32), np.array(g["candidate"], np.float32)
  np.testing.assert_array_equal(orc.retrieval_scores(q, c)[0], g["scores"])
  np.testing.assert_allclose(orc.retrieval_loss(q, c), g["loss"], rtol=1e-9)
  g = G["retrieval_multipoint"]
  q, c = np.array(g["query"], np.float32), np.array(g["candidate"], np.float32)
  np.testing.assert_array_equal(orc.retrieval_scores(q, c)[0], g["maxsim"])
  np.testing.assert_allclose(orc.retrieval_loss(q, c), g["loss"], rtol=1e-9)


def test_cross_fixtures():
  g = G["cross"]
  x0, x = np.array(g["x0"], np.float32), np.array(g["x"], np.float32)
  ones = np.ones((3, 3), np.float32)
  np.testing.assert_allclose(orc.cross(x0, x, ones), g["full_ones"], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, None, ones), g["one_input_ones"], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, x, ones, bias=np.ones(3)), g["bias_ones"], rtol=1e-6)
  np.testing.assert_allclose(orc.cross(x0, x, ones, diag_scale=1.0), g["diag_scale_1"], rtol=1e-6)


def test_feature_interaction_fixtures():
  g = G["multi_layer_dcn"]
  x0 = np.array(g["x0"], np.float32)
  np.testing.assert_allclose(orc.multi_layer_dcn(x0, [np.ones((3, 3))], [np.ones((3, 3))]), g["full_rank_p3_1layer_nobias"], rtol=1e-6)
  np.testing.assert_allclose(orc.multi_layer_dcn(x0, [np.ones((3, 1))], [np.ones((1, 3))]), g["low_rank_p1_1layer_nobias"], rtol=1e-6)
  np.testing.assert_allclose(orc.multi_layer_dcn(x0, [np.ones((3, 1))] * 3, [np.ones((1, 3))] * 3, [np.ones(3)] * 3),
                             g["low_rank_p1_3layers_bias_ones"], rtol=1e-5)
  g = G["dot_interaction"]
  feats = [np.array(g["feature%d" % i], np.float32) for i in (1, 2, 3)]
  np.testing.assert_allclose(orc.dot_interaction(feats, True, False)[0], [g[k] for k in g["order_self_interaction"]], rtol=1e-6)
  np.testing.assert_allclose(orc.dot_interaction(feats, False, False)[0], [g[k] for k in g["order_no_self"]], rtol=1e-6)
