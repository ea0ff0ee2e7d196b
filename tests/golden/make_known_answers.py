"""Recomputes the closed-form expectations stored in known_answers.json (pure NumPy; no reference import is
possible here).  Run: python tests/golden/make_known_answers.py"""
import numpy as np


def sig(x):
  return 1.0 / (1 + np.exp(-x))


print("retrieval_2x2 loss", -np.log(sig(3.0)) - np.log(1 - sig(4.0)))
print("retrieval_2x2 weighted", -0.7 * np.log(sig(3.0)) - 0.3 * np.log(1 - sig(4.0)))
print("extra negatives", -np.log(1 / (1 + np.exp(1) + np.exp(3))) - np.log(np.exp(4) / (1 + np.exp(4) + np.exp(2))))
print("multipoint", -np.log(1 / (1 + np.exp(3) + np.exp(3))) - np.log(np.exp(5) / (np.exp(1) + np.exp(5) + np.exp(5))))

# multi_layer_dcn_test.py:28-59 -- x0 = [.1,.2,.3], kernels of ones
x0 = np.array([0.1, 0.2, 0.3])
print("mldcn full p=3", x0 * (np.ones((3, 3)) @ (np.ones((3, 3)) @ x0)) + x0)
print("mldcn low-rank p=1", x0 * (x0.sum() * np.ones(3)) + x0)
xl = x0
for _ in range(3):
  xl = x0 * (xl.sum() * np.ones(3) + 1.0) + xl
print("mldcn 3 layers, bias ones", xl)
# dot_interaction_test.py:27-41
f = [np.array([0.1, -4.3, 0.2, 1.1, 0.3]), np.array([2.0, 3.2, -1.0, 0.0, 1.0]), np.array([0.0, 1.0, -3.0, -2.2, -0.2])]
print("dot interaction f11 f12 f22 f13 f23 f33", f[0] @ f[0], f[0] @ f[1], f[1] @ f[1], f[0] @ f[2], f[1] @ f[2], f[2] @ f[2])
