"""Recomputes the closed-form expectations stored in known_answers.json (pure NumPy; no reference import is
possible here).  Run: python tests/golden/make_known_answers.py"""
import numpy as np


def sig(x):
  return 1.0 / (1 + np.exp(-x))


print("retrieval_2x2 loss", -np.log(sig(3.0)) - np.log(1 - sig(4.0)))
print("retrieval_2x2 weighted", -0.7 * np.log(sig(3.0)) - 0.3 * np.log(1 - sig(4.0)))
print("extra negatives", -np.log(1 / (1 + np.exp(1) + np.exp(3))) - np.log(np.exp(4) / (1 + np.exp(4) + np.exp(2))))
print("multipoint", -np.log(1 / (1 + np.exp(3) + np.exp(3))) - np.log(np.exp(5) / (np.exp(1) + np.exp(5) + np.exp(5))))
