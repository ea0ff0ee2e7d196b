"""Times one tensor-core Cross layer at config-5 size (65536 x 845) and the exact CUDA-core kernel beside it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommenders_b200 import ops
B, D = 65536, 845
x0 = torch.rand((B, D), device="cuda"); x = torch.rand((B, D), device="cuda")
W = torch.randn((D, D), device="cuda") * 0.05; b = torch.zeros(D, device="cuda")


def t(fn, n=5):
  for _ in range(2): fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n


with torch.no_grad():
  ms_tc = t(lambda: ops.cross(x0, x, W, b, 0.0))
  y_tc = ops.cross(x0, x, W, b, 0.0)
  old = ops.CROSS_TC_MIN_B; ops.CROSS_TC_MIN_B = 1 << 60
  ms_cc = t(lambda: ops.cross(x0, x, W, b, 0.0), 2)
  y_cc = ops.cross(x0, x, W, b, 0.0)
  ops.CROSS_TC_MIN_B = old
print("cross layer 65536x845: tensor-core %.3f ms (%.1f TFLOP/s algorithmic), CUDA-core %.3f ms; max rel diff %.2e" %
      (ms_tc, 2.0 * B * D * D / ms_tc / 1e9, ms_cc, float((y_tc - y_cc).abs().max() / y_cc.abs().max())))

# ---- backward: tensor-core GEMMs (dx, dW) vs the exact CUDA-core path, same inputs
g = torch.randn((B, D), device="cuda")


def fwd_bwd():
  xs = [x0.detach().requires_grad_(True), x.detach().requires_grad_(True), W.detach().requires_grad_(True), b.detach().requires_grad_(True)]
  y = ops.cross(xs[0], xs[1], xs[2], xs[3], 0.25)
  y.backward(g)
  return [v.grad for v in xs]


ms_tc = t(fwd_bwd, 3)
g_tc = fwd_bwd()
old = ops.CROSS_TC_MIN_B; ops.CROSS_TC_MIN_B = 1 << 60
ms_cc = t(fwd_bwd, 1)
g_cc = fwd_bwd()
ops.CROSS_TC_MIN_B = old
errs = [float((a - c).abs().max() / c.abs().max()) for a, c in zip(g_tc, g_cc)]
print("cross fwd+bwd 65536x845: tensor-core %.3f ms, CUDA-core %.3f ms; max |tc - exact| / max|exact| for dx0, dx, dW, db = %s" %
      (ms_tc, ms_cc, ", ".join("%.2e" % e for e in errs)))
