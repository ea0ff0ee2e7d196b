"""One Cross layer forward + backward at config-5 size (65536 x 845) a few times -- for an ncu launch list.
usage: python tools/cross_step_probe.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import ops

B, D = 65536, 845
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator(device="cuda"); g.manual_seed(5)
x0 = torch.rand((B, D), generator=g, device="cuda"); x = torch.rand((B, D), generator=g, device="cuda")
W = torch.randn((D, D), generator=g, device="cuda") * 0.05; b = torch.zeros(D, device="cuda")
go = torch.randn((B, D), generator=g, device="cuda")
for _ in range(iters):
  xs = [x0.detach().requires_grad_(True), x.detach().requires_grad_(True), W.detach().requires_grad_(True), b.detach().requires_grad_(True)]
  y = ops.cross(xs[0], xs[1], xs[2], xs[3], 0.25)
  y.backward(go)
torch.cuda.synchronize()
print("ok", float(xs[2].grad.abs().max()))
