"""cfg3 in-batch softmax step (B = C = 16384, d = 64) a few times -- for an ncu launch list / --set full capture.
usage: python tools/softmax_probe.py [B] [d] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda"); g.manual_seed(3)
q = ((torch.rand((B, d), generator=g, device="cuda") - 0.5) * 0.1).requires_grad_(True)
c = ((torch.rand((B, d), generator=g, device="cuda") - 0.5) * 0.1).requires_grad_(True)
for _ in range(iters):
  q.grad = None; c.grad = None
  loss = ops.inbatch_softmax_loss(q, c)
  loss.backward()
torch.cuda.synchronize()
print("loss", float(loss.detach()), "dq", float(q.grad.abs().max()), "dc", float(c.grad.abs().max()))

# tensor-core backward vs the exact CUDA-core backward at this size (same lse): error relative to the gradient scale
with torch.no_grad():
  _, lse = ops.inbatch_softmax_tc(q, c)
  tq, tc_ = ops.inbatch_softmax_tc_bwd(q, c, lse)
  eq, ec = ops.inbatch_softmax_bwd_exact(q, c, lse)
  for name, a, b in (("dq", tq, eq), ("dc", tc_, ec)):
    print(name, "max|tc - exact| / max|exact| =", float((a - b).abs().max() / b.abs().max()))
