"""Every secondary kernel of the path once at its BASELINE size -- the driver for ncu captures (launch list and
`--set full` on the gather / Adagrad / Cross / softmax kernels):
  cfg5 gather (26 x 1M x 32, B = 65536, uniform and Zipf ids), cfg3 sparse Adagrad (uniform and Zipf), one Cross layer forward +
  backward (65536 x 845), the low-rank Cross layer (p = 256), cfg3 in-batch softmax forward + backward (plain and with
  accidental-hit removal).
usage: python tools/kernel_step_probe.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)


def zipf(n_rows, n):
  u = torch.rand((n,), generator=g, device=dev, dtype=torch.float64)
  s = 1.05
  return (((u * (n_rows ** (1 - s) - 1) + 1) ** (1 / (1 - s))).clamp(1, n_rows).to(torch.int64) - 1).to(torch.int32)


# ---- cfg5 gather
tables = [torch.rand((1_000_000, 32), generator=g, device=dev) * 0.1 - 0.05 for _ in range(26)]
act = torch.zeros((65536, 848), device=dev)
uid = [torch.randint(0, 1_000_000, (65536,), generator=g, device=dev, dtype=torch.int32) for _ in range(26)]
zid = [zipf(1_000_000, 65536) for _ in range(26)]
for _ in range(iters):
  ops.gather(tables, uid, out=act)
  ops.gather(tables, zid, out=act)
del tables, act
torch.cuda.empty_cache()

# ---- cfg3 sparse Adagrad
table = torch.rand((1_000_000, 64), generator=g, device=dev) * 0.1 - 0.05
acc = torch.full_like(table, 0.1)
grad = torch.randn((16384, 64), generator=g, device=dev) * 1e-3
ids_u = torch.randint(0, 1_000_000, (16384,), generator=g, device=dev, dtype=torch.int32)
ids_z = zipf(1_000_000, 16384)
for _ in range(iters):
  ops.sparse_adagrad_(table, acc, ids_u, grad, 0.1)
  ops.sparse_adagrad_(table, acc, ids_z, grad, 0.1)
del table, acc
torch.cuda.empty_cache()

# ---- cfg5 Cross layer (full rank, then low rank p = 256)
B, D, P = 65536, 845, 256
x0 = torch.rand((B, D), generator=g, device=dev); x = torch.rand((B, D), generator=g, device=dev)
W = torch.randn((D, D), generator=g, device=dev) * 0.05; b = torch.zeros(D, device=dev)
U = torch.randn((D, P), generator=g, device=dev) * 0.05; V = torch.randn((P, D), generator=g, device=dev) * 0.05
go = torch.randn((B, D), generator=g, device=dev)
for _ in range(iters):
  xs = [t.detach().requires_grad_(True) for t in (x0, x, W, b)]
  ops.cross(xs[0], xs[1], xs[2], xs[3], 0.0).backward(go)
  ys = [t.detach().requires_grad_(True) for t in (x0, x, U, V, b)]
  ops.cross_lowrank(ys[0], ys[1], ys[2], ys[3], ys[4], 0.0).backward(go)
del x0, x, go, xs, ys
torch.cuda.empty_cache()

# ---- cfg3 in-batch softmax (plain; with accidental-hit removal)
Bq = 16384
q = ((torch.rand((Bq, 64), generator=g, device=dev) - 0.5) * 0.1).requires_grad_(True)
c = ((torch.rand((Bq, 64), generator=g, device=dev) - 0.5) * 0.1).requires_grad_(True)
cid = zipf(1_000_000, Bq).to(torch.int64)
for _ in range(iters):
  q.grad = None; c.grad = None
  ops.inbatch_softmax_loss(q, c).backward()
  ops.inbatch_softmax_loss(q, c, candidate_ids=cid).backward()
torch.cuda.synchronize()
print("ok")
