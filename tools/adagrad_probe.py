"""cfg3 sparse-Adagrad step (B = 16384 ids on a 1M x 64 table), uniform and Zipf(1.05) ids: run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split, or alone for CUDA-event totals.
usage: python tools/adagrad_probe.py [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)
rows, d, B = 1_000_000, 64, 16384
table = torch.rand((rows, d), generator=g, device=dev) * 0.1 - 0.05
acc = torch.full_like(table, 0.1)
grad = torch.randn((B, d), generator=g, device=dev) * 1e-3
u = torch.rand((B,), generator=g, device=dev, dtype=torch.float64)
s = 1.05
zipf = (((u * (rows ** (1 - s) - 1) + 1) ** (1 / (1 - s))).clamp(1, rows).to(torch.int64) - 1).to(torch.int32)
uni = torch.randint(0, rows, (B,), generator=g, device=dev, dtype=torch.int32)
out = {}
for name, ids in (("uniform", uni), ("zipf", zipf)):
  for _ in range(2):
    ops.sparse_adagrad_(table, acc, ids, grad, 0.1)
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  torch.cuda.synchronize(); e0.record()
  for _ in range(reps):
    ops.sparse_adagrad_(table, acc, ids, grad, 0.1)
  e1.record(); torch.cuda.synchronize()
  cnt = torch.bincount(ids.to(torch.int64))
  out[name] = {"us": e0.elapsed_time(e1) / reps * 1e3, "unique": int((cnt > 0).sum()), "max_run": int(cnt.max())}
print(json.dumps(out))
