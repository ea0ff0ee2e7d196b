"""Cross forward (tensor-core path) at B = 65536 for several widths: is the epilogue sensitive to row alignment (845 floats
per row = 4-byte aligned rows, 848 = 64-byte aligned, 832 = 128-byte aligned)?   usage: python tools/cross_align_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import ops

B = 65536
g = torch.Generator(device="cuda"); g.manual_seed(5)
out = {}
for D in (845, 848, 832, 896):
  x0 = torch.rand((B, D), generator=g, device="cuda"); x = torch.rand((B, D), generator=g, device="cuda")
  W = torch.randn((D, D), generator=g, device="cuda") * 0.05; b = torch.zeros(D, device="cuda")
  with torch.no_grad():
    for _ in range(2):
      ops.cross(x0, x, W, b, 0.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
      ops.cross(x0, x, W, b, 0.0)
    e1.record(); torch.cuda.synchronize()
  out[D] = {"ms_per_layer": e0.elapsed_time(e1) / 5, "TFLOPs": 2.0 * B * D * D / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12}
  del x0, x, W
print(json.dumps(out))
