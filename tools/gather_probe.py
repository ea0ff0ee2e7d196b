"""cfg5 gather (26 tables 1M x 32, batch 65536 -> [65536, 848]) with both kernels (A/B), GB/s vs the measured HBM peak.
usage: python tools/gather_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import _ffi, ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(7)
F, V, D, B = 26, 1_000_000, 32, 65536
tables = [torch.rand((V, D), generator=g, device=dev) - 0.5 for _ in range(F)]
ids_sets = [[torch.randint(0, V, (B,), generator=g, device=dev, dtype=torch.int32) for _ in range(F)] for _ in range(4)]
act = torch.zeros((B, 848), device=dev)
peaks = {}
try:
  peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
  pass
HBM = peaks.get("hbm_gbs", 6650.0)
out = {"hbm_peak_gbs": HBM}
ref = None
for variant, name in [(0, "lane_per_item"), (1, "warp_chunk")]:
  _ffi.check(_ffi.lib().tfrs_debug_set_gather_variant(variant), "variant")
  st = {"i": 0}

  def f():
    ops.gather(tables, ids_sets[st["i"] % 4], out=act); st["i"] += 1
  for _ in range(5):
    f()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(20):
    f()
  e1.record(); torch.cuda.synchronize()
  t = e0.elapsed_time(e1) / 20 * 1e-3
  nbytes = B * F * D * 4 * 2 + B * F * 4
  ops.gather(tables, ids_sets[0], out=act)
  if ref is None:
    ref = act.clone()
  out[name] = {"us": round(t * 1e6, 1), "GBps": round(nbytes / t / 1e9, 1), "frac_of_hbm": round(nbytes / t / 1e9 / HBM, 3),
               "same_result": bool(torch.equal(ref, act))}
_ffi.check(_ffi.lib().tfrs_debug_set_gather_variant(1), "variant")
print(json.dumps(out))
