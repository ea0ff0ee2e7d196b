"""Debug driver for the tensor-core top-K path: compares tfrs_topk_tc_f32 with the exact CUDA-core
scan on one shape and prints mismatch statistics.  usage: python tools/tc_check.py Q N d k [seed]"""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recommenders_b200 import ops

Q, N, d, k = [int(x) for x in sys.argv[1:5]]
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(seed)
c = torch.randn((N, d), generator=g, device=dev); q = torch.randn((Q, d), generator=g, device=dev)
print("tc_supported:", ops.tc_supported(Q, N, d, k), flush=True)
t0 = time.time(); idx = ops.index_build(c); torch.cuda.synchronize(); print("index_build ok %.3fs" % (time.time() - t0), flush=True)
s, i = ops.topk_tc(q, c, idx, k); torch.cuda.synchronize(); print("topk_tc ok", flush=True)
Qc = min(Q, 512)
es, ei = ops.topk_scan(q[:Qc], c, k); torch.cuda.synchronize()
print("ids equal:", bool((i[:Qc] == ei).all()), " scores equal:", bool((s[:Qc] == es).all()))
if not (i[:Qc] == ei).all():
  bad = (i[:Qc] != ei).any(1).nonzero().flatten()
  print("bad rows:", bad.numel(), bad[:10].tolist())
  r = int(bad[0]); print(i[r][:12].tolist()); print(ei[r][:12].tolist()); print(s[r][:6].tolist()); print(es[r][:6].tolist())
ops.profile_enable(True)
for _ in range(5): ops.topk_tc(q, c, idx, k)
ms, calls = ops.profile_read(); ops.profile_enable(False)
print("stage ms/call (prep, sample, filter, finalize):", [round(m / calls, 4) for m in ms], "total %.4f" % (sum(ms) / calls))
print("filter TFLOP/s: %.1f   q/s: %.0f" % (2.0 * Q * N * d / (ms[2] / calls * 1e-3) / 1e12, Q / (sum(ms) / calls * 1e-3)))
