"""Secondary measurements for BASELINE.md §4 (configs 3 and 5): embedding gather GB/s, sparse Adagrad,
in-batch softmax step, Cross layer.  CUDA events, warm-up, inputs larger than L2 or rotated between
iterations.  Prints one JSON object.   usage: python tools/bench_kernels.py [--quick]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommenders_b200 import ops

dev = torch.device("cuda", 0)
quick = "--quick" in sys.argv
peaks = {}
try:
  peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
  pass
HBM = peaks.get("hbm_gbs", 6650.0)


def timeit(fn, iters=20, warm=5):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


out = {"hbm_peak_gbs": HBM}
g = torch.Generator(device=dev); g.manual_seed(7)

# ---- config 5 gather: 26 tables 1M x 32, 65536 ids each -> [65536, 26*32 (+13 dense, ld 848)]
F, V, D, B = (26, 1_000_000, 32, 65536) if not quick else (26, 100_000, 32, 8192)
tables = [torch.rand((V, D), generator=g, device=dev) - 0.5 for _ in range(F)]
ids_sets = [[torch.randint(0, V, (B,), generator=g, device=dev, dtype=torch.int32) for _ in range(F)] for _ in range(4)]
act = torch.zeros((B, 848), device=dev)
state = {"i": 0}


def gather5():
  ops.gather(tables, ids_sets[state["i"] % 4], out=act); state["i"] += 1


def graph_time(fn, iters=20):
  """One call captured in a CUDA graph and replayed: device time.  (The Python binding of a 26-table gather costs more host
  time than the kernel runs, so a plain Python loop measures the host, with run-to-run differences of 25 %.)"""
  fn(); torch.cuda.synchronize()
  gr = torch.cuda.CUDAGraph()
  with torch.cuda.graph(gr):
    fn()
  return timeit(gr.replay, iters=iters)


t = sum(graph_time(lambda ids=ids: ops.gather(tables, ids, out=act)) for ids in ids_sets) / len(ids_sets)
t_host = timeit(gather5)
bytes5 = B * F * D * 4 * 2 + B * F * 4
out["cfg5_gather"] = {"seconds": t, "algorithmic_bytes": bytes5, "GBps": bytes5 / t / 1e9, "frac_of_hbm": bytes5 / t / 1e9 / HBM,
                      "python_loop_seconds": t_host, "timing": "CUDA-graph replays (device time); python_loop_seconds is host-bound",
                      "shape": f"{F} tables {V}x{D}, batch {B}, ids int32"}

# ---- config 5 Cross: B=65536, D=845 (ld 848 padded activations -> use D=848 contiguous here), 3 layers fwd
Dc = 845
x0 = torch.rand((B, Dc), generator=g, device=dev)
Ws = [torch.randn((Dc, Dc), generator=g, device=dev) * 0.05 for _ in range(3)]
bs = [torch.zeros((Dc,), device=dev) for _ in range(3)]


def cross3():
  x = x0
  for W, b in zip(Ws, bs):
    x = ops.cross(x0, x, W, b, 0.0)
  return x


with torch.no_grad():
  t = timeit(cross3, iters=5 if not quick else 3, warm=2)
flops = 3 * 2.0 * B * Dc * Dc
out["cfg5_cross_fwd_3layers"] = {"seconds": t, "TFLOPs": flops / t / 1e12, "flops": flops, "path": "tcgen05 fp16 hi/lo split GEMM + fused epilogue (B>=1024), exact CUDA-core SGEMM otherwise"}
# one Cross layer forward + backward (the training step's share), then the low-rank variants (p = 256) -- SURVEY 8f-4
go = torch.randn((B, Dc), generator=g, device=dev)
def cross_fb():
  xs = [t.detach().requires_grad_(True) for t in (x0, x0, Ws[0], bs[0])]
  ops.cross(xs[0], xs[1], xs[2], xs[3], 0.0).backward(go)
t = timeit(cross_fb, iters=5 if not quick else 3, warm=2)
out["cfg5_cross_layer_fwd_bwd"] = {"seconds": t, "TFLOPs": 3 * 2.0 * B * Dc * Dc / t / 1e12, "note": "fwd + dx + dW GEMMs, algorithmic flops"}
P = 256
Us = [torch.randn((Dc, P), generator=g, device=dev) * 0.05 for _ in range(3)]
Vs = [torch.randn((P, Dc), generator=g, device=dev) * 0.05 for _ in range(3)]
def lowrank3():
  x = x0
  for U, V, b in zip(Us, Vs, bs):
    x = ops.cross_lowrank(x0, x, U, V, b, 0.0)
  return x
def lowrank3_unfused():
  x = x0
  for U, V, b in zip(Us, Vs, bs):
    x = x0 * (ops.matmul(ops.matmul(x, U), V) + b) + x
  return x
with torch.no_grad():
  t = timeit(lowrank3, iters=5 if not quick else 3, warm=2)
  t0 = timeit(lowrank3_unfused, iters=3, warm=1)
fl = 3 * 2 * 2.0 * B * Dc * P
out["cfg5_multilayer_dcn_p256_fwd_3layers"] = {"seconds": t, "TFLOPs": fl / t / 1e12, "flops": fl, "unfused_cuda_core_seconds": t0,
                                               "path": "2 tcgen05 split-fp16 GEMMs per layer, cross formula in the second one's epilogue"}
def lowrank_fb():
  ys = [t.detach().requires_grad_(True) for t in (x0, x0, Us[0], Vs[0], bs[0])]
  ops.cross_lowrank(ys[0], ys[1], ys[2], ys[3], ys[4], 0.0).backward(go)
t = timeit(lowrank_fb, iters=5 if not quick else 3, warm=2)
out["cfg5_lowrank_layer_fwd_bwd"] = {"seconds": t, "TFLOPs": 3 * 2 * 2.0 * B * Dc * P / t / 1e12, "note": "2 fwd + 4 bwd GEMMs, algorithmic flops"}
del tables, ids_sets, act, x0, Ws, Us, Vs, go

# ---- config 3: two-tower step pieces, 10M users / 1M items, d=64, batch 16384
U, I, d, Bt = (10_000_000, 1_000_000, 64, 16384) if not quick else (1_000_000, 100_000, 64, 4096)
ut = (torch.rand((U, d), generator=g, device=dev) - 0.5) * 0.1
it = (torch.rand((I, d), generator=g, device=dev) - 0.5) * 0.1
uacc = torch.full_like(ut, 0.1); iacc = torch.full_like(it, 0.1)
uid = [torch.randint(0, U, (Bt,), generator=g, device=dev) for _ in range(4)]
iid = [torch.randint(0, I, (Bt,), generator=g, device=dev) for _ in range(4)]
state["i"] = 0


def gather3():
  k = state["i"] % 4; state["i"] += 1
  return ops.gather([ut], [uid[k]]), ops.gather([it], [iid[k]])


t = timeit(gather3)
bytes3 = 2 * Bt * d * 4 * 2 + 2 * Bt * 8
out["cfg3_gather_2tables"] = {"seconds": t, "algorithmic_bytes": bytes3, "GBps": bytes3 / t / 1e9, "frac_of_hbm": bytes3 / t / 1e9 / HBM,
                              "note": "2 launches of 4 MB each: launch-latency bound at this size"}
qe, ce = gather3()
qe = qe.requires_grad_(True); ce = ce.requires_grad_(True)


def softmax_fwd_bwd():
  qe.grad = None; ce.grad = None
  loss = ops.inbatch_softmax_loss(qe, ce)
  loss.backward()
  return loss


t = timeit(softmax_fwd_bwd, iters=5 if not quick else 3, warm=2)
out["cfg3_inbatch_softmax_fwd_bwd"] = {"seconds": t, "TFLOPs": 8.0 * Bt * Bt * d / t / 1e12, "flops": 8.0 * Bt * Bt * d,
                                       "path": "tcgen05 split-fp16 forward (online log-sum-exp) + tcgen05 backward (G in TMEM)"}
with torch.no_grad():
  t = timeit(lambda: ops.inbatch_softmax_tc(qe, ce), iters=10, warm=3)
  out["cfg3_inbatch_softmax_fwd_only"] = {"seconds": t, "TFLOPs": 2.0 * Bt * Bt * d / t / 1e12}
  _, lse_t = ops.inbatch_softmax_tc(qe, ce)
  t = timeit(lambda: ops.inbatch_softmax_tc_bwd(qe, ce, lse_t), iters=10, warm=3)
  out["cfg3_inbatch_softmax_bwd_only"] = {"seconds": t, "TFLOPs": 4.0 * Bt * Bt * d / t / 1e12,
                                          "note": "algorithmic flops (dq + dc); the kernels also recompute S twice"}
gq = qe.grad.detach().clone()


def adagrad3():
  k = state["i"] % 4; state["i"] += 1
  ops.sparse_adagrad_(ut, uacc, uid[k], gq, 0.5)


t = timeit(adagrad3)
out["cfg3_sparse_adagrad_user_table"] = {"seconds": t, "rows": Bt, "GBps_algorithmic": (Bt * d * 4 * 5) / t / 1e9}
# ---- Streaming over a corpus that lives in HOST memory: 1M x 64 rows in pinned memory, dataset batches of
#      8192 rows coalesced into 262144-row chunks, pinned double-buffered H2D overlapped with the tensor-core scan per chunk
try:
  import recommenders_b200 as tfrs
  del ut, it, uacc, iacc
  torch.cuda.empty_cache()
  Ns, Qs, ks = (1_000_000, 4096, 100) if not quick else (200_000, 1024, 100)
  corpus_host = torch.randn((Ns, 64), generator=torch.Generator().manual_seed(1)).pin_memory()
  qd = torch.randn((Qs, 64), generator=g, device=dev)
  layer = tfrs.layers.factorized_top_k.Streaming(k=ks).index_from_dataset(
      tfrs.data.Dataset.from_tensor_slices(corpus_host).batch(8192))
  t = timeit(lambda: layer(qd), iters=3, warm=1)
  on_dev = tfrs.layers.factorized_top_k.Streaming(k=ks).index_from_dataset(
      tfrs.data.Dataset.from_tensor_slices(corpus_host.to(dev)).batch(8192))
  t_dev = timeit(lambda: on_dev(qd), iters=3, warm=1)
  out["streaming_host_corpus"] = {"seconds": t, "queries_per_s": Qs / t, "h2d_bytes": Ns * 64 * 4, "h2d_GBps": Ns * 64 * 4 / t / 1e9,
                                  "device_resident_seconds": t_dev, "device_resident_queries_per_s": Qs / t_dev,
                                  "shape": f"{Qs} queries x {Ns}x64 corpus in pinned host memory, top-{ks}"}
except Exception as e:  # keep the other figures if this leg fails
  out["streaming_host_corpus"] = {"error": repr(e)}
print(json.dumps(out))
