"""Where does a sharded BruteForce step go?  Times, with CUDA events on the launching stream, the three pieces of
BruteForce._sharded_topk -- local scan, the single all-gather, the merge -- plus the whole step and the host-side
enqueue time per step.  Run under torchrun (real all-gather) or alone with --fake-world W (merge only, the receive
buffer is filled locally).   usage: torchrun ... tools/scale_probe.py [--n 1000000] [--d 64] [--q 4096] [--k 100]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from recommenders_b200 import ops
from recommenders_b200.layers import factorized_top_k as ftk

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--d", type=int, default=64)
ap.add_argument("--q", type=int, default=4096)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--fake-world", type=int, default=0)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
  dist.init_process_group("nccl", device_id=dev)


def timeit(fn, iters=args.iters, warm=5):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter()
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  t_host = time.perf_counter() - t0
  torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1) / iters, 1e3 * t_host / iters], device=dev)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return [round(float(x), 4) for x in ms]  # [device ms, host-enqueue ms]


W = args.fake_world if world == 1 and args.fake_world else world
Q, k, d = args.q, args.k, args.d
g = torch.Generator(device=dev); g.manual_seed(2)
lo, hi = ftk.shard_bounds(args.n, rank if world > 1 else 0, W)
corpus = torch.randn((args.n, d), generator=g, device=dev)[lo:hi].contiguous()
g.manual_seed(1)
queries = torch.randn((Q, d), generator=g, device=dev)
out = {"world": W, "real_collective": world > 1, "shard_rows": hi - lo}

layer = ftk.BruteForce(k=k)
if world > 1:
  layer.index_shard(corpus, lo)
else:
  layer.index(corpus)

idx_off = (Q * k * 4 + 7) // 8 * 8
block = idx_off + Q * k * 8
send = torch.empty(block, dtype=torch.uint8, device=dev)
out_s = send[:Q * k * 4].view(torch.float32).view(Q, k)
out_i = send[idx_off:].view(torch.int64).view(Q, k)
recv = torch.empty(W * block, dtype=torch.uint8, device=dev)

def note(key, val):
  out[key] = val
  if rank == 0:
    print(key, val, flush=True)


note("local_scan", timeit(lambda: layer._local_topk(queries, k, lo, out=(out_s, out_i))))
if world > 1:
  note("all_gather", timeit(lambda: dist.all_gather_into_tensor(recv, send)))
else:
  for r in range(W):
    recv[r * block:(r + 1) * block].copy_(send)
note("merge", timeit(lambda: ops.topk_merge_packed(recv, W, Q, k, k, idx_off, block)))
out["block_mb"] = block / 1e6
if world > 1:
  out["full_step"] = timeit(lambda: layer(queries))
else:
  out["full_step"] = timeit(lambda: layer(queries))
if rank == 0:
  print(json.dumps(out))
if world > 1:
  dist.destroy_process_group()
