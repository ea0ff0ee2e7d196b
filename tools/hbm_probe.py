"""What does the memory system give for the embedding gather's access pattern?  128-byte rows, same thread layout as
gather.cu; copy / random-row read / write-only / random read + sequential write / random read + gather-strided write.
Prints GB/s per mode (CUDA events, buffers far larger than L2).   usage: python tools/hbm_probe.py"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from recommenders_b200 import _ffi
import _probe  # tools/libtfrs_b200_probe.so: the probes are not in the product library


def _rc(rc, what="probe"):
  if rc:
    raise RuntimeError(f"{what}: rc={rc}")


dev = torch.device("cuda", 0)
SRC_ROWS = 26_000_000          # 3.3 GB of 128-byte rows (the 26 cfg5 tables)
N = 65536 * 26                 # rows moved per launch (218 MB)
src = torch.empty((SRC_ROWS, 32), device=dev).uniform_(-1, 1)
dst = torch.zeros((65536, 848), device=dev)   # the cfg5 activation (ld 848)
sink = torch.zeros(4, device=dev)
st = _ffi.stream()


def run(mode, iters=20):
  def f():
    _rc(_probe.lib().tfrs_debug_hbm_probe(mode, _ffi.ptr(src), SRC_ROWS, _ffi.ptr(dst), N, 65536, 848, _ffi.ptr(sink), st), "probe")
  for _ in range(3):
    f()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    f()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters * 1e-3


row_bytes = N * 128
out = {}
small = src[:1_000_000]  # one cfg5 table (128 MB): the footprint a table-at-a-time gather touches
for mode, name, nbytes in [(0, "copy_seq", 2 * row_bytes), (1, "random_read_only", row_bytes), (2, "write_only_seq", row_bytes),
                           (3, "random_read_seq_write", 2 * row_bytes), (4, "random_read_gather_strided_write", 2 * row_bytes)]:
  t = run(mode)
  out[name] = {"us": round(t * 1e6, 1), "GBps": round(nbytes / t / 1e9, 1)}
SRC_ROWS = 1_000_000
for mode, name, nbytes in [(1, "random_read_only_128MB_table", row_bytes), (4, "random_read_128MB_table_gather_strided_write", 2 * row_bytes)]:
  t = run(mode)
  out[name] = {"us": round(t * 1e6, 1), "GBps": round(nbytes / t / 1e9, 1)}
print(json.dumps(out))
