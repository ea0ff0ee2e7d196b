"""Cycle counts of the tensor-pipe / TMEM building blocks (csrc/probe.cu: tc_rate_probe_kernel), one CTA per SM.
Answers, for the next optimisation round: how many bytes/clk does tcgen05.ld move with and without 16-bit packing,
what do SS (smem x smem) and TS (TMEM x smem) MMAs cost per instruction, and what does interleaving them cost.
usage: python tools/tc_rate_probe.py [rounds]"""
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


rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
n_ctas = torch.cuda.get_device_properties(dev).multi_processor_count
cyc = torch.zeros(n_ctas, dtype=torch.int64, device=dev)
sink = torch.zeros(4, dtype=torch.int32, device=dev)
names = {0: "tcgen05.ld x64 (16 warps x 8 KB per round)", 1: "tcgen05.ld x64 pack::16b", 2: "12 SS MMAs M128 N128 K16",
         3: "24 TS MMAs M128 N64 K16", 4: "12 SS + 24 TS interleaved", 5: "tcgen05.st 2 x x32 (16 warps x 8 KB per round)"}
out = {}
for mode in range(6):
  for _ in range(2):
    _rc(_probe.lib().tfrs_debug_tc_rate_probe(mode, rounds, n_ctas, _ffi.ptr(cyc), _ffi.ptr(sink), _ffi.stream()), "probe")
  torch.cuda.synchronize()
  c = cyc.float()
  per_round = float(c.median()) / rounds
  entry = {"what": names[mode], "cycles_per_round_median": round(per_round, 1), "min": round(float(c.min()) / rounds, 1),
           "max": round(float(c.max()) / rounds, 1)}
  if mode in (0, 1, 5):
    entry["bytes_per_clk_per_sm"] = round(16 * 8192 / per_round, 1)
  if mode == 2:
    entry["cycles_per_mma"] = round(per_round / 12, 1)
  if mode == 3:
    entry["cycles_per_mma"] = round(per_round / 24, 1)
  out[str(mode)] = entry
print(json.dumps(out, indent=1))
