"""Opcode histogram per kernel of libtfrs_b200.so (cuobjdump -sass): shows which kernels carry tcgen05 MMA (UTCHMMA),
TMEM loads/stores (LDTM/STTM), bulk-TMA copies (UBLKCP), mbarrier waits (SYNCS) and how many local-memory (spill)
instructions each has.   usage: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "recommenders_b200", "libtfrs_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern = None
hist = collections.OrderedDict()
arch = set()
for line in txt.splitlines():
  m = re.match(r"\s*Function : (\S+)", line)
  if m:
    kern = m.group(1); hist[kern] = collections.Counter(); continue
  m = re.match(r"\s*arch = (\S+)", line)
  if m:
    arch.add(m.group(1))
  m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
  if m and kern:
    hist[kern][m.group(1)] += 1
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "UTCATOMSWS", "MUFU", "FMNMX3", "FMNMX", "SHFL", "REDUX", "ATOMG", "RED", "LDL", "STL"]
print(f"# SASS opcode summary of {os.path.basename(lib)}; arch = {sorted(arch)}; {len(hist)} kernels")
print("# kernel | total instructions | " + " ".join(KEY))
tot = collections.Counter()
for k, h in hist.items():
  name = demangle(k)
  name = re.sub(r"\(.*\)$", "", name)[:100]
  n = sum(h.values())
  tot.update(h)
  print(f"{name} | {n} | " + " ".join(f"{key}={h[key]}" for key in KEY if h[key]))
print("# library total: " + " ".join(f"{key}={tot[key]}" for key in KEY))
