"""Builds and loads tools/libtfrs_b200_probe.so: the hardware probes (tools/csrc/probe.cu) live OUTSIDE the product
library.  `python tools/_probe.py` builds it (nvcc, sm_100a)."""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libtfrs_b200_probe.so")
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)


def build(force=False):
  srcs = [os.path.join(HERE, "csrc", "probe.cu"), os.path.join(ROOT, "recommenders_b200", "csrc", "api.cu")]
  if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
    return LIB
  cmd = ["/usr/local/cuda/bin/nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", "-cudart", "static", "-o", LIB, *srcs]
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode:
    raise RuntimeError(r.stdout + r.stderr)
  return LIB


def lib():
  l = ctypes.CDLL(build())
  c_p, c_i, c_l = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
  l.tfrs_debug_umma_probe.argtypes = [c_p, c_p, ctypes.c_uint32, c_i, c_p, c_p]
  l.tfrs_debug_hbm_probe.argtypes = [c_i, c_p, c_l, c_p, c_l, c_l, c_l, c_p, c_p]
  l.tfrs_debug_tc_rate_probe.argtypes = [c_i, c_i, c_i, c_p, c_p, c_p]
  l.tfrs_last_error.restype = ctypes.c_char_p
  return l


if __name__ == "__main__":
  print(build(force=True))
