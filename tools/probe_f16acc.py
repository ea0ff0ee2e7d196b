"""Probe: what does tcgen05.mma kind::f16 write to TMEM when the accumulator format is F16 (c_format=0)?
Builds two 128x64 fp16 tiles with the library's own image builder, runs one UMMA tile with D=f32 and with
D=f16, dumps raw TMEM and tests layout hypotheses."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from recommenders_b200 import ops, _ffi
import _probe  # tools/libtfrs_b200_probe.so: the probes are not in the product library


def _rc(rc, what="probe"):
  if rc:
    raise RuntimeError(f"{what}: rc={rc}")

dev = torch.device("cuda", 0)
torch.manual_seed(0)
a = torch.randn((128, 64), device=dev); b = torch.randn((128, 64), device=dev)
os.environ["TFRS_TC_FP16_TARGET"] = "3"       # small scale so that fp16 accumulators cannot overflow
ia = ops.index_build(a); ib = ops.index_build(b)
ea = int(ia[:64].view(torch.int32)[2]); eb = int(ib[:64].view(torch.int32)[2])
ah = torch.ldexp(a, torch.tensor(ea, device=dev)).half().float(); bh = torch.ldexp(b, torch.tensor(eb, device=dev)).half().float()
ref = (ah @ bh.T)  # [128 rows of A, 128 rows of B]
base = (1 << 17 + 0) * 0
IDESC_F32 = (1 << 4) | ((128 >> 3) << 17) | ((128 >> 4) << 24)
IDESC_F16 = (0 << 4) | ((128 >> 3) << 17) | ((128 >> 4) << 24)
for name, idesc in (("D=f32", IDESC_F32), ("D=f16", IDESC_F16)):
  out = torch.zeros((128, 128), dtype=torch.int32, device=dev)
  _rc(_probe.lib().tfrs_debug_umma_probe(ctypes.c_void_p(ia[1024:].data_ptr()), ctypes.c_void_p(ib[1024:].data_ptr()),
                                              idesc, 128, ctypes.c_void_p(out.data_ptr()), None), "probe")
  torch.cuda.synchronize()
  raw = out.cpu().numpy().view(np.uint32)
  as_f32 = raw.view(np.float32)
  lo = (raw & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
  hi = (raw >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
  r = ref.cpu().numpy()
  print("==", name)
  print(" f32 view max abs err vs ref      :", float(np.abs(as_f32 - r).max()))
  print(" lo16 of col c == ref[:, c]       :", float(np.abs(lo - r).max()), " hi16 nonzero frac:", float((hi != 0).mean()))
  pk = np.empty_like(r); pk[:, 0::2] = lo[:, :64]; pk[:, 1::2] = hi[:, :64]
  print(" packed pairs in cols 0..63       :", float(np.abs(pk - r).max()))
  pk2 = np.concatenate([lo[:, :64], hi[:, :64]], 1)
  print(" lo=cols 0..63, hi=cols 64..127   :", float(np.abs(pk2 - r).max()))
  print(" raw[0,:4] hex", [hex(int(x)) for x in raw[0, :4]], "ref[0,:4]", r[0, :4].tolist(), " cols 64.. nonzero:", float((raw[:, 64:] != 0).mean()))
