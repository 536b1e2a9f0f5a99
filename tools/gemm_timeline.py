"""Bring-up aid: per-tile event timeline (SM cycles) of CTA 0 (pair leader) of the GEMM kernel.

  python tools/gemm_timeline.py [case]     # case: epi0 | epi1 | epi2 | epi4 | epi4cs | k3072
"""
import ctypes
import os
import sys

os.environ["BV_GEMM_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from big_vision_b200 import lib as L  # noqa: E402
from big_vision_b200 import ops  # noqa: E402

EV = ["m_acc_free", "m_first_full", "m_all_issued", "p_first", "p_last", "e_tfull", "e_drained", "e_stored",
      "e_ready"]


def main():
  case = sys.argv[1] if len(sys.argv) > 1 else "epi0"
  M = 200704
  dev = "cuda"
  x768 = torch.randn(M, 768, device=dev).bfloat16()
  x3072 = torch.randn(M, 3072, device=dev).bfloat16()
  w0 = (torch.randn(768, 3072, device=dev) * 0.03).bfloat16()
  w1 = (torch.randn(3072, 768, device=dev) * 0.03).bfloat16()
  b3072 = torch.randn(3072, device=dev)
  cs = torch.zeros(3072, device=dev)
  out3072 = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
  out3072b = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
  out768 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
  fns = {
      "epi0": lambda: ops.gemm(x768, w0, b_mn=True, out=out3072),
      "epi1": lambda: ops.gemm(x768, w0, b_mn=True, bias=b3072, out=out3072),
      "epi2": lambda: ops.gemm(x768, w0, b_mn=True, bias=b3072, out=out3072, out2=out3072b,
                               epilogue=L.EPI_BIAS_GELU),
      "epi4": lambda: ops.gemm(x768, w1, aux=x3072, out=out3072, epilogue=L.EPI_DGELU),
      "epi4cs": lambda: ops.gemm(x768, w1, aux=x3072, out=out3072, epilogue=L.EPI_DGELU, colsum=cs),
      "k3072": lambda: ops.gemm(x3072, w0, out=out768),
  }
  fn = fns[case]
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  fn()
  t1.record()
  torch.cuda.synchronize()
  print(f"{case}: {t0.elapsed_time(t1) * 1e3:.1f} us")
  buf = (ctypes.c_longlong * 512)()
  lib = L.load()
  lib.bv_debug_gemm_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
  lib.bv_debug_gemm_timeline(buf, 512)
  vals = [buf[i] for i in range(512)]
  base = min(x for i, x in enumerate(vals) if x > 0 and i % 16 != 9)
  print("tile " + " ".join(f"{e[:12]:>12s}" for e in EV))
  for i in range(4, 14):
    row = vals[i * 16:(i + 1) * 16]
    print(f"{i:4d} " + " ".join(f"{(x - base) if x else -1:12d}" for x in row[:9]))


if __name__ == "__main__":
  main()
