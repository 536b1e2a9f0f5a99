"""Bring-up aid: time the step's GEMM shapes / epilogues in isolation (CUDA events, 5 reps).

  python tools/gemm_shapes.py            # the ViT-B/16 MLP + attention projections at 200704 tokens
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from big_vision_b200 import lib as L  # noqa: E402
from big_vision_b200 import ops  # noqa: E402


def timeit(fn, reps=5):
  if os.environ.get("GEMM_SHAPES_ONCE") == "1":      # under ncu: exactly one launch per case
    fn()
    torch.cuda.synchronize()
    return float("nan")
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


def main():
  M = int(sys.argv[1]) if len(sys.argv) > 1 else 200704
  dev = "cuda"
  x768 = torch.randn(M, 768, device=dev).bfloat16()
  x3072 = torch.randn(M, 3072, device=dev).bfloat16()
  w0 = (torch.randn(768, 3072, device=dev) * 0.03).bfloat16()    # Dense_0 kernel [K=768, N=3072]
  w1 = (torch.randn(3072, 768, device=dev) * 0.03).bfloat16()    # Dense_1 kernel [K=3072, N=768]
  b3072 = torch.randn(3072, device=dev)
  b768 = torch.randn(768, device=dev)
  cs = torch.zeros(3072, device=dev)
  out3072 = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
  out3072b = torch.empty(M, 3072, device=dev, dtype=torch.bfloat16)
  out768 = torch.empty(M, 768, device=dev, dtype=torch.bfloat16)
  cases = [
      ("fwd  x.W0        epi0      ", 3072, 768, lambda: ops.gemm(x768, w0, b_mn=True, out=out3072)),
      ("fwd  x.W0 +bias  epi1      ", 3072, 768, lambda: ops.gemm(x768, w0, b_mn=True, bias=b3072, out=out3072)),
      ("fwd  x.W0 gelu   epi2 dual ", 3072, 768, lambda: ops.gemm(x768, w0, b_mn=True, bias=b3072, out=out3072,
                                                                 out2=out3072b, epilogue=L.EPI_BIAS_GELU)),
      ("dgrad dy.W1^T    epi0      ", 3072, 768, lambda: ops.gemm(x768, w1, out=out3072)),
      ("dgrad dy.W1^T gelu' epi4   ", 3072, 768, lambda: ops.gemm(x768, w1, aux=x3072, out=out3072,
                                                                 epilogue=L.EPI_DGELU)),
      ("dgrad + colsum   epi4      ", 3072, 768, lambda: ops.gemm(x768, w1, aux=x3072, out=out3072,
                                                                 epilogue=L.EPI_DGELU, colsum=cs)),
      ("fwd  h.W1 +resid epi3      ", 768, 3072, lambda: ops.gemm(x3072, w1, b_mn=True, bias=b768, aux=x768,
                                                                 out=out768, epilogue=L.EPI_BIAS_RESID)),
      ("dgrad dh.W0^T    epi0      ", 768, 3072, lambda: ops.gemm(x3072, w0, out=out768)),
  ]
  for name, N, K, fn in cases:
    ms = timeit(fn)
    print(f"{name} M={M} N={N} K={K}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms * 1e-9:7.1f} TFLOP/s")


if __name__ == "__main__":
  main()
