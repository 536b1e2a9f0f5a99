"""Bring-up aid: LayerNorm forward / backward bandwidth at the step's shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from big_vision_b200 import ops  # noqa: E402


def timeit(fn, reps=10):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


for rows in (200704, 65536):
  d = 768
  x = torch.randn(rows, d, device="cuda").bfloat16()
  dy = torch.randn(rows, d, device="cuda").bfloat16()
  dres = torch.randn(rows, d, device="cuda").bfloat16()
  sc, bi = torch.ones(d, device="cuda"), torch.zeros(d, device="cuda")
  y, mean, rstd = ops.layernorm_fwd(x, sc, bi)
  ds, db, cs = (torch.zeros(d, device="cuda") for _ in range(3))
  t = timeit(lambda: ops.layernorm_fwd(x, sc, bi))
  print(f"ln_fwd rows={rows}: {t * 1e3:7.1f} us  {rows * d * 4 / t * 1e-9:6.2f} TB/s")
  t = timeit(lambda: ops.layernorm_bwd(dy, x, sc, mean, rstd, dres=dres, dscale=ds, dbias=db, dx_colsum=cs))
  print(f"ln_bwd rows={rows}: {t * 1e3:7.1f} us  {rows * d * 8 / t * 1e-9:6.2f} TB/s")
