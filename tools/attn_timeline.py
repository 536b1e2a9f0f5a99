"""Bring-up aid: per-tile event timeline (SM cycles) of CTA 0 of the attention forward kernel.

  BV_ATTN_DBG=1 python tools/attn_timeline.py [N] [B]
"""
import ctypes
import os
import sys

os.environ["BV_ATTN_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from big_vision_b200 import lib as L  # noqa: E402
from big_vision_b200 import ops  # noqa: E402

EV = ["tma_issued", "S_in_ready", "S_committed", "sm_s_full", "sm_max_done", "sm_p_empty",
      "sm_p_full_arr", "PV_ready", "PV_committed", "ep_o_full", "ep_stage_free", "ep_store",
      "sm_ld_done", "sm_max_local", "sm_exp_done", "sm_sts_done"]


def main():
  N = int(sys.argv[1]) if len(sys.argv) > 1 else 196
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
  H = 12
  d = H * 64
  qkv = torch.randn(B, N, 3 * d, device="cuda").bfloat16()
  q, k, v = qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:]
  for _ in range(3):
    o, lse = ops.attention_fwd(q, k, v, H)
  torch.cuda.synchronize()
  t0 = torch.cuda.Event(enable_timing=True)
  t1 = torch.cuda.Event(enable_timing=True)
  t0.record()
  o, lse = ops.attention_fwd(q, k, v, H)
  t1.record()
  torch.cuda.synchronize()
  tiles = B * H * ((N + 127) // 128)
  print(f"N={N} B={B}: {t0.elapsed_time(t1) * 1e3:.1f} us, {tiles} tiles, "
        f"{t0.elapsed_time(t1) * 1e3 / (tiles / 148):.2f} us per tile per SM")
  buf = (ctypes.c_longlong * (32 * 16))()
  lib = L.load()
  lib.bv_debug_attn_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
  n = lib.bv_debug_attn_timeline(buf, 32 * 16)
  vals = [buf[i] for i in range(n)]
  base = min(x for x in vals if x > 0)
  print("tile " + " ".join(f"{e[:11]:>11s}" for e in EV))
  for i in range(4, 12):
    row = vals[i * 16:(i + 1) * 16]
    print(f"{i:4d} " + " ".join(f"{(x - base) if x else -1:11d}" for x in row[:len(EV)]))


if __name__ == "__main__":
  main()
