"""Bring-up aid: per-pair event timeline (SM cycles) of CTA 0 of the attention backward kernel.
  python tools/attn_bwd_timeline.py [N] [B]"""
import ctypes
import os
import sys

os.environ["BV_ATTN_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from big_vision_b200 import lib as L  # noqa: E402
from big_vision_b200 import ops  # noqa: E402

EV = ["c_sdp_full", "c_exp_done", "c_pds_empty", "-", "c_pds_arr", "m_sdp_iss", "m_grad_iss", "c_in_full",
      "m_pds_full", "m_acc_free", "m_grad_enter", "c_pds_arr_w7", "-", "tma_issued"]


def main():
  N = int(sys.argv[1]) if len(sys.argv) > 1 else 196
  B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
  H = 12
  d = H * 64
  qkv = torch.randn(B, N, 3 * d, device="cuda").bfloat16()
  q, k, v = qkv[:, :, 0:d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:]
  do = torch.randn(B, N, d, device="cuda").bfloat16()
  o, lse = ops.attention_fwd(q, k, v, H)
  for _ in range(2):
    ops.attention_bwd(do, q, k, v, o, lse, H)
  torch.cuda.synchronize()
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  ops.attention_bwd(do, q, k, v, o, lse, H)
  t1.record()
  torch.cuda.synchronize()
  items = B * H
  print(f"N={N} B={B}: {t0.elapsed_time(t1) * 1e3:.1f} us, {items} items, "
        f"{t0.elapsed_time(t1) * 1e3 / (items / 148):.2f} us per item per SM")
  cs = torch.zeros(3, d, device="cuda")
  for _ in range(2):
    ops.attention_bwd(do, q, k, v, o, lse, H, dq_colsum=cs[0], dk_colsum=cs[1], dv_colsum=cs[2])
  t0.record()
  ops.attention_bwd(do, q, k, v, o, lse, H, dq_colsum=cs[0], dk_colsum=cs[1], dv_colsum=cs[2])
  t1.record()
  torch.cuda.synchronize()
  print(f"   with fused bias gradients: {t0.elapsed_time(t1) * 1e3:.1f} us")
  buf = (ctypes.c_longlong * 512)()
  lib = L.load()
  lib.bv_debug_attn_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
  lib.bv_debug_attn_timeline(buf, 512)
  vals = [buf[i] for i in range(256, 512)]
  base = min(x for x in vals if x > 0)
  print("pair " + " ".join(f"{e[:11]:>11s}" for e in EV))
  for i in range(12):
    row = vals[i * 16:(i + 1) * 16]
    print(f"{i:4d} " + " ".join(f"{(x - base) if x else -1:11d}" for x in row[:len(EV)]))


if __name__ == "__main__":
  main()
