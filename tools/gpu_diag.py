"""Kernel-by-kernel diagnostics against plain torch fp32 math (development aid).

  python tools/gpu_diag.py            # run every check, each in its own subprocess + timeout
  python tools/gpu_diag.py gemm_fwd   # run one check in-process

Prints max-abs / relative errors so a single GPU session says which kernel is wrong and
how (layout vs. scale vs. garbage).  The pytest suite (tests/, -m gpu) is the gate; this
script is for bring-up.
"""
import math
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _err(name, got, ref, tol):
  import torch
  got = got.float()
  ref = ref.float()
  diff = (got - ref).abs()
  denom = ref.abs().max().item() + 1e-12
  rel = diff.max().item() / denom
  bad = torch.isnan(got).any().item() or torch.isinf(got).any().item()
  status = "OK " if (rel <= tol and not bad) else "BAD"
  print(f"  [{status}] {name}: max|d|={diff.max().item():.4e} rel={rel:.4e} "
        f"(ref max {denom:.3e}) nan/inf={bad}", flush=True)
  if status == "BAD":
    idx = diff.flatten().argmax().item()
    print(f"        worst flat index {idx} of shape {tuple(got.shape)}: got "
          f"{got.flatten()[idx].item():.5f} ref {ref.flatten()[idx].item():.5f}")
    # coarse map of where errors are: by row-block / col-block
    if got.dim() == 2:
      R, C = got.shape
      rb, cb = max(1, R // 8), max(1, C // 8)
      m = diff[: rb * 8, : cb * 8].reshape(8, rb, 8, cb).amax(dim=(1, 3))
      print("        block max-abs error map (8x8):")
      for r in range(8):
        print("         " + " ".join(f"{m[r, c].item():9.2e}" for c in range(8)))
  return status == "OK "


def check_gemm(kind):
  import torch
  from big_vision_b200 import ops, lib as L
  torch.manual_seed(0)
  dev = "cuda"
  ok = True
  shapes = [(256, 256, 128), (300, 768, 192), (1000, 2304, 768), (4096, 3072, 768), (130, 1000, 512)]
  if kind == "wgrad":
    shapes = [(768, 768, 1024), (768, 3072, 4000), (3072, 768, 1568), (384, 1000, 333 * 8)]
  for (M, N, K) in shapes:
    a32 = torch.randn(M, K, device=dev) * 0.5
    b32 = torch.randn(N, K, device=dev) * 0.5
    a = a32.bfloat16()
    b = b32.bfloat16()
    ref = a.float() @ b.float().t()
    if kind == "fwd":        # A K-major [M,K], B stored [K,N] (MN-major)
      bias = torch.randn(N, device=dev)
      out = ops.gemm(a, b.t().contiguous(), b_mn=True, bias=bias)
      ok &= _err(f"fwd(K,MN)+bias {M}x{N}x{K}", out, ref + bias, 1e-2)
      resid = torch.randn(M, N, device=dev).bfloat16()
      out = ops.gemm(a, b.t().contiguous(), b_mn=True, bias=bias, aux=resid, epilogue=L.EPI_BIAS_RESID)
      ok &= _err(f"fwd+bias+resid {M}x{N}x{K}", out, (ref + bias).bfloat16().float() + resid.float(), 1e-2)
      act, pre = ops.gemm(a, b.t().contiguous(), b_mn=True, bias=bias, epilogue=L.EPI_BIAS_GELU)
      pre_ref = (ref + bias)
      ok &= _err(f"fwd+bias+gelu(pre) {M}x{N}x{K}", pre, pre_ref, 1e-2)
      ok &= _err(f"fwd+bias+gelu(act) {M}x{N}x{K}", act,
                 torch.nn.functional.gelu(pre.float(), approximate="tanh"), 1e-2)
      pe = torch.randn(50, N, device=dev).bfloat16()
      out = ops.gemm(a, b.t().contiguous(), b_mn=True, bias=bias, aux=pe, aux_row_mod=50,
                     epilogue=L.EPI_BIAS_RESID)
      rows = torch.arange(M, device=dev) % 50
      ok &= _err(f"fwd+bias+posemb(mod) {M}x{N}x{K}", out, (ref + bias).bfloat16().float() + pe.float()[rows], 1e-2)
    elif kind == "dgrad":    # both K-major
      out = ops.gemm(a, b)
      ok &= _err(f"dgrad(K,K) {M}x{N}x{K}", out, ref, 1e-2)
      out32 = ops.gemm(a, b, out_dtype=torch.float32)
      ok &= _err(f"dgrad(K,K) f32out {M}x{N}x{K}", out32, ref, 2e-3)
      pre = torch.randn(M, N, device=dev).bfloat16()
      out = ops.gemm(a, b, aux=pre, epilogue=L.EPI_DGELU)
      x = pre.float().requires_grad_(True)
      torch.nn.functional.gelu(x, approximate="tanh").sum().backward()
      ok &= _err(f"dgrad+dgelu {M}x{N}x{K}", out, ref * x.grad, 1e-2)
      for bn in (128, 256):
        out = ops.gemm(a, b, block_n=bn)
        ok &= _err(f"dgrad(K,K) block_n={bn} {M}x{N}x{K}", out, ref, 1e-2)
    elif kind == "wgrad":    # both MN-major, fp32 reduce-add output, split-K
      at = a.t().contiguous()   # [K, M]
      bt = b.t().contiguous()   # [K, N]
      out = torch.zeros(M, N, device=dev)
      ops.gemm(at, bt, a_mn=True, b_mn=True, out=out, reduce_out=True)
      ok &= _err(f"wgrad(MN,MN) splitK auto {M}x{N}x{K}", out, ref, 2e-3)
      ops.gemm(at, bt, a_mn=True, b_mn=True, out=out, reduce_out=True, splits=1)
      ok &= _err(f"wgrad accumulate twice {M}x{N}x{K}", out, 2 * ref, 2e-3)
      out = torch.zeros(M, N, device=dev)
      ops.gemm(at, b, a_mn=True, b_mn=False, out=out, reduce_out=True, splits=3)
      ok &= _err(f"(MN,K) splits=3 {M}x{N}x{K}", out, ref, 2e-3)
  return ok


def check_gemm_fwd():
  return check_gemm("fwd")


def check_gemm_dgrad():
  return check_gemm("dgrad")


def check_gemm_wgrad():
  return check_gemm("wgrad")


def check_layernorm():
  import torch
  from big_vision_b200 import ops
  torch.manual_seed(0)
  ok = True
  for rows, d in [(1000, 768), (77, 384), (513, 1024), (64, 1152)]:
    x = (torch.randn(rows, d, device="cuda") * 2 + 0.5).bfloat16()
    g = torch.randn(d, device="cuda") * 0.2 + 1
    b = torch.randn(d, device="cuda") * 0.1
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xr = x.float().requires_grad_(True)
    gr = g.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (d,), gr, br, eps=1e-6)
    ok &= _err(f"ln fwd {rows}x{d}", y, yr, 1e-2)
    y32, _, _ = ops.layernorm_fwd(x, g, b, out_dtype=torch.float32)
    ok &= _err(f"ln fwd f32 out {rows}x{d}", y32, yr, 1e-4)
    dy = torch.randn(rows, d, device="cuda").bfloat16()
    dres = torch.randn(rows, d, device="cuda").bfloat16()
    yr.backward(dy.float())
    dscale = torch.zeros(d, device="cuda")
    dbias = torch.zeros(d, device="cuda")
    dcs = torch.zeros(d, device="cuda")
    dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dres=dres, dscale=dscale, dbias=dbias, dx_colsum=dcs)
    ok &= _err(f"ln bwd dx {rows}x{d}", dx, xr.grad + dres.float(), 1e-2)
    ok &= _err(f"ln bwd dscale {rows}x{d}", dscale, gr.grad, 1e-3)
    ok &= _err(f"ln bwd dbias {rows}x{d}", dbias, br.grad, 1e-3)
    ok &= _err(f"ln bwd colsum {rows}x{d}", dcs, dx.float().sum(0), 1e-3)
  return ok


def _ref_attn(q, k, v, heads, scale):
  import torch
  B, Nq, _ = q.shape
  Nk = k.shape[1]
  qh = q.float().reshape(B, Nq, heads, 64).transpose(1, 2)
  kh = k.float().reshape(B, Nk, heads, 64).transpose(1, 2)
  vh = v.float().reshape(B, Nk, heads, 64).transpose(1, 2)
  s = (qh @ kh.transpose(-1, -2)) * scale
  p = torch.softmax(s, dim=-1)
  o = (p @ vh).transpose(1, 2).reshape(B, Nq, heads * 64)
  lse = torch.logsumexp(s, dim=-1)
  return o, lse


def check_attention_fwd():
  import torch
  from big_vision_b200 import ops
  torch.manual_seed(0)
  ok = True
  for (B, H, Nq, Nk) in [(3, 2, 64, 64), (2, 12, 196, 196), (5, 3, 197, 197), (4, 2, 1, 196),
                         (300, 12, 196, 196), (2, 2, 256, 256)]:
    d = H * 64
    qkv = (torch.randn(B, max(Nq, Nk), 3 * d, device="cuda")).bfloat16()
    q = qkv[:, :Nq, 0:d]
    k = qkv[:, :Nk, d:2 * d]
    v = qkv[:, :Nk, 2 * d:3 * d]
    o, lse = ops.attention_fwd(q, k, v, H)
    o_ref, lse_ref = _ref_attn(q, k, v, H, 0.125)
    ok &= _err(f"attn fwd o B{B} H{H} {Nq}x{Nk}", o.reshape(B * Nq, d), o_ref.reshape(B * Nq, d), 2e-2)
    ok &= _err(f"attn fwd lse B{B} H{H} {Nq}x{Nk}", lse.reshape(B * H, Nq), lse_ref.reshape(B * H, Nq), 2e-3)
  return ok


def check_attention_bwd():
  import torch
  from big_vision_b200 import ops
  torch.manual_seed(0)
  ok = True
  for (B, H, Nq, Nk) in [(3, 2, 64, 64), (2, 12, 196, 196), (5, 3, 197, 197), (4, 2, 1, 196),
                         (160, 12, 196, 196), (2, 2, 256, 256)]:
    d = H * 64
    qkv = (torch.randn(B, max(Nq, Nk), 3 * d, device="cuda")).bfloat16()
    q = qkv[:, :Nq, 0:d]
    k = qkv[:, :Nk, d:2 * d]
    v = qkv[:, :Nk, 2 * d:3 * d]
    o, lse = ops.attention_fwd(q, k, v, H)
    do = torch.randn(B, Nq, d, device="cuda").bfloat16()
    qr = q.float().detach().requires_grad_(True)
    kr = k.float().detach().requires_grad_(True)
    vr = v.float().detach().requires_grad_(True)
    o_ref, _ = _ref_attn(qr, kr, vr, H, 0.125)
    o_ref.backward(do.float())
    dqkv = torch.zeros(B, max(Nq, Nk), 3 * d, device="cuda", dtype=torch.bfloat16)
    dq, dk, dv = ops.attention_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :Nq, 0:d],
                                   dk=dqkv[:, :Nk, d:2 * d], dv=dqkv[:, :Nk, 2 * d:3 * d])
    ok &= _err(f"attn bwd dq B{B} H{H} {Nq}x{Nk}", dq.reshape(B * Nq, d), qr.grad.reshape(B * Nq, d), 3e-2)
    ok &= _err(f"attn bwd dk B{B} H{H} {Nq}x{Nk}", dk.reshape(B * Nk, d), kr.grad.reshape(B * Nk, d), 3e-2)
    ok &= _err(f"attn bwd dv B{B} H{H} {Nq}x{Nk}", dv.reshape(B * Nk, d), vr.grad.reshape(B * Nk, d), 3e-2)
  return ok


def check_elementwise():
  import torch
  from big_vision_b200 import ops
  torch.manual_seed(0)
  ok = True
  # patchify
  n, H, W, C, P = 3, 32, 48, 3, 16
  img = torch.rand(n, H, W, C, device="cuda") * 2 - 1
  pt = ops.patchify(img, P)
  ref = img.reshape(n, H // P, P, W // P, P, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, P * P * C)
  ok &= _err("patchify", pt, ref, 5e-3)
  # embed
  ids = torch.randint(0, 100, (5, 7), device="cuda", dtype=torch.int32)
  table = torch.randn(100, 64, device="cuda")
  pos = torch.randn(7, 64, device="cuda")
  e = ops.embed_fwd(ids, table, pos, out_dtype=torch.float32)
  ok &= _err("embed fwd", e, (table[ids.long()] + pos[None]).reshape(-1, 64), 1e-6)
  dy = torch.randn(35, 64, device="cuda")
  dt = torch.zeros_like(table)
  dp = torch.zeros_like(pos)
  ops.embed_bwd(ids, dy, dt, dp)
  dt_ref = torch.zeros_like(table).index_add_(0, ids.flatten().long(), dy)
  ok &= _err("embed bwd table", dt, dt_ref, 1e-5)
  ok &= _err("embed bwd pos", dp, dy.reshape(5, 7, 64).sum(0), 1e-5)
  # colsum
  x = torch.randn(1234, 776, device="cuda").bfloat16()
  out = torch.zeros(776, device="cuda")
  ops.colsum(x, out)
  ok &= _err("colsum", out, x.float().sum(0), 1e-4)
  # l2norm
  x = torch.randn(33, 768, device="cuda")
  z, nrm = ops.l2norm_fwd(x)
  xr = x.clone().requires_grad_(True)
  zr = xr / (xr.norm(dim=1, keepdim=True) + 1e-8)
  ok &= _err("l2norm fwd", z, zr, 1e-5)
  dz = torch.randn_like(z)
  zr.backward(dz)
  ok &= _err("l2norm bwd", ops.l2norm_bwd(dz, z, nrm), xr.grad, 1e-4)
  # pool
  x = torch.randn(4 * 10, 64, device="cuda")
  ok &= _err("pool gap", ops.pool_fwd(x, 4, 10, 0), x.reshape(4, 10, 64).mean(1), 1e-5)
  ok &= _err("pool tok", ops.pool_fwd(x, 4, 10, 1, tok=9), x.reshape(4, 10, 64)[:, 9], 1e-6)
  dy = torch.randn(4, 64, device="cuda")
  ref = torch.zeros(4, 10, 64, device="cuda")
  ref[:, 9] = dy
  ok &= _err("pool tok bwd", ops.pool_bwd(dy, 4, 10, 1, tok=9, dx_dtype=torch.float32), ref.reshape(40, 64), 1e-6)
  ok &= _err("pool gap bwd", ops.pool_bwd(dy, 4, 10, 0, dx_dtype=torch.float32),
             (dy[:, None] / 10).expand(4, 10, 64).reshape(40, 64), 1e-6)
  # cast, tanh
  x = torch.randn(1001, device="cuda")
  y = torch.empty(1001, device="cuda", dtype=torch.bfloat16)
  ok &= _err("cast", ops.cast(x, y), x.bfloat16(), 1e-6)
  ok &= _err("tanh", ops.tanh_fwd(x), torch.tanh(x), 1e-5)
  ok &= _err("gelu", ops.gelu_fwd(x), torch.nn.functional.gelu(x, approximate="tanh"), 1e-5)
  ok &= _err("broadcast", ops.broadcast_row(x[:64].reshape(1, 64), 9), x[:64].expand(9, 64), 1e-6)
  xt = torch.randn(3 * 196, 64, device="cuda").bfloat16()
  tt = ops.transpose_tokens(xt, 3, 196, 64)
  ref = torch.zeros(3, 64, 200, device="cuda")
  ref[:, :, :196] = xt.float().reshape(3, 196, 64).transpose(1, 2)
  ok &= _err("transpose_tokens", tt, ref.reshape(3 * 64, 200), 1e-6)
  return ok


def check_loss():
  import torch
  from big_vision_b200 import ops
  torch.manual_seed(0)
  ok = True
  n, B, D, off = 64, 256, 768, 128
  zi = torch.nn.functional.normalize(torch.randn(n, D, device="cuda"), dim=1)
  zt = torch.nn.functional.normalize(torch.randn(B, D, device="cuda"), dim=1)
  zt[off:off + n] = torch.nn.functional.normalize(zi + 0.5 * zt[off:off + n], dim=1)
  dots = (zi @ zt.t()).contiguous()
  tp = torch.tensor([math.log(10.0)], device="cuda")
  bp = torch.tensor([-10.0], device="cuda")
  loss = torch.zeros(1, device="cuda")
  dt = torch.zeros(1, device="cuda")
  db = torch.zeros(1, device="cuda")
  G = ops.siglip_loss(dots, off, tp, bp, B, loss, dt, db)
  dr = dots.clone().requires_grad_(True)
  tr = tp.clone().requires_grad_(True)
  br = bp.clone().requires_grad_(True)
  x = dr * tr.exp() + br
  m = -torch.ones(n, B, device="cuda")
  m[torch.arange(n), off + torch.arange(n)] = 1
  l = -(torch.nn.functional.logsigmoid(m * x)).sum() / B
  l.backward()
  ok &= _err("siglip loss", loss, l.detach().reshape(1), 1e-5)
  ok &= _err("siglip G", G, dr.grad, 1e-2)
  ok &= _err("siglip dt", dt, tr.grad, 1e-4)
  ok &= _err("siglip db", db, br.grad, 1e-4)
  for fn, name in ((ops.sigmoid_xent, "sigmoid_xent"), (ops.softmax_xent, "softmax_xent")):
    lg = torch.randn(37, 1000, device="cuda") * 3
    lab = torch.nn.functional.one_hot(torch.randint(0, 1000, (37,), device="cuda"), 1000).float()
    lab = 0.9 * lab + 0.1 * lab.roll(1, 0)
    loss = torch.zeros(1, device="cuda")
    dl = fn(lg, lab, loss)
    lr = lg.clone().requires_grad_(True)
    if name == "sigmoid_xent":
      ref = -(lab * torch.nn.functional.logsigmoid(lr) + (1 - lab) * torch.nn.functional.logsigmoid(-lr)).sum(-1).mean()
    else:
      ref = -(lab * torch.log_softmax(lr, -1)).sum(-1).mean()
    ref.backward()
    ok &= _err(name, loss, ref.detach().reshape(1), 1e-5)
    ok &= _err(name + " grad", dl, lr.grad, 1e-4)
  return ok


def check_adam():
  import torch
  from big_vision_b200 import ops
  torch.manual_seed(0)
  ok = True
  n = 4096 * 3 + 4
  for mu_dt in (torch.float32, torch.bfloat16):
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 3
    mu = torch.zeros(n, device="cuda", dtype=mu_dt)
    nu = torch.zeros(n, device="cuda")
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    gsq = torch.zeros(1, device="cuda")
    ops.sumsq(g, gsq)
    ok &= _err("sumsq", gsq, (g * g).sum().reshape(1), 1e-5)
    pr, mr, vr = p.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    usq = torch.zeros(1, device="cuda")
    psq = torch.zeros(1, device="cuda")
    for step in (1, 2, 3):
      ops.adam_step(p, g, mu, nu, p16, lr_eff=1e-3, b1=0.9, b2=0.95, eps=1e-8, wd_eff=1e-4,
                    step=step, clip_norm=1.0, gnorm_sq=gsq, upd_sq=usq, param_sq=psq)
      gc = g * (1.0 / g.norm())
      mr = 0.9 * mr + 0.1 * gc
      vr = 0.95 * vr + 0.05 * gc * gc
      d = (mr / (1 - 0.9 ** step)) / ((vr / (1 - 0.95 ** step)).sqrt() + 1e-8)
      pr = pr - (1e-3 * d + 1e-4 * pr)
      if mu_dt == torch.bfloat16:
        mr = mr.bfloat16().float()
    ok &= _err(f"adam p mu={mu_dt}", p, pr, 1e-5)
    ok &= _err(f"adam p16 mu={mu_dt}", p16, pr.bfloat16(), 1e-6)
    ok &= _err(f"adam nu mu={mu_dt}", nu, vr, 1e-5)
  return ok


CHECKS = {
    "elementwise": check_elementwise,
    "layernorm": check_layernorm,
    "loss": check_loss,
    "adam": check_adam,
    "gemm_dgrad": check_gemm_dgrad,
    "gemm_fwd": check_gemm_fwd,
    "gemm_wgrad": check_gemm_wgrad,
    "attention_fwd": check_attention_fwd,
    "attention_bwd": check_attention_bwd,
}


def main():
  if len(sys.argv) == 2 and sys.argv[1] in CHECKS:
    ok = CHECKS[sys.argv[1]]()
    import torch
    torch.cuda.synchronize()
    print("RESULT", sys.argv[1], "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
  names = sys.argv[1:] or list(CHECKS)
  summary = {}
  for name in names:
    print(f"=== {name} ===", flush=True)
    t0 = time.time()
    try:
      r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240,
                         capture_output=True, text=True)
      print(r.stdout[-6000:])
      if r.returncode != 0:
        print(r.stderr[-3000:])
      summary[name] = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
    except subprocess.TimeoutExpired as e:
      print((e.stdout or b"").decode("utf-8", "replace")[-3000:] if isinstance(e.stdout, bytes) else (e.stdout or "")[-3000:])
      summary[name] = "TIMEOUT (hang?)"
    print(f"--- {name}: {summary[name]} in {time.time() - t0:.1f}s", flush=True)
  print("SUMMARY", summary)


if __name__ == "__main__":
  main()
