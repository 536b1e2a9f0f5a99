"""GPU parity of the key-block streaming attention kernels (attention_stream.cu) against the fp64
reference attention: long sequences (config 5: 576 / 577 keys; ragged and very long cases) and the
short shapes of the resident kernels forced through the streaming path.  Tolerances as in
test_kernels_gpu.py (bf16 operands, un-normalised bf16 probabilities, fp32 accumulation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(x):
  return x.to(torch.bfloat16)


def _close(got, ref, tol):
  got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
  assert not torch.isnan(got).any()
  err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
  assert err <= tol, f"rel err {err:.3e} > {tol}"


def _ref_attention(q, k, v, H):
  B, Nq, d = q.shape
  Nk = k.shape[1]
  qh = q.reshape(B, Nq, H, 64).transpose(1, 2)
  kh = k.reshape(B, Nk, H, 64).transpose(1, 2)
  vh = v.reshape(B, Nk, H, 64).transpose(1, 2)
  s = qh @ kh.transpose(-1, -2) / 8.0
  return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Nq, d), torch.logsumexp(s, -1)


@pytest.fixture(scope="module")
def ops():
  from big_vision_b200 import lib, ops as _ops
  assert lib.load().bv_device_supported() == 1
  return _ops


SHAPES = [(2, 16, 576, 576), (2, 3, 577, 577), (1, 2, 300, 700), (3, 2, 1, 576), (1, 1, 1025, 130),
          (3, 2, 64, 64), (2, 12, 196, 196), (5, 3, 197, 197), (4, 2, 1, 196), (2, 1, 16, 16), (2, 2, 130, 7)]


@pytest.mark.parametrize("B,H,Nq,Nk", SHAPES)
def test_stream_forward(ops, monkeypatch, B, H, Nq, Nk):
  monkeypatch.setenv("BV_ATTN_FWD", "stream")
  g = torch.Generator().manual_seed(B * 1000 + Nq)
  d = H * 64
  qkv = _bf(torch.randn(B, max(Nq, Nk), 3 * d, generator=g))
  # a few large scores per row: the per-block maxima differ by far more than the bf16 range of P
  qkv[:, ::37, 0:d] *= 4.0
  q64, k64, v64 = qkv[:, :Nq, 0:d].double(), qkv[:, :Nk, d:2 * d].double(), qkv[:, :Nk, 2 * d:].double()
  o_ref, lse_ref = _ref_attention(q64, k64, v64, H)
  c = qkv.cuda()
  o, lse = ops.attention_fwd(c[:, :Nq, 0:d], c[:, :Nk, d:2 * d], c[:, :Nk, 2 * d:], H)
  torch.cuda.synchronize()
  _close(o, o_ref, 2 ** -6)
  _close(lse, lse_ref, 1e-5)


@pytest.mark.parametrize("sm", ["0", "109"])
@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 16, 576, 576), (2, 3, 577, 577), (1, 2, 300, 700), (3, 2, 1, 576),
                                       (2, 12, 196, 196), (3, 2, 64, 64), (2, 2, 130, 7)])
def test_stream_forward_variants(ops, monkeypatch, sm, B, H, Nq, Nk):
  """The non-default builds of the streaming forward stay correct: BV_ATTN_SM=0 (exponentials split
  between MUFU and the FMA-pipe polynomial) and 109 (P kept in tensor memory, tcgen05.st + A-from-TMEM
  P.V product)."""
  monkeypatch.setenv("BV_ATTN_FWD", "stream")
  monkeypatch.setenv("BV_ATTN_SM", sm)
  g = torch.Generator().manual_seed(B * 1000 + Nq + 7)
  d = H * 64
  qkv = _bf(torch.randn(B, max(Nq, Nk), 3 * d, generator=g))
  q64, k64, v64 = qkv[:, :Nq, 0:d].double(), qkv[:, :Nk, d:2 * d].double(), qkv[:, :Nk, 2 * d:].double()
  o_ref, lse_ref = _ref_attention(q64, k64, v64, H)
  c = qkv.cuda()
  o, lse = ops.attention_fwd(c[:, :Nq, 0:d], c[:, :Nk, d:2 * d], c[:, :Nk, 2 * d:], H)
  torch.cuda.synchronize()
  _close(o, o_ref, 2 ** -6)
  _close(lse, lse_ref, 1e-5)


def test_stream_forward_matches_resident_kernel_on_short_sequences(ops, monkeypatch):
  """Same inputs through both forward kernels: equal up to the bf16 rounding of the output."""
  g = torch.Generator().manual_seed(3)
  B, H, N = 8, 12, 196
  c = _bf(torch.randn(B, N, 3 * H * 64, generator=g)).cuda()
  d = H * 64
  monkeypatch.setenv("BV_ATTN_FWD", "resident")
  o1, l1 = ops.attention_fwd(c[:, :, 0:d], c[:, :, d:2 * d], c[:, :, 2 * d:], H)
  monkeypatch.setenv("BV_ATTN_FWD", "stream")
  o2, l2 = ops.attention_fwd(c[:, :, 0:d], c[:, :, d:2 * d], c[:, :, 2 * d:], H)
  _close(o2, o1, 2 ** -7)
  _close(l2, l1, 1e-5)


def test_stream_forward_rows_are_convex_combinations_at_config5_size(ops):
  """Size-independent property at the config-5 shape (576 keys, 16 heads): with v constant per
  column every output row equals that constant, whatever the scores are."""
  B, H, N = 32, 16, 576
  d = H * 64
  g = torch.Generator(device="cuda").manual_seed(1)
  q = torch.randn(B, N, d, generator=g, device="cuda").to(torch.bfloat16) * 3
  k = torch.randn(B, N, d, generator=g, device="cuda").to(torch.bfloat16)
  col = torch.randn(1, 1, d, generator=g, device="cuda").to(torch.bfloat16)
  o, _ = ops.attention_fwd(q, k, col.expand(B, N, d).contiguous(), H)
  assert (o.float() - col.float()).abs().max().item() <= 2 ** -7 * col.float().abs().max().item()


BWD_SHAPES = [(2, 16, 576, 576), (2, 3, 577, 577), (1, 2, 300, 700), (3, 2, 1, 576),
              (3, 2, 64, 64), (2, 12, 196, 196), (5, 3, 197, 197), (4, 2, 1, 196), (2, 2, 130, 7)]


@pytest.mark.parametrize("B,H,Nq,Nk", BWD_SHAPES)
def test_stream_backward(ops, monkeypatch, B, H, Nq, Nk):
  """Key-tile streaming backward (fp32 dQ accumulation by TMA reduce-add, delta pre-kernel) against
  autograd through the fp64 reference; the short shapes are forced through it with BV_ATTN_BWD."""
  monkeypatch.setenv("BV_ATTN_BWD", "stream")
  g = torch.Generator().manual_seed(B * 1000 + Nq)
  d = H * 64
  qkv = _bf(torch.randn(B, max(Nq, Nk), 3 * d, generator=g))
  do = _bf(torch.randn(B, Nq, d, generator=g))
  qr = qkv[:, :Nq, 0:d].double().requires_grad_(True)
  kr = qkv[:, :Nk, d:2 * d].double().requires_grad_(True)
  vr = qkv[:, :Nk, 2 * d:].double().requires_grad_(True)
  o_ref, _ = _ref_attention(qr, kr, vr, H)
  o_ref.backward(do.double())
  c = qkv.cuda()
  q, k, v = c[:, :Nq, 0:d], c[:, :Nk, d:2 * d], c[:, :Nk, 2 * d:]
  o, lse = ops.attention_fwd(q, k, v, H)
  dqkv = torch.zeros_like(c)
  cs = torch.ones(3, d, device="cuda")
  dq, dk, dv = ops.attention_bwd(do.cuda(), q, k, v, o, lse, H, dq=dqkv[:, :Nq, 0:d], dk=dqkv[:, :Nk, d:2 * d],
                                 dv=dqkv[:, :Nk, 2 * d:], dq_colsum=cs[0], dk_colsum=cs[1], dv_colsum=cs[2])
  torch.cuda.synchronize()
  _close(dq, qr.grad, 2 ** -5)
  _close(dk, kr.grad, 2 ** -5)
  _close(dv, vr.grad, 2 ** -5)
  for i, t in enumerate((dq, dk, dv)):      # fused bias gradients: column sums over the valid rows
    ref = 1 + t.double().sum((0, 1))
    assert (cs[i].double() - ref).abs().max().item() <= 1e-4 * (ref.abs().max().item() + 1)
  # rows past Nq / Nk of the destination buffers were not touched
  assert float(dqkv[:, Nq:, 0:d].abs().max() if Nq < dqkv.shape[1] else 0) == 0
  assert float(dqkv[:, Nk:, d:].abs().max() if Nk < dqkv.shape[1] else 0) == 0


def test_stream_backward_matches_resident_kernel(ops, monkeypatch):
  g = torch.Generator().manual_seed(4)
  B, H, N = 8, 12, 196
  d = H * 64
  c = _bf(torch.randn(B, N, 3 * d, generator=g)).cuda()
  do = _bf(torch.randn(B, N, d, generator=g)).cuda()
  q, k, v = c[:, :, 0:d], c[:, :, d:2 * d], c[:, :, 2 * d:]
  o, lse = ops.attention_fwd(q, k, v, H)
  monkeypatch.setenv("BV_ATTN_BWD", "resident")
  r = ops.attention_bwd(do, q, k, v, o, lse, H)
  monkeypatch.setenv("BV_ATTN_BWD", "stream")
  t = ops.attention_bwd(do, q, k, v, o, lse, H)
  for a, b in zip(t, r):
    _close(a, b, 2 ** -6)
