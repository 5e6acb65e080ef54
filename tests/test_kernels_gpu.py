"""Operator-level parity: tcgen05 GEMM vs torch fp32 matmul of the same fp16 operands, tcgen05
FlashAttention vs the oracle's exact softmax attention (oracle/dit_oracle.py:softmax_attention,
restating attention.py:24-130).  Tolerances (floating point): GEMM fp32-accumulate
rel-L2 <= 3e-4 for fp32 output / 6e-4 for fp16 output, attention <= 2e-3."""
import math

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half()


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 128, 64, 128), (128, 256, 64, 256), (256, 256, 128, 256), (1560, 1536, 1536, 0), (1560, 4608, 1536, 256),
    (1560, 8960, 1536, 0), (1560, 1536, 8960, 128), (300, 384, 200, 128), (77, 96, 72, 128), (6240, 1536, 1536, 256),
    # tile widths that are not multiples of 32 (16-column epilogue chunks) / do not divide N
    (3120, 1536, 1536, 144), (3120, 4608, 1536, 192), (300, 384, 200, 144), (3120, 1536, 8960, 144),
    (3120, 1536, 8960, 0), (1000, 1000, 328, 192),
    # CTA-pair kernels (tcgen05 cta_group::2, 256 x BN tiles): width + 1000
    (3120, 4608, 1536, 1256), (300, 384, 200, 1128), (6240, 1536, 1536, 1256), (1560, 8960, 1536, 1192),
    (128, 256, 64, 1256), (3120, 1536, 8960, 1192), (2000, 1000, 328, 1224),
])
def test_linear_f32(M, N, K, bn):
    import b200dit
    a, w = _mk((M, K), 1).cuda(), _mk((N, K), 2, 1 / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).cuda()
    out = b200dit.linear(a, w, bias, "f32", bn)
    ref = a.float() @ w.float().t() + bias
    assert rel_l2(out, ref) < 3e-4, (M, N, K)


@pytest.mark.parametrize("epi,bn", [("f16", 0), ("gelu", 0), ("f16", 144), ("gelu", 192), ("f16", 192), ("f16", 1128),
                                    ("gelu", 1224), ("gelu", 1256)])
def test_linear_f16_epilogues(epi, bn):
    import b200dit
    M, N, K = 777, 640, 512
    a, w = _mk((M, K), 4).cuda(), _mk((N, K), 5, 1 / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=torch.Generator().manual_seed(6)).cuda()
    out = b200dit.linear(a, w, bias, epi, bn)
    ref = a.float() @ w.float().t() + bias
    if epi == "gelu":
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    assert out.dtype == torch.float16
    assert rel_l2(out.float(), ref) < 6e-4


@pytest.mark.parametrize("B,Lq,Lk,H,klens", [
    (1, 128, 128, 1, None), (1, 256, 256, 2, None), (1, 1560, 1560, 12, None), (2, 300, 512, 3, [77, 512]),
    (1, 200, 257, 2, None), (2, 1560, 512, 12, [512, 1]), (1, 130, 1000, 1, [999]),
    (2, 200, 257, 2, [257, 100]), (3, 70, 15, 1, None),      # odd key counts: every item's V^T starts on a multiple of 8
    # >= 1024 keys: the two-tiles-per-CTA kernel (ragged key counts, a lone last tile, partial second tile)
    (2, 300, 1104, 2, [1104, 1030]), (1, 200, 1560, 1, None), (1, 128, 1152, 1, None), (1, 640, 1280, 3, [1025]),
    # tail split: 312 / 624 tiles on 296 co-resident slots leave 16 / 32 tiles for the last wave, which are cut
    # into 4 (self-attention, 25 key steps) or 3 (cross-attention, 6+ key steps) parts along the key axis
    (2, 1560, 1560, 12, None), (4, 1560, 512, 12, [512, 400, 384, 512]),
])
def test_flash_attention(B, Lq, Lk, H, klens):
    import b200dit
    from oracle import dit_oracle as O
    q, k, v = _mk((B, Lq, H, 128), 7).cuda(), _mk((B, Lk, H, 128), 8).cuda(), _mk((B, Lk, H, 128), 9).cuda()
    out = b200dit.flash_attention(q, k, v, k_lens=torch.tensor(klens) if klens else None)
    assert out.shape == q.shape and out.dtype == q.dtype
    for b in range(B):
        ref = O.softmax_attention(q[b].cpu().float(), k[b].cpu().float(), v[b].cpu().float(), klens[b] if klens else None)
        assert rel_l2(out[b].cpu().float(), ref) < 2e-3, (b,)


def test_flash_attention_tail_split_matches_unsplit(monkeypatch):
    """B200_ATTN_SPLIT=0 runs every tile in one CTA; the split path must agree to fp16 rounding of the output."""
    import b200dit
    B, L, H = 2, 1560, 12
    q, k, v = _mk((B, L, H, 128), 13).cuda(), _mk((B, L, H, 128), 14).cuda(), _mk((B, L, H, 128), 15).cuda()
    a = b200dit.flash_attention(q, k, v)
    monkeypatch.setenv("B200_ATTN_SPLIT", "0")
    b = b200dit.flash_attention(q, k, v)
    monkeypatch.delenv("B200_ATTN_SPLIT")
    c = b200dit.flash_attention(q, k, v)
    assert torch.equal(a, c)                                   # deterministic
    assert rel_l2(a.float(), b.float()) < 3e-4
    assert (a.float() - b.float()).abs().max() < 4e-3


@pytest.mark.parametrize("B,Lq,Lk,H,klens", [
    (1, 2304, 2304, 2, None),                 # long enough for the two-tiles-per-CTA kernel by default
    (2, 2100, 2500, 1, [2500, 2049]),         # ragged keys, partial second tile
])
def test_flash_attention_long_sequences(B, Lq, Lk, H, klens):
    test_flash_attention(B, Lq, Lk, H, klens)


@pytest.mark.parametrize("B,Lq,Lk,H,klens", [
    (1, 128, 128, 1, None), (2, 300, 512, 3, [77, 512]), (1, 1560, 1560, 12, None), (3, 70, 15, 1, None),
    (2, 200, 257, 2, [257, 100]), (1, 640, 1280, 3, [1025]),
])
def test_flash_attention_two_tile_kernel_forced(monkeypatch, B, Lq, Lk, H, klens):
    """B200_ATTN_PAIR=1 sends every shape through the long-sequence kernel (lone / partial second tiles, short and
    ragged key counts)."""
    monkeypatch.setenv("B200_ATTN_PAIR", "1")
    test_flash_attention(B, Lq, Lk, H, klens)


def test_flash_attention_T21_L32760():
    """The T = 21 sequence length of configs 3 / 5: Lq = Lk = 32 760 (255 full query tiles + a 120-row tail, 511 full
    key steps + a 56-key tail) on two heads through the long-sequence kernel, every row against the query-chunked
    CPU oracle (attention.py:24-130 semantics)."""
    import b200dit
    from oracle import dit_oracle as O
    L, H = 32760, 2
    q, k, v = _mk((1, L, H, 128), 21).cuda(), _mk((1, L, H, 128), 22).cuda(), _mk((1, L, H, 128), 23).cuda()
    out = b200dit.flash_attention(q, k, v)
    ref = O.softmax_attention(q[0].cpu().float(), k[0].cpu().float(), v[0].cpu().float(), None)
    assert rel_l2(out[0].cpu().float(), ref) < 2e-3
    assert rel_l2(out[0, -120:].cpu().float(), ref[-120:]) < 2e-3          # the partial last tile
    # ragged: the second item's keys stop inside a tile
    out2 = b200dit.flash_attention(q[:, :4096], k, v, k_lens=torch.tensor([20001]))
    ref2 = O.softmax_attention(q[0, :4096].cpu().float(), k[0].cpu().float(), v[0].cpu().float(), 20001)
    assert rel_l2(out2[0].cpu().float(), ref2) < 2e-3


def test_flash_attention_large_logits():
    """rows whose running max keeps growing exercise the thresholded O rescale path"""
    import b200dit
    from oracle import dit_oracle as O
    B, L, H = 1, 640, 1
    q, k, v = _mk((B, L, H, 128), 10, 3.0).cuda(), _mk((B, L, H, 128), 11, 3.0).cuda(), _mk((B, L, H, 128), 12).cuda()
    k[0, :, 0, :] *= torch.linspace(0.2, 2.0, L, device="cuda").half()[:, None]     # later keys score higher
    out = b200dit.flash_attention(q, k, v)
    ref = O.softmax_attention(q[0].cpu().float(), k[0].cpu().float(), v[0].cpu().float(), None)
    assert rel_l2(out[0].cpu().float(), ref) < 3e-3


def _attn_grads_reference(q, k, v, do, klens):
    """fp32 autograd through the exact masked softmax attention (attention.py:24-130 semantics) on the CPU."""
    q, k, v = (u.cpu().float().clone().requires_grad_(True) for u in (q, k, v))
    outs = []
    for b in range(q.shape[0]):
        s = torch.einsum("qhd,khd->hqk", q[b], k[b]) / 128 ** 0.5
        if klens is not None and klens[b] < k.shape[1]:
            s = s.masked_fill(torch.arange(k.shape[1])[None, None, :] >= klens[b], float("-inf"))
        outs.append(torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v[b]))
    torch.stack(outs).backward(do.cpu().float())
    return q.grad, k.grad, v.grad


@pytest.mark.parametrize("B,Lq,Lk,H,klens", [
    (1, 128, 128, 1, None), (1, 64, 128, 1, None), (1, 256, 256, 2, None),
    (2, 304, 512, 3, [77, 512]),              # cross-attention shape: masked keys, whole key tiles past klen
    (1, 200, 264, 2, [257]),                  # partial query and key tiles
    (3, 72, 16, 1, [16, 15, 2]),              # one key tile mostly empty; two valid keys
    (1, 1560, 1560, 12, None),                # self-attention of the bench latent: 13 key tiles x 25 query steps
    (2, 1560, 512, 12, [512, 300]),
])
def test_flash_attention_backward(B, Lq, Lk, H, klens):
    """b200_flash_attention_backward (the fused tcgen05 kernel of csrc/attn_bwd_tc.cu) against fp32 autograd through
    the exact attention: dq, dk, dv rel-L2 <= 2e-3 per item; keys past k_lens get exactly zero gradients."""
    import b200dit
    q, k, v = _mk((B, Lq, H, 128), 31).cuda(), _mk((B, Lk, H, 128), 32).cuda(), _mk((B, Lk, H, 128), 33).cuda()
    do = _mk((B, Lq, H, 128), 34).cuda()
    dq, dk, dv = b200dit.flash_attention_backward(q, k, v, do, k_lens=klens)
    rq, rk, rv = _attn_grads_reference(q, k, v, do, klens)
    for b in range(B):
        n = klens[b] if klens else Lk
        assert rel_l2(dq[b].cpu(), rq[b]) < 2e-3, ("dq", b, rel_l2(dq[b].cpu(), rq[b]))
        assert rel_l2(dk[b, :n].cpu(), rk[b, :n]) < 2e-3, ("dk", b, rel_l2(dk[b, :n].cpu(), rk[b, :n]))
        assert rel_l2(dv[b, :n].cpu().float(), rv[b, :n]) < 2e-3, ("dv", b, rel_l2(dv[b, :n].cpu().float(), rv[b, :n]))
        if n < Lk:
            assert float(dk[b, n:].abs().max()) == 0.0 and float(dv[b, n:].float().abs().max()) == 0.0
    dq2, dk2, dv2 = b200dit.flash_attention_backward(q, k, v, do, k_lens=klens)
    assert torch.equal(dk, dk2) and torch.equal(dv, dv2)               # dK / dV have one writer: bit-reproducible
    assert rel_l2(dq2.cpu(), dq.cpu()) < 1e-5                          # dQ sums key tiles with fp32 L2 atomics
