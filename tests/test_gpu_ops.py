"""Operator-level parity on a real B200: every kernel behind the C ABI against a plain torch fp32 restatement of
the reference op it replaces (floating-point kernels -> torch fp32 reference with the reference's rounding points).
Tolerances are stated per test; bf16 outputs are compared in units of bf16 ulps where that is meaningful."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from kandinsky import ops

    return ops


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def bf16_ulp_err(a, b):
    """max |a-b| in units of the bf16 spacing at max(|b|, 5 % of rms(b)): elements that are small only through
    cancellation are measured against the scale of the terms that produced them."""
    a, b = a.float(), b.float()
    floor = 0.05 * b.pow(2).mean().sqrt().clamp_min(1e-6)
    mag = torch.maximum(b.abs(), floor.expand_as(b))
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)
    return float(((a - b).abs() / ulp).max())


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 256, 192), (1024, 1792, 1792), (37, 1792, 3584),
                                   (1000, 64, 1792), (257, 128, 512), (4096, 7168, 1792), (2048, 1792, 7168)])
def test_gemm_store_bias(M, N, K):
    a, w = _rand((M, K), 1), _rand((N, K), 2, K ** -0.5)
    bias = _rand((N,), 3).float()
    out = _ops().linear(a, w, bias)
    ref = (a.float() @ w.float().t() + bias).to(torch.bfloat16)
    # fp32 accumulation order differs from torch's: allow 1 bf16 ulp on isolated elements
    assert rel_l2(out, ref) < 2e-3
    assert bf16_ulp_err(out, ref) <= 2.0
    out2 = _ops().linear(a, w, None)
    assert rel_l2(out2, (a.float() @ w.float().t()).to(torch.bfloat16)) < 2e-3


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 192, 512), (1536, 6144, 512), (257, 64, 1792)])
def test_gemm_f32_scores_are_unrounded(M, N, K):
    """K5_EPI_F32: the fp32 accumulators leave unrounded (scores of the VAE mid-block softmax, vae.py:343-359 + SDPA)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    a, w = _rand((M, K), 11), _rand((N, K), 12, K ** -0.5)
    out = _ops().linear(a, w, None, epilogue="f32")
    ref = a.float() @ w.float().t()
    assert out.dtype == torch.float32 and out.shape == (M, N)
    assert float((out - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))    # far below a bf16 ulp (4e-3)
    bias = _rand((N,), 13).float()
    out = _ops().linear(a, w, bias, epilogue="f32")
    assert float((out - (ref + bias)).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()) + 4.0)


def test_gemm_gelu():
    M, N, K = 640, 1024, 256
    a, w = _rand((M, K), 4), _rand((N, K), 5, K ** -0.5)
    out = _ops().linear(a, w, None, epilogue="gelu")
    h = (a.float() @ w.float().t()).to(torch.bfloat16)
    ref = torch.nn.functional.gelu(h.float()).to(torch.bfloat16)         # nn.py:356, exact erf
    assert rel_l2(out, ref) < 3e-3


def test_gemm_gate_residual_in_place():
    M, N, K = 517, 1792, 1792
    a, w = _rand((M, K), 6), _rand((N, K), 7, K ** -0.5)
    bias, gate = _rand((N,), 8).float(), _rand((N,), 9, dtype=torch.float32)
    x = _rand((M, N), 10)
    x0 = x.clone()
    out = _ops().linear(a, w, bias, epilogue="gate", resid=x, gate=gate, out=x)   # out aliases resid
    lin = (a.float() @ w.float().t() + bias).to(torch.bfloat16)
    ref = (x0.float() + gate * lin.float()).to(torch.bfloat16)                  # nn.py:30-33
    assert out.data_ptr() == x.data_ptr()
    assert rel_l2(out, ref) < 3e-3


def _heads_ref(a, w, bias, wq, wk, split, norm_cols, rope_cols, rope):
    y = (a.float() @ w.float().t() + bias).to(torch.bfloat16)
    M, N = y.shape
    yh = y.float().view(M, N // 64, 64)
    out = yh.clone()
    nh = norm_cols // 64
    if nh:
        wsel = torch.stack([wq if h * 64 < split else wk for h in range(nh)])          # [nh,64]
        n = yh[:, :nh]
        n = n * torch.rsqrt(n.pow(2).mean(-1, keepdim=True) + torch.finfo(torch.float32).eps) * wsel
        n = n.to(torch.bfloat16).float()                                               # nn.py:248-249
        rh = rope_cols // 64
        if rh:
            c, s = rope[..., 0][:, None, :], rope[..., 1][:, None, :]                  # [M,1,32]
            x0, x1 = n[:, :rh, 0::2], n[:, :rh, 1::2]
            r = torch.stack([c * x0 + (-s) * x1, s * x0 + c * x1], dim=-1).flatten(-2)   # nn.py:35-40
            n = torch.cat([r.to(torch.bfloat16).float(), n[:, rh:]], dim=1)
        out[:, :nh] = n
    return out.view(M, N).to(torch.bfloat16)


def test_gemm_heads_qkv_norm_rope():
    M, D = 333, 512
    a, w = _rand((M, D), 11), _rand((3 * D, D), 12, D ** -0.5)
    bias = _rand((3 * D,), 13).float()
    wq, wk = 1 + 0.1 * _rand((64,), 14, dtype=torch.float32), 1 + 0.1 * _rand((64,), 15, dtype=torch.float32)
    ang = _rand((M, 32), 16, 3.0, torch.float32)
    rope = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).contiguous()
    out = _ops().linear(a, w, bias, epilogue="heads", norm_w0=wq, norm_w1=wk, norm_split=D, norm_cols=2 * D,
                        rope_cols=2 * D, rope=rope)
    ref = _heads_ref(a, w, bias, wq, wk, D, 2 * D, 2 * D, rope)
    assert rel_l2(out, ref) < 4e-3
    # cross-attention flavour: k normed, v plain, no rope (nn.py:343-349)
    out2 = _ops().linear(a, w[: 2 * D], bias[: 2 * D].contiguous(), epilogue="heads", norm_w0=wk, norm_w1=wk, norm_split=D,
                         norm_cols=D, rope_cols=0)
    ref2 = _heads_ref(a, w[: 2 * D], bias[: 2 * D], wk, wk, D, D, 0, rope)
    assert rel_l2(out2, ref2) < 4e-3


def _attn_ref(q, k, v, heads):
    Sq, Sk = q.shape[0], k.shape[0]
    qh = q.float().view(Sq, heads, 64).transpose(0, 1)
    kh = k.float().view(Sk, heads, 64).transpose(0, 1)
    vh = v.float().view(Sk, heads, 64).transpose(0, 1)
    p = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1)
    return (p @ vh).transpose(0, 1).reshape(Sq, heads * 64)


@pytest.mark.parametrize("Sq,Sk,heads", [(256, 128, 1), (256, 256, 2), (512, 512, 4), (300, 37, 3), (64, 24, 28),
                                         (1000, 777, 2), (2048, 4096, 28)])
def test_attention_dense(Sq, Sk, heads):
    q, k, v = _rand((Sq, heads * 64), 20), _rand((Sk, heads * 64), 21), _rand((Sk, heads * 64), 22)
    out = _ops().attention(q, k, v, heads)
    ref = _attn_ref(q, k, v, heads)
    # P is rounded to bf16 before P.V (as in flash-attn) and the output is bf16
    assert rel_l2(out, ref) < 8e-3
    assert float((out.float() - ref).abs().max()) < 0.05


def test_attention_large_logits_lazy_rescale():
    """Row maxima that keep growing along the KV axis exercise the lazy O rescale path."""
    Sq, Sk, heads = 256, 1024, 2
    q = _rand((Sq, heads * 64), 23, 2.0)
    k = _rand((Sk, heads * 64), 24, 2.0) * torch.linspace(0.2, 3.0, Sk, device="cuda")[:, None].to(torch.bfloat16)
    v = _rand((Sk, heads * 64), 25)
    out = _ops().attention(q, k, v, heads)
    assert rel_l2(out, _attn_ref(q, k, v, heads)) < 1e-2


def _rms_heads(x, w=1.0):
    """per-head RMSNorm with a scalar weight (nn.py:246-250): every 64-column head slice gets norm <= 8 w"""
    x4 = x.float().reshape(x.shape[0], -1, 64)
    x4 = x4 * torch.rsqrt(x4.pow(2).mean(-1, keepdim=True) + torch.finfo(torch.float32).eps) * w
    return x4.reshape(x.shape).to(torch.bfloat16)


def _bound(wq=1.0, wk=1.0):
    return 8.0 * wq * 8.0 * wk * 0.125 * math.log2(math.e) * 1.02


@pytest.mark.parametrize("Sq,Sk,heads,wq,wk", [(256, 128, 1, 1.0, 1.0), (512, 512, 4, 1.0, 1.0), (300, 37, 3, 1.3, 0.7),
                                               (64, 24, 28, 1.0, 1.0), (1000, 777, 2, 2.2, 2.2), (2048, 4096, 28, 1.0, 1.0),
                                               (4000, 256, 4, 1.5, 1.5)])
def test_attention_bounded_matches_torch_and_general_kernel(Sq, Sk, heads, wq, wk):
    """Fixed-offset softmax (k5_attention_bounded) on RMS-normalised q / k with the bound the engine derives from the
    norm weights: same tolerance as the general kernel against torch fp32, and the two kernels agree with each other
    to bf16 rounding (they compute the same softmax; only the scale of P before normalisation differs)."""
    q, k = _rms_heads(_rand((Sq, heads * 64), 30), wq), _rms_heads(_rand((Sk, heads * 64), 31), wk)
    v = _rand((Sk, heads * 64), 32)
    assert _bound(wq, wk) <= 60.0
    out = _ops().attention(q, k, v, heads, score_bound=_bound(wq, wk))
    ref = _attn_ref(q, k, v, heads)
    assert rel_l2(out, ref) < 8e-3
    assert float((out.float() - ref).abs().max()) < 0.05
    gen = _ops().attention(q, k, v, heads)
    assert rel_l2(out, gen) < 6e-3


@pytest.mark.parametrize("Sq,Sk,split,split2,heads", [(512, 512, 128, 0, 2), (768, 1920, 1152, 0, 4), (300, 1024, 512, 0, 3),
                                                       (6144, 12288, 1536, 0, 28), (512, 768, 256, 384, 2), (300, 1920, 128, 1152, 3),
                                                       (6144, 12288, 1536, 6144, 28)])
def test_attention_split_by_key_rows_is_bit_identical(Sq, Sk, split, split2, heads):
    """k5_attention_bounded_split: key rows [0, split) in one launch (unnormalised fp32 partials of O and the row sums),
    the rest in a second launch that starts from them (with split2: a middle launch over [split, split2) that starts from
    the partials AND leaves them again) - what a shard rank does to start on its local K | V slab while the foreign slabs
    arrive.  No row maximum exists under a score bound, so the partials are additive and the launches must reproduce
    the single launch BIT FOR BIT (same accumulation order per row)."""
    q, k = _rms_heads(_rand((Sq, heads * 64), 40)), _rms_heads(_rand((Sk, heads * 64), 41))
    v = _rand((Sk, heads * 64), 42)
    one = _ops().attention(q, k, v, heads, score_bound=_bound())
    two = _ops().attention_split(q, k, v, heads, _bound(), split, split_row2=split2)
    assert torch.equal(one, two)
    assert rel_l2(two, _attn_ref(q, k, v, heads)) < 8e-3
    with pytest.raises(ValueError):
        _ops().attention_split(q, k, v, heads, _bound(), split + 64)
    with pytest.raises(ValueError):
        _ops().attention_split(q, k, v, heads, _bound(), split, split_row2=split)


def test_attention_bound_above_limit_selects_general_kernel():
    """A bound above 60 (or none) must not run the fixed-offset kernel: large logits still come out right."""
    Sq, Sk, heads = 256, 1024, 2
    q, k, v = _rand((Sq, heads * 64), 23, 2.0), _rand((Sk, heads * 64), 24, 3.0), _rand((Sk, heads * 64), 25)
    out = _ops().attention(q, k, v, heads, score_bound=500.0)
    assert rel_l2(out, _attn_ref(q, k, v, heads)) < 1e-2


def test_attention_strided_views_of_fused_qkv():
    S, heads = 640, 4
    D = heads * 64
    qkv = _rand((S, 3 * D), 26)
    out = _ops().attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], heads)
    ref = _attn_ref(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], heads)
    assert rel_l2(out, ref) < 8e-3


@pytest.mark.parametrize("S,D", [(1000, 1792), (37, 256), (8, 2048)])
def test_ln_modulate(S, D):
    x = _rand((S, D), 30, 2.0)
    scale, shift = _rand((D,), 31, 0.3, torch.float32), _rand((D,), 32, 0.3, torch.float32)
    out = _ops().ln_rows(x, scale, shift, plus_one=True)
    ref = (torch.nn.functional.layer_norm(x.float(), (D,)) * (scale + 1.0) + shift).to(torch.bfloat16)   # nn.py:25-28
    assert bf16_ulp_err(out, ref) <= 1.0
    assert (out != ref).float().mean() < 0.02


def test_invalid_arguments_raise_value_error():
    a, w = _rand((64, 100), 40), _rand((64, 100), 41)          # K not a multiple of 8
    with pytest.raises(ValueError):
        _ops().linear(a, w)


# ----------------------------------------------------------------------------- NABLA
def test_sta_mask_matches_oracle():
    from oracle import dit_oracle as O

    for (T, Hb, Wb, w) in [(4, 2, 2, (3, 3, 3)), (61, 4, 6, (11, 3, 3)), (5, 3, 2, (1, 1, 5))]:
        got = _ops().sta_mask(T, Hb, Wb, *w).bool().cpu()
        assert torch.equal(got, O.sta_mask(T, Hb, Wb, *w))          # bit-exact (integer work)


def _lists_to_mask(cnt, idx):
    h, nb, nk = idx.shape
    keep = torch.arange(nk, device=idx.device)[None, None, :] < cnt[..., None]
    # only the first kv_count entries of a row are defined (include/k5.h): park the rest on a spare column
    safe = torch.where(keep, idx.long(), torch.full_like(idx, nk, dtype=torch.long))
    assert int(safe.min()) >= 0 and int(safe.max()) <= nk
    return torch.zeros(h, nb, nk + 1, dtype=torch.bool, device=idx.device).scatter_(-1, safe, keep)[..., :nk]


@pytest.mark.parametrize("S,heads,P,use_sta", [(1024, 4, 0.6, True), (2048, 28, 0.9, False), (6144, 2, 0.5, True)])
def test_nabla_select_matches_oracle(S, heads, P, use_sta):
    from oracle import dit_oracle as O

    q, k = _rand((S, heads * 64), 50), _rand((S, heads * 64), 51)
    # concentrate the block-pooled scores so that the cumulative-mass threshold is selective
    q = (q.float() + 1.5 * _rand((S // 64, 1, heads * 64), 52).float().expand(-1, 64, -1).reshape(S, -1)).to(torch.bfloat16)
    k = (k.float() + 1.5 * _rand((S // 64, 1, heads * 64), 53).float().expand(-1, 64, -1).reshape(S, -1)).to(torch.bfloat16)
    nb = S // 64
    sta_cpu = O.sta_mask(nb // 4, 2, 2, 3, 3, 3) if use_sta else None
    sta = sta_cpu.to(torch.uint8).cuda() if use_sta else None
    cnt, idx = _ops().nabla_select(q, k, heads, P, sta)
    got = _lists_to_mask(cnt, idx).cpu()
    ref = O.nabla_block_mask(q.cpu().view(S, heads, 64), k.cpu().view(S, heads, 64), sta_cpu, P, "cuda")
    agree = float((got == ref).float().mean())
    dens = float(ref.float().mean())
    print(f"nabla_select S={S}: density {dens:.3f}, agreement {agree:.5f}")
    # the selection is a threshold on a cumulative sum: summation order may flip isolated borderline blocks
    assert agree > 0.998
    assert 0.02 < dens < 0.98
    # lists are ascending and counts consistent
    assert int(cnt.min()) >= 1
    first = idx[0, 0, : int(cnt[0, 0])]
    assert torch.all(first[1:] > first[:-1])


@pytest.mark.parametrize("S,heads,dens", [(512, 2, 0.5), (1024, 4, 0.2), (4096, 3, 0.08), (1088, 2, 0.4)])
def test_attention_block_sparse_matches_masked_dense(S, heads, dens):
    from oracle import dit_oracle as O

    nb = S // 64
    g = torch.Generator().manual_seed(60)
    mask = torch.rand(heads, nb, nb, generator=g) < dens
    mask |= torch.eye(nb, dtype=torch.bool)[None]                      # every row keeps at least its diagonal
    cnt = mask.sum(-1).to(torch.int32)
    idx = torch.argsort(mask.int(), dim=-1, descending=True, stable=True).to(torch.int32)
    q, k, v = _rand((S, heads * 64), 61), _rand((S, heads * 64), 62), _rand((S, heads * 64), 63)
    out = _ops().attention(q, k, v, heads, kv_count=cnt.cuda(), kv_index=idx.cuda().contiguous())
    ref = O.attention(q.cpu().view(S, heads, 64), k.cpu().view(S, heads, 64), v.cpu().view(S, heads, 64), "cuda", mask)
    assert rel_l2(out.cpu(), ref) < 8e-3
