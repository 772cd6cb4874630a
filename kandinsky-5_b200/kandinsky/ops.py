"""Operator-level wrappers over the C ABI: thin argument marshalling for torch CUDA tensors.
These mirror the reference's op call sites (nn.Linear + following elementwise op, flash_attn_func,
apply_scale_shift_norm, nablaT_v2) and are what the parity tests drive."""
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _bf16(t):
    assert t.is_cuda and t.dtype == torch.bfloat16 and t.stride(-1) == 1, "expected a CUDA bf16 row-major tensor"
    return t


def _f32(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
    return t


def linear(a, w, bias=None, epilogue="store", resid=None, gate=None, norm_w0=None, norm_w1=None, norm_split=0,
           norm_cols=0, rope_cols=0, rope=None, out=None):
    """out[M,N] = epilogue(a[M,K] @ w[N,K]^T).  bias / gate / norm weights float32; rope float32 [M,32,2]."""
    a, w = _bf16(a), _bf16(w)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if epilogue == "f32" else torch.bfloat16)
    epi = {"store": _lib.EPI_STORE, "gelu": _lib.EPI_GELU, "gate": _lib.EPI_GATE, "heads": _lib.EPI_HEADS,
           "f32": _lib.EPI_F32}[epilogue]
    check(lib().k5_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K, epi, ptr(out), out.stride(0),
                             ptr(_f32(bias)), ptr(resid), 0 if resid is None else resid.stride(0), ptr(_f32(gate)),
                             ptr(_f32(norm_w0)), ptr(_f32(norm_w1)), norm_split, norm_cols, rope_cols, ptr(_f32(rope)),
                             stream_ptr()))
    return out


def attention(q, k, v, heads, scale=None, kv_count=None, kv_index=None, out=None, score_bound=None):
    """softmax(q k^T * scale) v per head; q [Sq, heads*64], k/v [Sk, heads*64] (views with a row pitch are fine).
    score_bound: a PROVEN bound on |q . k| * scale * log2(e) (k5_attention_bounded), None = general kernel."""
    q, k, v = _bf16(q), _bf16(k), _bf16(v)
    Sq, Sk = q.shape[0], k.shape[0]
    if out is None:
        out = torch.empty(Sq, heads * 64, device=q.device, dtype=torch.bfloat16)
    if score_bound is not None:
        check(lib().k5_attention_bounded(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(out),
                                         out.stride(0), Sq, Sk, heads, float(scale if scale is not None else 64 ** -0.5),
                                         ptr(kv_count), ptr(kv_index), float(score_bound), stream_ptr()))
        return out
    check(lib().k5_attention(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(out), out.stride(0), Sq, Sk,
                             heads, float(scale if scale is not None else 64 ** -0.5), ptr(kv_count), ptr(kv_index),
                             stream_ptr()))
    return out


def attention_split(q, k, v, heads, score_bound, split_row, scale=None, out=None, split_row2=0):
    """k5_attention_bounded_split: the bounded attention over two (three with split_row2) launches by key rows (partials in fp32)."""
    q, k, v = _bf16(q), _bf16(k), _bf16(v)
    Sq, Sk = q.shape[0], k.shape[0]
    if out is None:
        out = torch.empty(Sq, heads * 64, device=q.device, dtype=torch.bfloat16)
    ws = torch.empty(Sq * heads * 68, device=q.device, dtype=torch.float32)
    check(lib().k5_attention_bounded_split(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(out),
                                           out.stride(0), Sq, Sk, heads, float(scale if scale is not None else 64 ** -0.5),
                                           float(score_bound), int(split_row), int(split_row2), ptr(ws), stream_ptr()))
    return out


def ln_rows(x, mul, add, plus_one=True, eps=1e-5, out=None):
    """bf16(LayerNorm(x) * (mul + plus_one) + add) row-wise."""
    x = _bf16(x)
    S, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib().k5_ln_rows(ptr(x), x.stride(0), ptr(out), out.stride(0), S, D, ptr(_f32(mul)), ptr(_f32(add)),
                           1 if plus_one else 0, float(eps), stream_ptr()))
    return out


def sta_mask(T, Hb, Wb, wT, wH, wW, device="cuda"):
    n = T * Hb * Wb
    out = torch.empty(n, n, device=device, dtype=torch.uint8)
    check(lib().k5_sta_mask(T, Hb, Wb, wT, wH, wW, ptr(out), stream_ptr()))
    return out


def nabla_select(q, k, heads, P, sta=None):
    """NABLA block selection -> (kv_count [heads, nb] int32, kv_index [heads, nb, nb] int32)."""
    q, k = _bf16(q), _bf16(k)
    S = q.shape[0]
    nb = S // 64
    cnt = torch.empty(heads, nb, device=q.device, dtype=torch.int32)
    idx = torch.zeros(heads, nb, nb, device=q.device, dtype=torch.int32)      # entries past kv_count stay 0
    ws = torch.empty(heads * nb * nb + 2 * nb * heads * 64, device=q.device, dtype=torch.float32)
    check(lib().k5_nabla_select(ptr(q), q.stride(0), ptr(k), k.stride(0), S, heads, float(P), ptr(sta), ptr(cnt), ptr(idx),
                                ptr(ws), stream_ptr()))
    return cnt, idx
