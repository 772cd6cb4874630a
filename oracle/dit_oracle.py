"""CPU oracle for the Kandinsky-5 DiT denoising path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU restatement of the reference algorithm.  It is the
checker for the CUDA engine: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``kandinsky-5_b200/``) never imports anything from ``oracle/``.

Parity status: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the reference's own Python code executed
in the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.pt``),
with CUDA-autocast rounding points emulated on the CPU.  See DESIGN.md §3.

Every function cites the reference file:line (relative to /root/reference) it
restates.  State-dict keys follow the reference checkpoint contract
(kandinsky/utils.py:115-116).

Numerics modes
  mode="cuda"  rounding points of the reference's CUDA autocast path (SURVEY.md
               Appendix A): bf16 GEMM operands / fp32 accumulate / bf16 results,
               fp32 LayerNorm / modulation / time MLP, bf16 residual stream.
  mode="gold"  everything in fp32 (used only to calibrate tolerances).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BF16 = torch.bfloat16
F32 = torch.float32

LN_EPS = 1e-5                               # nn.LayerNorm default (nn.py:68,369; dit.py:27,30,52,55,58)
RMS_EPS = torch.finfo(torch.float32).eps    # nn.RMSNorm(eps=None) on fp32 input (nn.py:228-229)


# --------------------------------------------------------------------------- helpers
def _lin(x: Tensor, w: Tensor, b: Optional[Tensor], mode: str) -> Tensor:
    """nn.Linear under ``autocast(cuda, bf16)``: bf16 operands, fp32 accumulate, bf16 out."""
    if mode == "gold":
        return F.linear(x.float(), w.float(), None if b is None else b.float())
    return F.linear(x.to(BF16), w.to(BF16), None if b is None else b.to(BF16))


def _lin32(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """nn.Linear inside an ``autocast(cuda, float32)`` region (nn.py:56,162)."""
    return F.linear(x.float(), w.float(), None if b is None else b.float())


def _r(x: Tensor, mode: str) -> Tensor:
    """bf16 rounding point (identity in gold mode)."""
    return x.float() if mode == "gold" else x.to(BF16)


def get_freqs(dim: int, max_period: float = 10000.0) -> Tensor:
    """models/utils.py:21-28."""
    return torch.exp(-math.log(max_period) * torch.arange(0, dim, dtype=F32) / dim)


def scale_shift_norm(x: Tensor, scale: Tensor, shift: Tensor, mode: str) -> Tensor:
    """apply_scale_shift_norm, nn.py:25-28: bf16(LN_noaffine_fp32(x) * (scale + 1) + shift)."""
    y = F.layer_norm(x.float(), (x.shape[-1],), None, None, LN_EPS)
    return _r(y * (scale.float() + 1.0) + shift.float(), mode)


def gate_sum(x: Tensor, out: Tensor, gate: Tensor, mode: str) -> Tensor:
    """apply_gate_sum, nn.py:30-33: bf16(x + gate * out) in fp32."""
    return _r(x.float() + gate.float() * out.float(), mode)


def apply_rotary(x: Tensor, cos: Tensor, sin: Tensor, mode: str) -> Tensor:
    """apply_rotary, nn.py:35-40 with the rotation matrix [[cos,-sin],[sin,cos]] of
    RoPE1D/RoPE3D (nn.py:110-116,147-150).  x: [S, heads, 2*P]; cos/sin: [S, P] fp32."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    x0, x1 = xf[..., 0], xf[..., 1]
    c, s = cos[:, None, :], sin[:, None, :]
    # reference: (rope * x_).sum(-1): two products then one add, all fp32
    o0 = c * x0 + (-s) * x1
    o1 = s * x0 + c * x1
    return _r(torch.stack([o0, o1], dim=-1).reshape(x.shape), mode)


def rms_norm_heads(x: Tensor, w: Tensor, mode: str) -> Tensor:
    """norm_qk, nn.py:246-250: RMSNorm over head_dim in fp32, learned weight, back to bf16."""
    return _r(F.rms_norm(x.float(), (x.shape[-1],), w.float(), RMS_EPS), mode)


def attention(q: Tensor, k: Tensor, v: Tensor, mode: str, block_mask: Optional[Tensor] = None) -> Tensor:
    """FA(q,k,v) of nn.py:201,254,336: softmax(q k^T / sqrt(d)) v, non-causal,
    fp32 softmax and accumulation, bf16 in/out.  q:[Sq,h,d] k,v:[Sk,h,d] -> [Sq,h*d].
    block_mask: optional bool [h, Sq/64, Sk/64] (NABLA; all selected 64x64 blocks are full)."""
    qh, kh, vh = (t.transpose(0, 1) for t in (q, k, v))          # [h,S,d]
    if block_mask is None and mode != "gold":
        o = F.scaled_dot_product_attention(qh[None], kh[None], vh[None])[0]
    else:
        s = torch.matmul(qh.float(), kh.float().transpose(-1, -2)) / math.sqrt(q.shape[-1])
        if block_mask is not None:
            m = block_mask.repeat_interleave(64, dim=-2).repeat_interleave(64, dim=-1)
            s = s.masked_fill(~m, float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = torch.matmul(p, vh.float())
    return _r(o.transpose(0, 1).flatten(-2, -1), mode)


def feed_forward(sd: Dict[str, Tensor], pfx: str, x: Tensor, mode: str) -> Tensor:
    """FeedForward, nn.py:352-361: W2 . GELU_erf(W1 x), no biases."""
    h = _lin(x, sd[pfx + "in_layer.weight"], None, mode)
    h = F.gelu(h)                                   # exact erf, bf16 storage / fp32 opmath
    return _lin(h, sd[pfx + "out_layer.weight"], None, mode)


def modulation(sd: Dict[str, Tensor], pfx: str, time_embed: Tensor) -> Tensor:
    """Modulation, nn.py:153-164: Linear(SiLU(time_embed)) in fp32."""
    return _lin32(F.silu(time_embed.float()), sd[pfx + "out_layer.weight"], sd[pfx + "out_layer.bias"])


# --------------------------------------------------------------------------- embeddings
def time_embeddings(sd, time: Tensor, model_dim: int) -> Tensor:
    """TimeEmbeddings, nn.py:43-61 (fp32)."""
    freqs = get_freqs(model_dim // 2).to(time.device)
    args = torch.outer(time.float(), freqs)
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    h = F.silu(_lin32(emb, sd["time_embeddings.in_layer.weight"], sd["time_embeddings.in_layer.bias"]))
    return _lin32(h, sd["time_embeddings.out_layer.weight"], sd["time_embeddings.out_layer.bias"])


def text_embeddings(sd, pfx: str, x: Tensor, mode: str) -> Tensor:
    """TextEmbeddings, nn.py:64-72: Linear(bf16) -> LayerNorm(affine, fp32) -> cast back."""
    h = _lin(x, sd[pfx + "in_layer.weight"], sd[pfx + "in_layer.bias"], mode)
    y = F.layer_norm(h.float(), (h.shape[-1],), sd[pfx + "norm.weight"].float(), sd[pfx + "norm.bias"].float(), LN_EPS)
    return _r(y, mode)


def visual_embeddings(sd, x: Tensor, patch: Sequence[int], mode: str) -> Tensor:
    """VisualEmbeddings, nn.py:75-96: patchify with inner order (pt,ph,pw,c), then Linear."""
    T, H, W, C = x.shape
    p0, p1, p2 = patch
    x = x.view(T // p0, p0, H // p1, p1, W // p2, p2, C).permute(0, 2, 4, 1, 3, 5, 6).flatten(3, 6)
    return _lin(x, sd["visual_embeddings.in_layer.weight"], sd["visual_embeddings.in_layer.bias"], mode)


def rope_1d(pos: Tensor, dim: int):
    """RoPE1D, nn.py:99-116 -> (cos, sin) [L, dim/2]."""
    args = torch.outer(torch.arange(1024, dtype=F32), get_freqs(dim // 2)).to(pos.device)[pos]
    return torch.cos(args), torch.sin(args)


def rope_3d(shape, pos, axes_dims, scale_factor):
    """RoPE3D, nn.py:119-150 -> (cos, sin) [T,H,W, sum(axes)/2]."""
    T, H, W = shape
    a = []
    for i, ad in enumerate(axes_dims):
        tab = torch.outer(torch.arange(128, dtype=F32), get_freqs(ad // 2)).to(pos[0].device)
        a.append(tab[pos[i]] / scale_factor[i])
    args = torch.cat(
        [
            a[0].view(T, 1, 1, -1).expand(T, H, W, -1),
            a[1].view(1, H, 1, -1).expand(T, H, W, -1),
            a[2].view(1, 1, W, -1).expand(T, H, W, -1),
        ],
        dim=-1,
    )
    return torch.cos(args), torch.sin(args)


# --------------------------------------------------------------------------- token order
def fractal_flatten(x: Tensor, shape, fractal: bool) -> Tensor:
    """fractal_flatten / local_patching, models/utils.py:31-41,54-78: plain (t,h,w) order, or
    (t, hb, wb, hi, wi) with 8x8 spatial tiles so that each 64-token block is one tile."""
    T, H, W = shape
    if not fractal:
        return x.reshape(T * H * W, *x.shape[3:])
    x = x.reshape(T, H // 8, 8, W // 8, 8, *x.shape[3:])
    x = x.permute(0, 1, 3, 2, 4, *range(5, x.dim()))
    return x.reshape(T * H * W, *x.shape[5:])


def fractal_unflatten(x: Tensor, shape, fractal: bool) -> Tensor:
    """fractal_unflatten / local_merge, models/utils.py:44-51,81-105."""
    T, H, W = shape
    if not fractal:
        return x.reshape(T, H, W, *x.shape[1:])
    x = x.reshape(T, H // 8, W // 8, 8, 8, *x.shape[1:])
    x = x.permute(0, 1, 3, 2, 4, *range(5, x.dim()))
    return x.reshape(T, H, W, *x.shape[5:])


# --------------------------------------------------------------------------- NABLA
def sta_mask(T: int, H: int, W: int, wT: int, wH: int, wW: int) -> Tensor:
    """fast_sta_nabla, models/utils.py:108-133: bool [T*H*W, T*H*W] over the block grid,
    row-major (t,h,w): keep iff |dt|<=wT//2 and |dh|<=wH//2 and |dw|<=wW//2."""
    t = torch.arange(T).view(T, 1, 1).expand(T, H, W).reshape(-1)
    h = torch.arange(H).view(1, H, 1).expand(T, H, W).reshape(-1)
    w = torch.arange(W).view(1, 1, W).expand(T, H, W).reshape(-1)
    return (
        ((t[:, None] - t[None, :]).abs() <= wT // 2)
        & ((h[:, None] - h[None, :]).abs() <= wH // 2)
        & ((w[:, None] - w[None, :]).abs() <= wW // 2)
    )


def nabla_block_mask(q: Tensor, k: Tensor, sta: Optional[Tensor], thr: float, mode: str) -> Tensor:
    """nablaT_v2, models/utils.py:136-163.  q,k: [S,h,d] (post-norm, post-RoPE, fractal order).
    Returns bool [h, S/64, S/64]: block kept iff the ascending cumulative softmax mass of the
    block-mean-pooled score row reaches >= 1-thr at it, OR it is in the STA mask."""
    S, h, d = q.shape
    nb = S // 64
    qh, kh = q.transpose(0, 1), k.transpose(0, 1)                  # [h,S,d]
    if mode == "gold":
        qa = qh.float().reshape(h, nb, 64, d).mean(-2)
        ka = kh.float().reshape(h, nb, 64, d).mean(-2)
        mp = torch.matmul(qa, ka.transpose(-1, -2)) / math.sqrt(d)
    else:
        qa = qh.reshape(h, nb, 64, d).mean(-2)                    # bf16 mean (no autocast rule)
        ka = kh.reshape(h, nb, 64, d).mean(-2)
        mp = torch.matmul(qa, ka.transpose(-1, -2)) / math.sqrt(d)  # bf16 matmul, bf16 divide
    mp = torch.softmax(mp.float(), dim=-1)
    vals, inds = mp.sort(-1)
    cvals = vals.cumsum(-1)
    mask = (cvals >= 1 - thr).int()
    mask = mask.gather(-1, inds.argsort(-1)).bool()
    if sta is not None:
        mask = mask | sta.reshape(1, nb, nb)
    return mask


# --------------------------------------------------------------------------- blocks
def _self_attention(sd, pfx, x, cos, sin, mode, sparse=None):
    """MultiheadSelfAttentionEnc/Dec.forward, nn.py:208-217, 286-298."""
    hd = sd[pfx + "query_norm.weight"].shape[0]
    q = _lin(x, sd[pfx + "to_query.weight"], sd[pfx + "to_query.bias"], mode)
    k = _lin(x, sd[pfx + "to_key.weight"], sd[pfx + "to_key.bias"], mode)
    v = _lin(x, sd[pfx + "to_value.weight"], sd[pfx + "to_value.bias"], mode)
    S = x.shape[0]
    q, k, v = (t.reshape(S, -1, hd) for t in (q, k, v))
    q = rms_norm_heads(q, sd[pfx + "query_norm.weight"], mode)
    k = rms_norm_heads(k, sd[pfx + "key_norm.weight"], mode)
    q = apply_rotary(q, cos, sin, mode)
    k = apply_rotary(k, cos, sin, mode)
    bm = None
    if sparse is not None:
        bm = nabla_block_mask(q, k, sparse.get("sta_mask"), sparse["P"], mode)
        if sparse.get("_record") is not None:
            sparse["_record"].append(bm)
    o = attention(q, k, v, mode, bm)
    return _lin(o, sd[pfx + "out_layer.weight"], sd[pfx + "out_layer.bias"], mode)


def _cross_attention(sd, pfx, x, cond, mode):
    """MultiheadCrossAttention.forward, nn.py:343-349 (no RoPE)."""
    hd = sd[pfx + "query_norm.weight"].shape[0]
    q = _lin(x, sd[pfx + "to_query.weight"], sd[pfx + "to_query.bias"], mode)
    k = _lin(cond, sd[pfx + "to_key.weight"], sd[pfx + "to_key.bias"], mode)
    v = _lin(cond, sd[pfx + "to_value.weight"], sd[pfx + "to_value.bias"], mode)
    q = q.reshape(x.shape[0], -1, hd)
    k = k.reshape(cond.shape[0], -1, hd)
    v = v.reshape(cond.shape[0], -1, hd)
    q = rms_norm_heads(q, sd[pfx + "query_norm.weight"], mode)
    k = rms_norm_heads(k, sd[pfx + "key_norm.weight"], mode)
    o = attention(q, k, v, mode)
    return _lin(o, sd[pfx + "out_layer.weight"], sd[pfx + "out_layer.bias"], mode)


def encoder_block(sd, pfx, x, time_embed, cos, sin, mode):
    """TransformerEncoderBlock.forward, dit.py:33-44."""
    m = modulation(sd, pfx + "text_modulation.", time_embed)
    sa, ff = torch.chunk(m, 2, dim=-1)
    shift, scale, gate = torch.chunk(sa, 3, dim=-1)
    out = scale_shift_norm(x, scale, shift, mode)
    out = _self_attention(sd, pfx + "self_attention.", out, cos, sin, mode)
    x = gate_sum(x, out, gate, mode)
    shift, scale, gate = torch.chunk(ff, 3, dim=-1)
    out = scale_shift_norm(x, scale, shift, mode)
    out = feed_forward(sd, pfx + "feed_forward.", out, mode)
    return gate_sum(x, out, gate, mode)


def decoder_block(sd, pfx, x, text_embed, time_embed, cos, sin, mode, sparse=None):
    """TransformerDecoderBlock.forward, dit.py:61-79."""
    m = modulation(sd, pfx + "visual_modulation.", time_embed)
    sa, ca, ff = torch.chunk(m, 3, dim=-1)
    shift, scale, gate = torch.chunk(sa, 3, dim=-1)
    out = scale_shift_norm(x, scale, shift, mode)
    out = _self_attention(sd, pfx + "self_attention.", out, cos, sin, mode, sparse)
    x = gate_sum(x, out, gate, mode)
    shift, scale, gate = torch.chunk(ca, 3, dim=-1)
    out = scale_shift_norm(x, scale, shift, mode)
    out = _cross_attention(sd, pfx + "cross_attention.", out, text_embed, mode)
    x = gate_sum(x, out, gate, mode)
    shift, scale, gate = torch.chunk(ff, 3, dim=-1)
    out = scale_shift_norm(x, scale, shift, mode)
    out = feed_forward(sd, pfx + "feed_forward.", out, mode)
    return gate_sum(x, out, gate, mode)


def out_layer(sd, x, time_embed, patch, mode):
    """OutLayer.forward, nn.py:374-400: modulation(2) order (shift, scale); un-patchify with
    inner order (c, pt, ph, pw)."""
    shift, scale = torch.chunk(modulation(sd, "out_layer.modulation.", time_embed), 2, dim=-1)
    y = scale_shift_norm(x, scale[:, None, None], shift[:, None, None], mode)
    y = _lin(y, sd["out_layer.out_layer.weight"], sd["out_layer.out_layer.bias"], mode)
    T, H, W, _ = y.shape
    p0, p1, p2 = patch
    y = y.view(T, H, W, -1, p0, p1, p2).permute(0, 4, 1, 5, 2, 6, 3)
    return y.flatten(0, 1).flatten(1, 2).flatten(2, 3)


# --------------------------------------------------------------------------- full forward
def dit_forward(
    sd: Dict[str, Tensor],
    cfg: dict,
    x: Tensor,
    text_embed: Tensor,
    pooled_text_embed: Tensor,
    time: Tensor,
    visual_rope_pos,
    text_rope_pos: Tensor,
    scale_factor=(1.0, 1.0, 1.0),
    sparse_params: Optional[dict] = None,
    mode: str = "cuda",
    taps: Optional[dict] = None,
) -> Tensor:
    """DiffusionTransformer3D.forward, dit.py:155-181.  x: [T,H,W,Cin] fp32 ->
    velocity [T*pt, H, W, Cout] (bf16 in cuda mode)."""
    patch = tuple(cfg["patch_size"])
    axes = tuple(cfg["axes_dims"])
    D = cfg["model_dim"]
    hd = sum(axes)
    # before_text_transformer_blocks, dit.py:130-137
    te = text_embeddings(sd, "text_embeddings.", text_embed, mode)
    tm = time_embeddings(sd, time, D)
    tm = tm + text_embeddings(sd, "pooled_text_embeddings.", pooled_text_embed, mode).float()
    ve = visual_embeddings(sd, x, patch, mode)
    tcos, tsin = rope_1d(text_rope_pos, hd)
    if taps is not None:
        taps["time_embed"] = tm.clone()
        taps["visual_embed"] = ve.clone()
    for i in range(cfg["num_text_blocks"]):
        te = encoder_block(sd, f"text_transformer_blocks.{i}.", te, tm, tcos, tsin, mode)
    if taps is not None:
        taps["text_embed"] = te.clone()
    # before_visual_transformer_blocks, dit.py:140-147
    shape = tuple(ve.shape[:-1])
    vcos, vsin = rope_3d(shape, visual_rope_pos, axes, scale_factor)
    fractal = bool(sparse_params["to_fractal"]) if sparse_params is not None else False
    ve = fractal_flatten(ve, shape, fractal)
    vcos = fractal_flatten(vcos, shape, fractal)
    vsin = fractal_flatten(vsin, shape, fractal)
    for i in range(cfg["num_visual_blocks"]):
        ve = decoder_block(sd, f"visual_transformer_blocks.{i}.", ve, te, tm, vcos, vsin, mode, sparse_params)
        if taps is not None:
            taps[f"visual_block_{i}"] = ve.clone()
    # after_blocks, dit.py:150-153
    ve = fractal_unflatten(ve, shape, fractal)
    return out_layer(sd, ve, tm, patch, mode)


# --------------------------------------------------------------------------- sampler
def schedule(num_steps: int, scheduler_scale: float) -> Tensor:
    """generation_utils.py:102-103."""
    t = torch.linspace(1, 0, num_steps + 1)
    return scheduler_scale * t / (1 + (scheduler_scale - 1) * t)


def model_input(img: Tensor, visual_cond: bool) -> Tensor:
    """generation_utils.py:107-114: append zero cond + zero mask channels."""
    if not visual_cond:
        return img
    return torch.cat([img, torch.zeros_like(img), torch.zeros([*img.shape[:-1], 1], dtype=img.dtype, device=img.device)], dim=-1)


def get_velocity(sd, cfg, x, t, text, null_text, visual_rope_pos, guidance_weight, scale_factor, sparse_params=None, mode="cuda"):
    """get_velocity, generation_utils.py:39-77; CFG combine runs in bf16."""
    v = dit_forward(sd, cfg, x, text["text_embeds"], text["pooled_embed"], t * 1000, visual_rope_pos,
                    torch.arange(text["text_embeds"].shape[0]), scale_factor, sparse_params, mode)
    if abs(guidance_weight - 1.0) > 1e-6:
        vu = dit_forward(sd, cfg, x, null_text["text_embeds"], null_text["pooled_embed"], t * 1000, visual_rope_pos,
                         torch.arange(null_text["text_embeds"].shape[0]), scale_factor, sparse_params, mode)
        v = vu + guidance_weight * (v - vu)
    return v


def generate(sd, cfg, img, num_steps, text, null_text, visual_rope_pos, guidance_weight, scheduler_scale,
             scale_factor, sparse_params=None, mode="cuda"):
    """generate, generation_utils.py:80-129, starting from the given noise ``img`` (the reference
    draws it with torch.Generator('cuda'), which the caller reproduces).  img fp32 [T,H,W,C]."""
    ts = schedule(num_steps, scheduler_scale)
    img = img.clone().float()
    for t, dt in zip(ts[:-1], torch.diff(ts)):
        x = model_input(img, cfg.get("visual_cond", False))
        v = get_velocity(sd, cfg, x, t.unsqueeze(0), text, null_text, visual_rope_pos, guidance_weight,
                         scale_factor, sparse_params, mode)
        img = img + dt * v          # 0-dim fp32 * bf16 -> bf16 product, fp32 add
    return img


# --------------------------------------------------------------------------- synthetic weights
def dit_state_dict_shapes(cfg: dict) -> Dict[str, tuple]:
    """The checkpoint key/shape contract (dit.py:82-127)."""
    D, Fd, Td = cfg["model_dim"], cfg["ff_dim"], cfg["time_dim"]
    hd = sum(cfg["axes_dims"])
    cin = (2 * cfg["in_visual_dim"] + 1) if cfg.get("visual_cond", False) else cfg["in_visual_dim"]
    pp = math.prod(cfg["patch_size"])
    s: Dict[str, tuple] = {}

    def lin(name, o, i, bias=True):
        s[name + ".weight"] = (o, i)
        if bias:
            s[name + ".bias"] = (o,)

    def attn(p):
        for n in ("to_query", "to_key", "to_value", "out_layer"):
            lin(p + n, D, D)
        s[p + "query_norm.weight"] = (hd,)
        s[p + "key_norm.weight"] = (hd,)

    lin("time_embeddings.in_layer", Td, D)
    lin("time_embeddings.out_layer", Td, Td)
    lin("text_embeddings.in_layer", D, cfg["in_text_dim"])
    s["text_embeddings.norm.weight"] = (D,)
    s["text_embeddings.norm.bias"] = (D,)
    lin("pooled_text_embeddings.in_layer", Td, cfg["in_text_dim2"])
    s["pooled_text_embeddings.norm.weight"] = (Td,)
    s["pooled_text_embeddings.norm.bias"] = (Td,)
    lin("visual_embeddings.in_layer", D, pp * cin)
    for i in range(cfg["num_text_blocks"]):
        p = f"text_transformer_blocks.{i}."
        lin(p + "text_modulation.out_layer", 6 * D, Td)
        attn(p + "self_attention.")
        lin(p + "feed_forward.in_layer", Fd, D, bias=False)
        lin(p + "feed_forward.out_layer", D, Fd, bias=False)
    for i in range(cfg["num_visual_blocks"]):
        p = f"visual_transformer_blocks.{i}."
        lin(p + "visual_modulation.out_layer", 9 * D, Td)
        attn(p + "self_attention.")
        attn(p + "cross_attention.")
        lin(p + "feed_forward.in_layer", Fd, D, bias=False)
        lin(p + "feed_forward.out_layer", D, Fd, bias=False)
    lin("out_layer.modulation.out_layer", 2 * D, Td)
    lin("out_layer.out_layer", pp * cfg["out_visual_dim"], D)
    return s


def is_fp32_key(key: str) -> bool:
    """Tensors that live in fp32 regions of the reference (time MLP, modulation, norms)."""
    return ("modulation" in key) or key.startswith("time_embeddings.") or ("norm" in key)


def synthetic_state_dict(cfg: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic synthetic checkpoint: Linear weights/biases
    ~ U(-1/sqrt(in), 1/sqrt(in)) like nn.Linear's default init; RMSNorm / LayerNorm weights
    perturbed around 1; every ``*modulation.out_layer`` re-randomised with N(0, 0.02) because
    the reference zero-inits them (nn.py:158-159), which would make every block the identity.
    GEMM operands are stored bf16, fp32-region tensors fp32.  Generated per key from its own
    generator so that any subset is reproducible."""
    out: Dict[str, Tensor] = {}
    shapes = dit_state_dict_shapes(cfg)
    for idx, (key, shape) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        if "modulation" in key:
            t = torch.randn(shape, generator=g) * 0.02
        elif key.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("norm.bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] if key.endswith(".weight") else shapes[key[: -len("bias")] + "weight"][1]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[key] = t.float() if is_fp32_key(key) else t.to(BF16)
    return out


LITE_CFG = dict(  # configs/config_5s_sft.yaml:11-29
    in_visual_dim=16, out_visual_dim=16, time_dim=512, patch_size=(1, 2, 2), model_dim=1792, ff_dim=7168,
    num_text_blocks=2, num_visual_blocks=32, axes_dims=(16, 24, 24), visual_cond=True, in_text_dim=3584,
    in_text_dim2=768,
)
