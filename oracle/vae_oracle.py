"""CPU oracle for the HunyuanVideo causal-conv3d VAE DECODER of Kandinsky-5 (config 5; SURVEY.md §8 a-V).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  Nothing under
kandinsky-5_b200/ may import this module.

A functional, plain-torch restatement of the reference decoder with the rounding points of its CUDA autocast path
(weights stored fp16, conv3d / linear / attention operands cast to bf16 with fp32 accumulation
and bf16 results, GroupNorm + SiLU + padding in fp32, residual adds in bf16).  Every function cites the reference
lines it follows (paths relative to /root/reference).

Pinning: tests/golden/vae_*.pt are produced by tests/golden/make_golden_vae.py, which executes the reference's own
kandinsky/models/vae.py (decode / _decode / _temporal_tiled_decode / blend_t and the whole HunyuanVideoDecoder3D) under
an emulation of the CUDA autocast policy; tests/test_vae_oracle.py checks this file against them.  ONE boundary is
"parity unpinned": the mid-block attention arithmetic lives in third-party diffusers.models.attention_processor
.Attention (absent from /root/reference and not installable here; requirements.txt:14 un-pinned), so the golden
generator and this oracle both restate its published behaviour for the arguments of vae.py:311-323, 355.
"""
import math

import torch
import torch.nn.functional as F

SCALING_FACTOR = 0.476986           # vae.py:732
GN_GROUPS, GN_EPS = 32, 1e-6        # vae.py:237,245,564,677
UP_FACTORS = ((1, 2, 2), (2, 2, 2), (2, 2, 2), None)   # vae.py:644-659 for time_compression 4, spatial 8


ROUNDING = True                     # False: the same graph in fp32 throughout ("gold", used to calibrate tolerances)


def _bf(x):
    return x.to(torch.bfloat16) if ROUNDING else x.float()


# ----------------------------------------------------------------------------- synthetic checkpoint
def decoder_shapes(widths=(128, 256, 512, 512), latent=16, out_ch=3):
    """Key -> shape of the decoder part of the diffusers checkpoint (vae.py:589-680, 769-771)."""
    s = {}
    rev = list(reversed(widths))
    top = rev[0]

    def conv(name, co, ci, k=3):
        s[name + ".weight"] = (co, ci, k, k, k)
        s[name + ".bias"] = (co,)

    def resnet(p, ci, co):
        s[p + "norm1.weight"], s[p + "norm1.bias"] = (ci,), (ci,)
        conv(p + "conv1.conv", co, ci)
        s[p + "norm2.weight"], s[p + "norm2.bias"] = (co,), (co,)
        conv(p + "conv2.conv", co, co)
        if ci != co:
            conv(p + "conv_shortcut.conv", co, ci, 1)

    conv("post_quant_conv", latent, latent, 1)
    conv("decoder.conv_in.conv", top, latent)
    resnet("decoder.mid_block.resnets.0.", top, top)
    resnet("decoder.mid_block.resnets.1.", top, top)
    a = "decoder.mid_block.attentions.0."
    s[a + "group_norm.weight"], s[a + "group_norm.bias"] = (top,), (top,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[a + n + ".weight"], s[a + n + ".bias"] = (top, top), (top,)
    prev = top
    for i, co in enumerate(rev):
        for j in range(3):                                   # layers_per_block + 1
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", prev if j == 0 else co, co)
        if UP_FACTORS[i] is not None:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv.conv", co, co)
        prev = co
    s["decoder.conv_norm_out.weight"], s["decoder.conv_norm_out.bias"] = (widths[0],), (widths[0],)
    conv("decoder.conv_out.conv", out_ch, widths[0])
    return s


def synthetic_state_dict(widths=(128, 256, 512, 512), seed=0):
    """Random-init decoder weights in fp16 (the reference loads the VAE with torch_dtype=float16, vae.py:1279):
    conv / linear ~ U(-1, 1) / sqrt(fan_in); norm weight 1 + 0.1 N(0,1), norm bias 0.05 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    shapes = decoder_shapes(widths)
    for k, shp in shapes.items():
        if "norm" in k:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g) if k.endswith("weight") else 0.05 * torch.randn(shp, generator=g)
        else:
            wshape = shp if k.endswith("weight") else shapes[k[:-4] + "weight"]
            fan_in = math.prod(wshape[1:])
            t = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
        sd[k] = t.to(torch.float16)
    return sd


# ----------------------------------------------------------------------------- layers
def causal_conv3d(x, w, b):
    """HunyuanVideoCausalConv3d.forward (vae.py:159-163): replicate pad (W 1,1; H 1,1; T k-1 in front, 0 behind), then
    Conv3d stride 1.  Autocast: operands bf16, fp32 accumulate, bf16 result.  x [B,C,T,H,W] any float dtype."""
    k = w.shape[-1]
    if k > 1:
        x = F.pad(x.float(), (k // 2, k // 2, k // 2, k // 2, k - 1, 0), mode="replicate")
    y = F.conv3d(_bf(x).float(), _bf(w).float(), None)
    return _bf(y + _bf(b).float().view(1, -1, 1, 1, 1))


def group_norm_silu(x, w, b, silu=True):
    """nn.GroupNorm(32, C, eps=1e-6) under autocast (fp32 in / out) followed by SiLU in fp32 (vae.py:258-264)."""
    y = F.group_norm(x.float(), GN_GROUPS, w.float(), b.float(), GN_EPS)
    return F.silu(y) if silu else y


def resnet_block(sd, p, x):
    """HunyuanVideoResnetBlockCausal3D.forward (vae.py:254-275); x bf16 [B,C,T,H,W]."""
    h = group_norm_silu(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    h = causal_conv3d(h, sd[p + "conv1.conv.weight"], sd[p + "conv1.conv.bias"])
    h = group_norm_silu(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    h = causal_conv3d(h, sd[p + "conv2.conv.weight"], sd[p + "conv2.conv.bias"])
    res = x
    if p + "conv_shortcut.conv.weight" in sd:
        res = causal_conv3d(x, sd[p + "conv_shortcut.conv.weight"], sd[p + "conv_shortcut.conv.bias"])
    return _bf(h.float() + res.float())


def frame_causal_mask(f, s, device=None):
    """prepare_causal_attention_mask (vae.py:110-122): token of frame i sees the tokens of frames <= i."""
    m = torch.ones(f, f, device=device).tril_().log_()
    return m.repeat_interleave(s, 0).repeat_interleave(s, 1)


def mid_attention(sd, p, x):
    """vae.py:343-359 + diffusers Attention (restated; see the module docstring): x bf16 [B,C,T,H,W]."""
    B, C, T, H, W = x.shape
    hs = x.permute(0, 2, 3, 4, 1).flatten(1, 3)                            # [B, N, C] bf16
    res = hs
    n = group_norm_silu(hs.transpose(1, 2), sd[p + "group_norm.weight"], sd[p + "group_norm.bias"], silu=False).transpose(1, 2)

    def lin(t, name):
        return _bf(_bf(t).float() @ _bf(sd[p + name + ".weight"]).float().t() + _bf(sd[p + name + ".bias"]).float())

    q, k, v = lin(n, "to_q"), lin(n, "to_k"), lin(n, "to_v")
    sc = (q.float() @ k.float().transpose(-1, -2)) * (C ** -0.5) + frame_causal_mask(T, H * W, x.device)
    pr = torch.softmax(sc, dim=-1)
    o = _bf(_bf(pr).float() @ v.float())
    o = lin(o, "to_out.0")
    o = _bf(o.float() + res.float())
    return o.unflatten(1, (T, H, W)).permute(0, 4, 1, 2, 3)


def upsample(sd, p, x, factor):
    """HunyuanVideoUpsampleCausal3D.forward (vae.py:187-205): the first frame is up-sampled in space only, the others
    by `factor` in (T, H, W), nearest neighbour; then the causal conv."""
    first, rest = x[:, :, :1], x[:, :, 1:]
    first = F.interpolate(first.squeeze(2).float(), scale_factor=tuple(float(f) for f in factor[1:]), mode="nearest").unsqueeze(2)
    if rest.shape[2] > 0:
        rest = F.interpolate(rest.float(), scale_factor=tuple(float(f) for f in factor), mode="nearest")
        x = torch.cat((first, rest), dim=2)
    else:
        x = first
    return causal_conv3d(_bf(x), sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"])


def decoder(sd, z):
    """post_quant_conv (vae.py:874) + HunyuanVideoDecoder3D.forward (vae.py:682-696).  z fp32 [B,16,T,H,W] -> bf16
    [B,3,4(T-1)+1,8H,8W]."""
    x = causal_conv3d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    x = causal_conv3d(x, sd["decoder.conv_in.conv.weight"], sd["decoder.conv_in.conv.bias"])
    x = resnet_block(sd, "decoder.mid_block.resnets.0.", x)
    x = mid_attention(sd, "decoder.mid_block.attentions.0.", x)
    x = resnet_block(sd, "decoder.mid_block.resnets.1.", x)
    for i in range(4):
        for j in range(3):
            x = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}.", x)
        if UP_FACTORS[i] is not None:
            x = upsample(sd, f"decoder.up_blocks.{i}.upsamplers.0.", x, UP_FACTORS[i])
    x = group_norm_silu(x, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"])
    return causal_conv3d(x, sd["decoder.conv_out.conv.weight"], sd["decoder.conv_out.conv.bias"])


# ----------------------------------------------------------------------------- temporal tiling
def temporal_tiling(num_sample_frames, height, width):
    """get_dec_optimal_tiling / get_enc_optimal_tiling (vae.py:1246-1273) restricted to what the T2V pipeline can ask
    for (sqrt(H W) <= 900: no spatial tiles).  Returns (sample_tile_frames, sample_stride_frames) or None when the
    video is decoded in one piece.  Table rows used by the 5 s / 10 s configs: 121 -> (17, 8), 241 -> (17, 8)."""
    table = {121: (17, 8), 241: (17, 8), 61: (13, 8), 33: (21, 12), 49: (17, 8), 97: (17, 8), 25: (17, 8), 21: (13, 8)}
    if math.sqrt(height * width) < 450 and num_sample_frames <= 97:
        return None
    ft, fs = table[num_sample_frames]
    return (ft, fs) if ft < num_sample_frames else None


def blend_t(a, b, extent):
    """vae.py:928-936, bf16 arithmetic with Python-float weights; modifies and returns b."""
    extent = min(a.shape[2], b.shape[2], extent)
    for x in range(extent):
        b[:, :, x] = a[:, :, -extent + x] * (1 - x / extent) + b[:, :, x] * (x / extent)
    return b


def decode(sd, z, tile=None):
    """AutoencoderKLHunyuanVideo.decode -> _decode -> _temporal_tiled_decode (vae.py:880-906, 847-877, 1144-1204).
    `tile` = (sample_tile_frames, sample_stride_frames) as applied by apply_tiling (vae.py:1230-1243), or None."""
    T = z.shape[2]
    if tile is None:
        return decoder(sd, z)
    ft, fs = tile
    min_frames = ft - 1                          # tile_sample_min_num_frames
    lat_min, lat_stride = min_frames // 4, fs // 4
    if T <= lat_min + 1:
        return decoder(sd, z)
    blend = min_frames - fs
    row = []
    for i in range(0, T - lat_min + 1, lat_stride):
        d = decoder(sd, z[:, :, i:i + lat_min + 1]).clone()
        row.append(d[:, :, 1:] if i > 0 else d)
    out = []
    for i, t in enumerate(row):
        if i > 0:
            t = blend_t(row[i - 1], t, blend)
            out.append(t[:, :, :(min_frames if i == len(row) - 1 else fs)])
        else:
            out.append(t[:, :, :fs + 1])
    return torch.cat(out, dim=2)[:, :, :(T - 1) * 4 + 1]


def to_uint8(images):
    """generation_utils.py:222: clamp, (x + 1) * 127.5 in bf16, truncating cast."""
    return ((images.clamp(-1.0, 1.0) + 1.0) * 127.5).to(torch.uint8)
