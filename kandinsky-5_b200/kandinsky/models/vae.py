"""Host-side mirror of the DECODE half of the reference's `kandinsky/models/vae.py` (AutoencoderKLHunyuanVideo):
same class name, `config.scaling_factor`, `decode(z).sample`, tiling selection and state-dict keys, but the decoder
is one call into libk5 (k5_vae_decode: implicit-GEMM causal convolutions on tcgen05, fused GroupNorm / SiLU / pad /
up-sample passes, temporal tile blending).  Encoding is not on the T2V path and is not provided."""
import ctypes
import json
import math
import os
from ctypes import c_int32, c_int64, c_void_p
from types import SimpleNamespace

import torch

from .. import _lib
from .._lib import K5VaeConfig, check, lib, ptr, stream_ptr

_DTYPE_CODE = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def _temporal_tiling_table():
    """(tile, stride) in sample frames per video length: the reference's OPT_TEMPORAL_TILING (vae.py:26-85).  From 61
    frames on it repeats with a period of 48 frames; the head is irregular."""
    head = {1: (1, 1), 17: (17, 17), 21: (13, 8), 25: (17, 8), 29: (17, 12), 33: (21, 12), 37: (21, 16), 41: (17, 12),
            45: (21, 12), 49: (17, 8), 53: (21, 16), 57: (21, 12)}
    period = [(13, 8), (17, 12), (21, 16), (17, 8), (17, 12), (21, 12), (21, 16), (17, 12), (21, 12), (17, 8), (21, 16),
              (21, 12)]
    table = dict(head)
    for n in range(61, 242, 4):
        table[n] = period[((n - 61) // 4) % 12]
    return table


OPT_TEMPORAL_TILING = _temporal_tiling_table()


def decoder_state_dict_shapes(block_out_channels=(128, 256, 512, 512), latent_channels=16, out_channels=3):
    """Checkpoint key -> shape of what decode needs (vae.py:589-680, 769-771): 138 decoder.* tensors
    and post_quant_conv."""
    s = {}
    rev = list(reversed(block_out_channels))
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

    conv("post_quant_conv", latent_channels, latent_channels, 1)
    conv("decoder.conv_in.conv", top, latent_channels)
    resnet("decoder.mid_block.resnets.0.", top, top)
    resnet("decoder.mid_block.resnets.1.", top, top)
    a = "decoder.mid_block.attentions.0."
    s[a + "group_norm.weight"], s[a + "group_norm.bias"] = (top,), (top,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s[a + n + ".weight"], s[a + n + ".bias"] = (top, top), (top,)
    prev = top
    for i, co in enumerate(rev):
        for j in range(3):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}.", prev if j == 0 else co, co)
        if i < 3:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv.conv", co, co)
        prev = co
    s["decoder.conv_norm_out.weight"], s["decoder.conv_norm_out.bias"] = (block_out_channels[0],), (block_out_channels[0],)
    conv("decoder.conv_out.conv", out_channels, block_out_channels[0])
    return s


class DecoderOutput:
    """diffusers.models.autoencoders.vae.DecoderOutput: `.sample` (and tuple-style [0])."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class AutoencoderKLHunyuanVideo:
    """Decode-only drop-in for `kandinsky.models.vae.AutoencoderKLHunyuanVideo` backed by the CUDA engine.
    `max_latent` = (frames per temporal tile, height, width) bounds the engine workspace; the default covers the
    (17, 8) tiling of the 5 s / 10 s videos at 512x768 (5 latent frames of 64x96)."""

    def __init__(self, in_channels=3, out_channels=3, latent_channels=16, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, act_fn="silu", norm_num_groups=32, scaling_factor=0.476986,
                 spatial_compression_ratio=8, temporal_compression_ratio=4, mid_block_add_attention=True,
                 max_latent=(6, 64, 96), **_ignored):
        if (layers_per_block, norm_num_groups, spatial_compression_ratio, temporal_compression_ratio) != (2, 32, 8, 4):
            raise ValueError("the engine implements the HunyuanVideo decoder: layers_per_block=2, 32 groups, 8x / 4x compression")
        if act_fn not in ("silu", "swish") or not mid_block_add_attention:
            raise ValueError("the engine implements SiLU and the mid-block attention of the HunyuanVideo decoder")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      act_fn=act_fn, norm_num_groups=norm_num_groups, scaling_factor=scaling_factor,
                                      spatial_compression_ratio=8, temporal_compression_ratio=4,
                                      mid_block_add_attention=True)
        self.max_latent = tuple(int(v) for v in max_latent)
        self.spatial_compression_ratio, self.temporal_compression_ratio = 8, 4
        self.tile_size = None
        self._tile_frames, self._stride_frames = 0, 0
        self._engine, self._device, self._pending = None, None, None
        self.dtype = torch.float16

    # ---- engine lifetime -------------------------------------------------------------------------------
    def _create_engine(self, device):
        c = K5VaeConfig()
        c.block_out_channels = (c_int32 * 4)(*self.config.block_out_channels)
        c.latent_channels, c.out_channels = self.config.latent_channels, self.config.out_channels
        c.max_tile_frames, c.max_height, c.max_width = self.max_latent
        handle = c_void_p()
        with torch.cuda.device(device):
            check(lib().k5_vae_create(ctypes.byref(c), ctypes.byref(handle)))
        self._engine, self._device = handle, torch.device(device)

    def __del__(self):
        eng = getattr(self, "_engine", None)
        if eng is not None and _lib._lib is not None:
            _lib._lib.k5_vae_destroy(eng)
            self._engine = None

    def _shapes(self):
        return decoder_state_dict_shapes(self.config.block_out_channels, self.config.latent_channels, self.config.out_channels)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        """Takes the full diffusers checkpoint (encoder.* / quant_conv.* entries are not needed for decoding and are
        skipped) or just the decoder part; every decoder tensor must be present with the reference's shape."""
        want = self._shapes()
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want and not k.startswith(("encoder.", "quant_conv."))]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                               f"unexpected keys {unexpected[:5]}{'...' if len(unexpected) > 5 else ''}")
        for k, shp in want.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(state_dict[k].shape)} vs model {tuple(shp)}")
        sd = {k: v for k, v in state_dict.items() if k in want}
        if self._engine is None:
            self._pending = sd
        else:
            self._upload(sd)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _upload(self, sd):
        with torch.cuda.device(self._device):
            for k, t in sd.items():
                t = t.detach()
                if t.dtype not in _DTYPE_CODE:
                    t = t.float()
                t = t.contiguous()
                shape = (c_int64 * t.dim())(*t.shape)
                check(lib().k5_vae_load_tensor(self._engine, k.encode(), c_void_p(t.data_ptr()), _DTYPE_CODE[t.dtype], shape,
                                               t.dim()))
            check(lib().k5_vae_finalize(self._engine))

    def to(self, device=None, *args, **kwargs):
        if device is not None and not isinstance(device, torch.dtype) and torch.device(device).type == "cuda":
            dev = torch.device(device)
            if dev.index is None:
                dev = torch.device("cuda", torch.cuda.current_device())
            if self._engine is None:
                self._create_engine(dev)
                if self._pending is not None:
                    self._upload(self._pending)
                    self._pending = None
            elif dev != self._device:
                raise RuntimeError("the engine is bound to one GPU; create a new model for another device")
        return self

    def eval(self):
        return self

    # ---- tiling selection (vae.py:1230-1273) ------------------------------------------------------------
    def get_enc_optimal_tiling(self, shape):
        _, _, num_frames, height, width = shape
        if math.sqrt(height * width) < 450 and num_frames <= 97:
            ft, fs = num_frames, num_frames
        else:
            ft, fs = OPT_TEMPORAL_TILING[num_frames]
        if math.sqrt(height * width) > 900:
            raise NotImplementedError("spatial tiling (videos larger than ~900x900) is outside the T2V path of this engine")
        return (1, ft, height, width), (fs, height, width)

    def get_dec_optimal_tiling(self, shape):
        b, _, f, h, w = shape
        return self.get_enc_optimal_tiling([b, 3, 4 * (f - 1) + 1, 8 * h, 8 * w])

    def apply_tiling(self, tile, stride):
        _, ft, ht, wt = tile
        fs, hs, ws = stride
        self._tile_frames, self._stride_frames = int(ft), int(fs)
        self._tile_hw = (ht, wt, hs, ws)

    # ---- decode ----------------------------------------------------------------------------------------
    def _decode(self, z, return_dict=True):
        if self._engine is None:
            raise RuntimeError("VAE is not on a CUDA device / no weights loaded: call load_state_dict() and .to('cuda')")
        if z.dim() != 5 or z.shape[0] != 1 or z.shape[1] != self.config.latent_channels:
            raise ValueError("z must be [1, latent_channels, T, H, W]")
        _, _, T, H, W = z.shape
        ht, wt, _, _ = getattr(self, "_tile_hw", (8 * H, 8 * W, 8 * H, 8 * W))
        if 8 * H > ht or 8 * W > wt:
            raise NotImplementedError("spatial tiling is outside the T2V path of this engine")
        with torch.cuda.device(self._device):
            zz = z[0].to(self._device, torch.float32).contiguous()
            out = torch.empty(1, 3, 4 * (T - 1) + 1, 8 * H, 8 * W, device=self._device, dtype=torch.bfloat16)
            check(lib().k5_vae_decode(self._engine, ptr(zz), T, H, W, self._tile_frames, self._stride_frames, ptr(out),
                                      stream_ptr()))
        return DecoderOutput(out) if return_dict else (out,)

    @torch.no_grad()
    def decode(self, z, return_dict=True):
        """vae.py:880-906: pick the tiling for this latent shape, then decode.  Returns bf16 [1, 3, 4(T-1)+1, 8H, 8W]
        (the reference's dtype under its bf16 autocast)."""
        tile_size, tile_stride = self.get_dec_optimal_tiling(z.shape)
        if tile_size != self.tile_size:
            self.tile_size = tile_size
            self.apply_tiling(tile_size, tile_stride)
        return self._decode(z, return_dict=return_dict)


def build_vae(conf):
    """vae.py:1276-1282: diffusers layout <checkpoint_path>/vae/{config.json, diffusion_pytorch_model.safetensors}."""
    name = conf["name"] if isinstance(conf, dict) else conf.name
    path = conf["checkpoint_path"] if isinstance(conf, dict) else conf.checkpoint_path
    assert name == "hunyuan", f"unknown vae name {name}"
    root = os.path.join(path, "vae")
    cfg = {}
    cfg_path = os.path.join(root, "config.json")
    if os.path.exists(cfg_path):
        with open(cfg_path) as f:
            cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
    weights = os.path.join(root, "diffusion_pytorch_model.safetensors")
    if not os.path.exists(weights):
        raise FileNotFoundError(f"VAE checkpoint not found: {weights} (downloads are not available)")
    from safetensors.torch import load_file

    vae = AutoencoderKLHunyuanVideo(**{k: v for k, v in cfg.items() if k in (
        "in_channels", "out_channels", "latent_channels", "block_out_channels", "layers_per_block", "act_fn",
        "norm_num_groups", "scaling_factor", "spatial_compression_ratio", "temporal_compression_ratio",
        "mid_block_add_attention")})
    vae.load_state_dict(load_file(weights))
    return vae
