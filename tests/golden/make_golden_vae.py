#!/usr/bin/env python
"""Mint VAE-decoder golden vectors by executing the REFERENCE's own `kandinsky/models/vae.py` on the CPU.

    python tests/golden/make_golden_vae.py        # writes tests/golden/vae_*.pt   (build container only)

`diffusers` is not installed here, so the handful of names vae.py imports from it are stubbed:
ConfigMixin / ModelMixin / register_to_config / apply_forward_hook / output dataclasses are plumbing with no
arithmetic; `get_activation("silu"|"swish")` is nn.SiLU; `Attention` (diffusers.models.attention_processor, un-pinned
in the reference's requirements.txt) is RESTATED below from its published behaviour for the constructor arguments
vae.py:311-323 passes -- GroupNorm on [B,C,N], to_q / to_k / to_v, scaled_dot_product_attention with the additive mask,
to_out[0], + residual.  That one module is therefore "parity unpinned"; everything else in the decoder (causal
convolutions, resnets, up-sampling, temporal tiling and blending) is the reference's own code.
The CUDA autocast dtype policy is emulated exactly like in make_golden.py (conv3d / linear / sdpa -> bf16,
group_norm -> fp32), so the vectors carry the rounding points of the reference's CUDA path.
"""
import importlib
import os
import sys
import types
from dataclasses import dataclass

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (autocast emulation shared with the DiT vectors)

MG._LOWER.update({F.conv3d, torch.conv3d})
MG._FP32.update({F.group_norm, torch.group_norm})


class Attention(nn.Module):
    """Restatement of diffusers' Attention for (heads=1, dim_head=C, norm_num_groups, residual_connection=True,
    bias=True, _from_deprecated_attn_block=True) with the default SDPA processor."""

    def __init__(self, query_dim, heads=8, dim_head=64, eps=1e-5, norm_num_groups=None, residual_connection=False,
                 bias=False, upcast_softmax=False, _from_deprecated_attn_block=False, rescale_output_factor=1.0):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.residual_connection, self.rescale_output_factor = heads, residual_connection, rescale_output_factor
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True) if norm_num_groups else None
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, attention_mask=None):
        residual = hidden_states
        b, n, _ = hidden_states.shape
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(hidden_states), self.to_k(hidden_states), self.to_v(hidden_states)
        hd = q.shape[-1] // self.heads
        q, k, v = (t.view(b, n, self.heads, hd).transpose(1, 2) for t in (q, k, v))
        mask = attention_mask.view(b, 1, n, n) if attention_mask is not None else None
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, n, self.heads * hd).to(q.dtype)
        o = self.to_out[1](self.to_out[0](o))
        if self.residual_connection:
            o = o + residual
        return o / self.rescale_output_factor


def install_diffusers_stub():
    @dataclass
    class DecoderOutput:
        sample: torch.Tensor

    @dataclass
    class AutoencoderKLOutput:
        latent_dist: object = None

    class ConfigMixin:
        pass

    class ModelMixin(nn.Module):
        pass

    def register_to_config(init):
        def wrapped(self, *a, **k):
            self.config = types.SimpleNamespace(**k)
            return init(self, *a, **k)

        return wrapped

    def get_activation(name):
        assert name in ("silu", "swish"), name
        return nn.SiLU()

    tree = {
        "diffusers": {},
        "diffusers.configuration_utils": dict(ConfigMixin=ConfigMixin, register_to_config=register_to_config),
        "diffusers.utils": {},
        "diffusers.utils.accelerate_utils": dict(apply_forward_hook=lambda f: f),
        "diffusers.models": {},
        "diffusers.models.activations": dict(get_activation=get_activation),
        "diffusers.models.attention_processor": dict(Attention=Attention),
        "diffusers.models.modeling_outputs": dict(AutoencoderKLOutput=AutoencoderKLOutput),
        "diffusers.models.modeling_utils": dict(ModelMixin=ModelMixin),
        "diffusers.models.autoencoders": {},
        "diffusers.models.autoencoders.vae": dict(DecoderOutput=DecoderOutput, DiagonalGaussianDistribution=object),
    }
    for name, attrs in tree.items():
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m


def import_reference_vae():
    install_diffusers_stub()
    for name, path in (("kandinsky", os.path.join(MG.REF, "kandinsky")),
                       ("kandinsky.models", os.path.join(MG.REF, "kandinsky", "models"))):
        if name not in sys.modules or not getattr(sys.modules[name], "__path__", None):
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    env = dict(os.environ)
    mod = importlib.import_module("kandinsky.models.vae")
    os.environ.clear()
    os.environ.update(env)               # vae.py:20-21 sets allocator / inductor variables at import; not ours to keep
    return mod


def build_ref_vae(mod, block_out_channels, sd):
    vae = mod.AutoencoderKLHunyuanVideo(block_out_channels=tuple(block_out_channels))
    dec_keys = {k: v for k, v in vae.state_dict().items() if k.startswith("decoder.") or k.startswith("post_quant_conv.")}
    assert set(dec_keys) == set(sd), (sorted(set(dec_keys) ^ set(sd))[:6])
    for k, v in dec_keys.items():
        assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
    vae.load_state_dict({**vae.state_dict(), **sd})
    return vae.to(torch.float16).eval()          # build_vae: torch_dtype=float16 (vae.py:1279)


def run_case(mod, name, widths, zshape, tiling=None, seed=3):
    from oracle import vae_oracle as VO

    sd = VO.synthetic_state_dict(widths, seed=0)
    vae = build_ref_vae(mod, widths, sd)
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(*zshape, generator=g)
    with torch.no_grad(), MG.AutocastEmu(), MG.EmuAutocast("cuda", dtype=torch.bfloat16):
        if tiling is None:
            out = vae.decode(z).sample
        else:
            # the reference picks (17, 8) temporal tiling for 121 / 241-frame videos at 512x768 (vae.py:54,84,1248-1273);
            # small test volumes never trigger it by themselves, so apply the same tiling explicitly
            vae.apply_tiling(*tiling)
            out = vae._decode(z).sample
    rec = dict(name=name, widths=tuple(widths), z=z, out=out.clone(), tiling=tiling, weight_seed=0)
    torch.save(rec, os.path.join(HERE, name + ".pt"))
    print(name, tuple(out.shape), out.dtype, float(out.float().abs().mean()))


def main():
    mod = import_reference_vae()
    import json

    with open(os.path.join(HERE, "vae_temporal_tiling.json"), "w") as f:       # the reference's tiling table, as data
        json.dump({str(k): list(v) for k, v in mod.OPT_TEMPORAL_TILING.items()}, f)
    # one un-tiled causal decode at the real decoder widths: 3 latent frames 8x8 -> 9 frames 64x64
    run_case(mod, "vae_full_width_3x8x8", (128, 256, 512, 512), (1, 16, 3, 8, 8))
    # temporal tiling + blending exactly as for the 5 s video ((17, 8): 5-latent-frame tiles, stride 2, blend 8 frames)
    # on a narrow decoder: 9 latent frames -> 3 tiles -> 33 frames
    run_case(mod, "vae_tiled_9x8x8", (64, 64, 128, 128), (1, 16, 9, 8, 8), tiling=((1, 17, 64, 64), (8, 64, 64)))


if __name__ == "__main__":
    main()
