#!/usr/bin/env python
"""Mint golden vectors by executing the REFERENCE's own code (/root/reference) on the CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.pt

The reference ships no tests or fixtures, so these vectors are the pin for
``oracle/dit_oracle.py`` and, through it, for the CUDA engine.

How the reference is made to run here:
  * ``kandinsky/__init__.py`` pulls omegaconf / diffusers (absent) -> register empty package
    objects and import ``kandinsky.models.{utils,nn,dit}`` + ``kandinsky.generation_utils`` directly;
  * ``nn.py:9`` calls ``torch.cuda.get_device_capability()`` at import -> stubbed;
  * ``nn.FA`` (flash_attn) -> SDPA wrapper with the same [B,S,H,D] contract;
  * ``flex_attention`` -> block-sparse math honouring the BlockMask's KV block lists (the eager CPU
    fallback silently ignores them; see ``flex_blocksparse`` below);
  * ``TORCHDYNAMO_DISABLE=1`` -> the ``@torch.compile`` decorators are no-ops (eager semantics);
  * the reference's ``torch.autocast(device_type="cuda", ...)`` regions do nothing on a CPU-only
    box, which would silently change its dtype policy.  ``EmuAutocast`` + ``AutocastEmu`` below
    re-create the CUDA autocast cast policy (lower-precision list: linear/matmul/bmm/sdpa ->
    autocast dtype; fp32 list: layer_norm/softmax -> float32; nesting and ``enabled=False``
    honoured) with a TorchFunctionMode, so the reference executes its *CUDA* rounding points
     using CPU kernels.
No reference source is copied: it is imported from where it lies.
"""
import importlib
import importlib.util
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import torch
import torch.nn.functional as F
from torch.overrides import TorchFunctionMode

REF = os.environ.get("K5_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# ----------------------------------------------------------------------------- autocast emulation
_STATE = [(False, None)]  # stack of (enabled, dtype)


class EmuAutocast:
    """Stand-in for ``torch.autocast`` (context manager + decorator) that records the region."""

    def __init__(self, device_type, dtype=None, enabled=True, cache_enabled=None):
        self.entry = (bool(enabled), dtype if dtype is not None else torch.float16) if device_type == "cuda" else None

    def __enter__(self):
        _STATE.append(self.entry if self.entry is not None else _STATE[-1])
        return self

    def __exit__(self, *a):
        _STATE.pop()
        return False

    def __call__(self, fn):
        import functools

        @functools.wraps(fn)
        def wrapped(*a, **k):
            with self:
                return fn(*a, **k)

        return wrapped


def _cast(obj, dtype):
    if isinstance(obj, torch.Tensor) and obj.is_floating_point() and obj.dtype != dtype:
        return obj.to(dtype)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cast(o, dtype) for o in obj)
    return obj


_LOWER = {F.linear, torch.matmul, torch.Tensor.matmul, torch.Tensor.__matmul__, torch.bmm, torch.mm,
          F.scaled_dot_product_attention}
_FP32 = {F.layer_norm, torch.layer_norm, torch.softmax, torch.Tensor.softmax, F.softmax,
         torch.cumsum, torch.Tensor.cumsum}
for _n in ("linear",):
    _LOWER.add(getattr(torch._C._nn, _n))


class AutocastEmu(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        enabled, dtype = _STATE[-1]
        if enabled:
            if func in _LOWER:
                args = _cast(args, dtype)
                kwargs = {k: _cast(v, dtype) for k, v in kwargs.items()}
            elif func in _FP32:
                args = _cast(args, torch.float32)
                kwargs = {k: _cast(v, torch.float32) for k, v in kwargs.items()}
        return func(*args, **kwargs)


# ----------------------------------------------------------------------------- reference import
def import_reference():
    torch.cuda.get_device_capability = lambda *a, **k: (10, 0)
    real_autocast = torch.autocast
    torch.autocast = EmuAutocast
    try:
        for name, path in (("kandinsky", os.path.join(REF, "kandinsky")),
                           ("kandinsky.models", os.path.join(REF, "kandinsky", "models"))):
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
        mods = {}
        for name in ("kandinsky.models.utils", "kandinsky.models.nn", "kandinsky.models.dit",
                     "kandinsky.generation_utils"):
            mods[name.split(".")[-1]] = importlib.import_module(name)
    finally:
        torch.autocast = real_autocast

    def fa(q, k, v):  # flash_attn_func contract: [B,S,H,D] in / out, non-causal, scale d^-0.5
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2)

    mods["nn"].FA = fa

    # flex_attention: on CUDA the reference runs the compiled block-sparse kernel, which visits only
    # the KV blocks listed in the BlockMask (all "full" blocks, nablaT_v2 passes mask_mod=None) with
    # fp32 scores / softmax and bf16 P.V.  The eager CPU fallback ignores the block lists when
    # mask_mod is None (it would compute DENSE attention), so substitute the kernel's semantics.
    def flex_blocksparse(q, k, v, block_mask=None):
        nb, idx = block_mask.full_kv_num_blocks, block_mask.full_kv_indices     # [B,h,nq], [B,h,nq,nk]
        keep = torch.arange(idx.shape[-1])[None, None, None, :] < nb[..., None]
        dense = torch.zeros(idx.shape, dtype=torch.bool).scatter_(-1, idx.long(), keep)
        m = dense.repeat_interleave(64, dim=-2).repeat_interleave(64, dim=-1)
        sc = (q.float() @ k.float().transpose(-1, -2)) * (q.shape[-1] ** -0.5)
        p = torch.softmax(sc.masked_fill(~m, float("-inf")), dim=-1)
        return (p.to(q.dtype).float() @ v.float()).to(q.dtype)

    mods["nn"].flex_attention = flex_blocksparse
    return mods


class Conf:
    """Minimal attribute view of a YAML dict (the reference reads conf.model.dit_params.* etc.)."""

    def __init__(self, d):
        self._d = d

    def __getattr__(self, k):
        v = self._d[k]
        return Conf(v) if isinstance(v, dict) else v


# ----------------------------------------------------------------------------- cases
SMALL = dict(in_visual_dim=16, out_visual_dim=16, time_dim=512, patch_size=(1, 2, 2), model_dim=1792,
             ff_dim=7168, num_text_blocks=2, num_visual_blocks=1, axes_dims=(16, 24, 24), visual_cond=True,
             in_text_dim=3584, in_text_dim2=768)
TINY = dict(in_visual_dim=16, out_visual_dim=16, time_dim=512, patch_size=(1, 2, 2), model_dim=256,
            ff_dim=1024, num_text_blocks=2, num_visual_blocks=2, axes_dims=(16, 24, 24), visual_cond=True,
            in_text_dim=3584, in_text_dim2=768)


def synth_inputs(T, H, W, L, seed=1):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, H, W, 16, generator=g)
    text = torch.randn(L, 3584, generator=g).to(torch.bfloat16)
    pooled = torch.randn(1, 768, generator=g).to(torch.bfloat16)
    return x, text, pooled


def build_ref_model(mods, cfg, sd):
    model = mods["dit"].get_dit(dict(cfg))
    missing = model.load_state_dict(sd, assign=True)  # strict: proves the key/shape contract
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()


def run_forward_case(mods, name, cfg, T, H, W, L, t, nabla=None):
    from oracle import dit_oracle as O

    sd = O.synthetic_state_dict(cfg, seed=0)
    model = build_ref_model(mods, cfg, sd)
    img, text, pooled = synth_inputs(T, H, W, L)
    x = O.model_input(img, True)
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    sparse = None
    if nabla is not None:
        sta = mods["utils"].fast_sta_nabla(T, H // 2 // 8, W // 2 // 8, nabla["wT"], nabla["wH"], nabla["wW"], device="cpu")
        sparse = {"sta_mask": sta[None, None], "to_fractal": True, "P": nabla["P"]}
    time = torch.tensor([t * 1000.0])
    masks = []
    real_nabla = mods["nn"].nablaT_v2

    def recording_nabla(q, k, sta, thr=0.9):
        bm = real_nabla(q, k, sta, thr=thr)
        nb, idx = bm.full_kv_num_blocks[0], bm.full_kv_indices[0]   # [h, nq], [h, nq, nk]
        dense = torch.zeros(idx.shape, dtype=torch.bool)
        keep = torch.arange(idx.shape[-1])[None, None, :] < nb[..., None]
        dense.scatter_(-1, idx.long(), keep)
        masks.append(dense)
        return bm

    mods["nn"].nablaT_v2 = recording_nabla
    try:
        with torch.no_grad(), AutocastEmu(), EmuAutocast("cuda", dtype=torch.bfloat16):
            out = model(x, text, pooled, time, pos, torch.arange(L), scale_factor=(1.0, 2.0, 2.0), sparse_params=sparse)
    finally:
        mods["nn"].nablaT_v2 = real_nabla
    rec = dict(name=name, cfg=cfg, T=T, H=H, W=W, L=L, t=t, nabla=nabla, out=out.clone(),
               scale_factor=(1.0, 2.0, 2.0), weight_seed=0, input_seed=1,
               block_masks=torch.stack(masks) if masks else None)
    torch.save(rec, os.path.join(HERE, name + ".pt"))
    print(name, tuple(out.shape), out.dtype, float(out.float().abs().mean()))


def run_sampler_case(mods, name, cfg, T, H, W, L, Ln, steps, w, sched):
    from oracle import dit_oracle as O

    sd = O.synthetic_state_dict(cfg, seed=0)
    model = build_ref_model(mods, cfg, sd)
    img, text, pooled = synth_inputs(T, H, W, L)
    _, ntext, npooled = synth_inputs(T, H, W, Ln, seed=2)
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    conf = Conf({"metrics": {"scale_factor": (1.0, 2.0, 2.0)},
                 "model": {"dit_params": dict(cfg), "attention": {"type": "flash"}}})
    gu = mods["generation_utils"]
    # generate() draws its noise from torch.Generator("cuda"); feed it ours instead.
    real_randn, real_gen = torch.randn, torch.Generator
    torch.randn = lambda *a, **k: img.clone()
    torch.Generator = lambda device=None: real_gen()
    gu.tqdm = lambda it, **k: it
    try:
        with torch.no_grad(), AutocastEmu(), EmuAutocast("cuda", dtype=torch.bfloat16):
            out = gu.generate(model, "cpu", (T, H, W, 16), steps,
                              {"text_embeds": text, "pooled_embed": pooled},
                              {"text_embeds": ntext, "pooled_embed": npooled},
                              pos, torch.arange(L), torch.arange(Ln), w, sched, conf, seed=6554)
    finally:
        torch.randn, torch.Generator = real_randn, real_gen
    rec = dict(name=name, cfg=cfg, T=T, H=H, W=W, L=L, Ln=Ln, steps=steps, guidance_weight=w,
               scheduler_scale=sched, out=out.clone(), scale_factor=(1.0, 2.0, 2.0), weight_seed=0)
    torch.save(rec, os.path.join(HERE, name + ".pt"))
    print(name, tuple(out.shape), out.dtype, float(out.abs().mean()))


def run_magcache_case(mods, name, cfg, T, H, W, L, Ln, steps, w, sched, mag_ratios):
    """generate() with the reference's MagCache forward (kandinsky/magcache_utils.py) swapped in; records which forwards
    skipped the visual blocks (number of visual-block executions per forward) and the final latent."""
    from oracle import dit_oracle as O

    mc = importlib.import_module("kandinsky.magcache_utils")
    sd = O.synthetic_state_dict(cfg, seed=0)
    model = build_ref_model(mods, cfg, sd)
    cls = model.__class__
    plain_forward = cls.forward
    mc.set_magcache_params(model, list(mag_ratios), steps, abs(w - 1.0) < 1e-6)
    img, text, pooled = synth_inputs(T, H, W, L)
    _, ntext, npooled = synth_inputs(T, H, W, Ln, seed=2)
    pos = [torch.arange(T), torch.arange(H // 2), torch.arange(W // 2)]
    conf = Conf({"metrics": {"scale_factor": (1.0, 2.0, 2.0)},
                 "model": {"dit_params": dict(cfg), "attention": {"type": "flash"}}})
    gu = mods["generation_utils"]
    calls = []
    blk = model.visual_transformer_blocks[0]
    real_blk_forward = blk.forward

    def counting(*a, **k):
        calls[-1] += 1
        return real_blk_forward(*a, **k)

    blk.forward = counting
    mag_forward = cls.forward

    def recording_forward(self, *a, **k):
        calls.append(0)
        return mag_forward(self, *a, **k)

    cls.forward = recording_forward
    real_randn, real_gen = torch.randn, torch.Generator
    torch.randn = lambda *a, **k: img.clone()
    torch.Generator = lambda device=None: real_gen()
    gu.tqdm = lambda it, **k: it
    try:
        with torch.no_grad(), AutocastEmu(), EmuAutocast("cuda", dtype=torch.bfloat16):
            out = gu.generate(model, "cpu", (T, H, W, 16), steps,
                              {"text_embeds": text, "pooled_embed": pooled},
                              {"text_embeds": ntext, "pooled_embed": npooled},
                              pos, torch.arange(L), torch.arange(Ln), w, sched, conf, seed=6554)
    finally:
        torch.randn, torch.Generator = real_randn, real_gen
        cls.forward = plain_forward
    skipped = [c == 0 for c in calls]
    rec = dict(name=name, cfg=cfg, T=T, H=H, W=W, L=L, Ln=Ln, steps=steps, guidance_weight=w, scheduler_scale=sched,
               mag_ratios=list(mag_ratios), skipped=skipped, out=out.clone(), scale_factor=(1.0, 2.0, 2.0), weight_seed=0)
    torch.save(rec, os.path.join(HERE, name + ".pt"))
    print(name, tuple(out.shape), "forwards", len(calls), "skipped", sum(skipped), float(out.abs().mean()))


def main():
    mods = import_reference()
    torch.manual_seed(0)
    # BASELINE.json configs[0]: single DiT block, 1x8x8 token grid (latent 1x16x16), full Lite widths
    run_forward_case(mods, "cfg1_block_1x8x8", SMALL, 1, 16, 16, 24, 0.7)
    # narrow model, several frames, plain token order, odd text length (tail masking)
    run_forward_case(mods, "tiny_flash_3x16x16", TINY, 3, 32, 32, 37, 0.35)
    # NABLA (fractal order + adaptive block mask through flex_attention), 4x16x16 token grid
    run_forward_case(mods, "tiny_nabla_4x16x16", TINY, 4, 32, 32, 24, 0.5,
                     nabla=dict(P=0.6, wT=3, wH=3, wW=3))
    # whole sampler with CFG (two forwards / step), 4 Euler steps
    run_sampler_case(mods, "tiny_sampler_cfg", TINY, 2, 16, 16, 24, 9, steps=4, w=5.0, sched=5.0)
    # MagCache: 10 CFG steps = 20 forwards, calibration curve shaped like configs/config_5s_sft.yaml's (some steps skip)
    ratios = [0.92, 0.915, 0.95, 0.953, 1.05, 1.049, 1.03, 1.031, 1.02, 1.021, 1.02, 1.019, 1.015, 1.016, 1.01, 1.011,
              0.99, 0.989, 0.94, 0.937]
    run_magcache_case(mods, "tiny_sampler_magcache", TINY, 2, 16, 16, 24, 9, steps=10, w=5.0, sched=5.0,
                      mag_ratios=ratios[:18])


if __name__ == "__main__":
    main()
