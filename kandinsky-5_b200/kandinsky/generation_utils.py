"""Drop-in for the reference's `kandinsky/generation_utils.py`: same function names, argument order and
results (`generate` is also called positionally by comfyui/nodes_kandinsky.py:221-226).  The Euler / CFG loop
runs on the device inside libk5 (k5_sample) when the model is the CUDA engine."""
import ctypes

import torch

from ._lib import K5Sparse, check, lib, ptr, stream_ptr
from .models.parallelize import gather_frames


def _get(conf, path):
    """conf may be an OmegaConf node, a plain dict or any attribute container."""
    cur = conf
    for p in path.split("."):
        cur = cur[p] if isinstance(cur, dict) else getattr(cur, p)
    return cur


def sta_block_mask(T, H, W, wT=3, wH=3, wW=3, device="cuda"):
    """Same matrix as the reference's fast_sta_nabla (models/utils.py:108-133): block (t, h, w) sees block (t', h', w')
    iff |t - t'| <= wT // 2, |h - h'| <= wH // 2 and |w - w'| <= wW // 2; bool [T H W, T H W].  On a CUDA device it is
    built by k5_sta_mask, elsewhere by index arithmetic."""
    device = torch.device(device)
    if device.type == "cuda":
        from . import ops

        with torch.cuda.device(device):
            return ops.sta_mask(T, H, W, wT, wH, wW, device=device).bool()
    i = torch.arange(T * H * W, device=device)
    coords = (i // (H * W), (i // W) % H, i % W)
    m = torch.ones(T * H * W, T * H * W, dtype=torch.bool, device=device)
    for c, win in zip(coords, (wT, wH, wW)):
        m &= (c[:, None] - c[None, :]).abs() <= win // 2
    return m


def get_sparse_params(conf, batch_embeds, device):
    """generation_utils.py:10-36, same keys and values ("sta_mask" is the [1, 1, n, n] bool block mask).  The engine
    rebuilds the mask on the device from (wT, wH, wW) and reads only P / to_fractal / the windows."""
    patch = _get(conf, "model.dit_params.patch_size")
    assert patch[0] == 1
    T, H, W, _ = batch_embeds["visual"].shape
    T, H, W = T // patch[0], H // patch[1], W // patch[2]
    att = _get(conf, "model.attention")
    att_type = att["type"] if isinstance(att, dict) else att.type
    if att_type != "nabla":
        return None

    def a(name, default=None):
        if isinstance(att, dict):
            return att.get(name, default)
        return getattr(att, name, default)

    return {
        "sta_mask": sta_block_mask(T, H // 8, W // 8, a("wT"), a("wH"), a("wW"), device=device)[None, None],
        "attention_type": att_type,
        "to_fractal": True,
        "P": a("P"),
        "wT": a("wT"),
        "wW": a("wW"),
        "wH": a("wH"),
        "add_sta": a("add_sta", True),
        "visual_shape": (T, H, W),
        "method": a("method", "topcdf"),
    }


@torch.no_grad()
def get_velocity(dit, x, t, text_embeds, null_text_embeds, visual_rope_pos, text_rope_pos, null_text_rope_pos,
                 guidance_weight, conf, sparse_params=None):
    """generation_utils.py:39-77: one forward, or two + bf16 CFG combine."""
    scale_factor = _get(conf, "metrics.scale_factor")
    v = dit(x, text_embeds["text_embeds"], text_embeds["pooled_embed"], t * 1000, visual_rope_pos, text_rope_pos,
            scale_factor=scale_factor, sparse_params=sparse_params)
    if abs(guidance_weight - 1.0) > 1e-6:
        vu = dit(x, null_text_embeds["text_embeds"], null_text_embeds["pooled_embed"], t * 1000, visual_rope_pos,
                 null_text_rope_pos, scale_factor=scale_factor, sparse_params=sparse_params)
        v = vu + guidance_weight * (v - vu)
    return v


def timesteps(num_steps, scheduler_scale, device):
    """generation_utils.py:102-103."""
    t = torch.linspace(1, 0, num_steps + 1, device=device)
    return scheduler_scale * t / (1 + (scheduler_scale - 1) * t)


@torch.no_grad()
def generate(model, device, shape, num_steps, text_embeds, null_text_embeds, visual_rope_pos, text_rope_pos,
             null_text_rope_pos, guidance_weight, scheduler_scale, conf, progress=False, seed=6554, noise=None):
    """generation_utils.py:80-129.  Returns the fp32 latent [T,H,W,C].  `noise` (optional, not in the reference)
    replaces the torch.Generator draw so that tests can inject identical noise on every backend."""
    device = torch.device(device)
    if noise is None:
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        img = torch.randn(*shape, device=device, generator=g)
    else:
        img = noise.to(device=device, dtype=torch.float32).clone().contiguous()
    sparse_params = get_sparse_params(conf, {"visual": img}, device)
    scale_factor = _get(conf, "metrics.scale_factor")
    arange_pos = (list(text_rope_pos.tolist()) == list(range(len(text_rope_pos)))
                  and list(null_text_rope_pos.tolist()) == list(range(len(null_text_rope_pos))))
    if hasattr(model, "_engine") and arange_pos:
        # whole loop on the device
        T, H, W, _ = img.shape
        fractal = bool(sparse_params["to_fractal"]) if sparse_params is not None else False
        with torch.cuda.device(device):
            model.set_grid((T, H, W), visual_rope_pos, scale_factor, fractal)
            text = text_embeds["text_embeds"].to(device, torch.bfloat16).contiguous()
            pooled = text_embeds["pooled_embed"].to(device, torch.bfloat16).contiguous().view(-1)
            cfg = abs(guidance_weight - 1.0) > 1e-6
            ntext = null_text_embeds["text_embeds"].to(device, torch.bfloat16).contiguous() if cfg else None
            npooled = null_text_embeds["pooled_embed"].to(device, torch.bfloat16).contiguous().view(-1) if cfg else None
            sp = model._sparse_struct(sparse_params)
            args = (model._engine, ptr(img), int(num_steps), float(guidance_weight), float(scheduler_scale), ptr(text),
                    text.shape[0], ptr(pooled), ptr(ntext), 0 if ntext is None else ntext.shape[0], ptr(npooled),
                    ctypes.byref(sp) if sp is not None else None)
            mag = getattr(model, "_magcache", None)
            if mag is None:
                check(lib().k5_sample(*args, stream_ptr()))
            else:
                # the skip decisions are host arithmetic on the calibrated ratios (magcache_utils.py:64-80): the state
                # machine is stepped once per forward the loop will run, in the loop's order, before the loop starts
                sched = (ctypes.c_uint8 * (2 * int(num_steps)))()
                for i in range(int(num_steps)):
                    for _ in range(2 if cfg else 1):
                        slot, skip = mag.next()
                        sched[2 * i + slot] = 1 if skip else 0
                check(lib().k5_sample_magcache(*args, sched, stream_ptr()))
        return gather_frames(model, img)
    # generic path: same loop as the reference, one engine forward per call
    ts = timesteps(num_steps, scheduler_scale, device)
    steps = list(zip(ts[:-1], torch.diff(ts)))
    if progress:
        from tqdm import tqdm

        steps = tqdm(steps)
    for timestep, timestep_diff in steps:
        time = timestep.unsqueeze(0)
        if model.visual_cond:
            visual_cond = torch.zeros_like(img)
            visual_cond_mask = torch.zeros([*img.shape[:-1], 1], dtype=img.dtype, device=img.device)
            model_input = torch.cat([img, visual_cond, visual_cond_mask], dim=-1)
        else:
            model_input = img
        v = get_velocity(model, model_input, time, text_embeds, null_text_embeds, visual_rope_pos, text_rope_pos,
                         null_text_rope_pos, guidance_weight, conf, sparse_params=sparse_params)
        img = img + timestep_diff * v
    return gather_frames(model, img) if hasattr(model, "_engine") else img


def generate_sample(shape, caption, dit, vae, conf, text_embedder, num_steps=25, guidance_weight=5.0, scheduler_scale=1,
                    negative_caption="", seed=6554, device="cuda", vae_device="cuda", text_embedder_device="cuda",
                    progress=True, offload=False):
    """generation_utils.py:132-228: text encode -> denoise -> (optional) VAE decode -> uint8.
    Returns uint8 [1,3,F,H,W] when a VAE is given, else the fp32 latent [T,H,W,C] (VAE decode is not part of
    this round's hot path)."""
    bs, duration, height, width, dim = shape
    type_of_content = "image" if duration == 1 else "video"
    with torch.no_grad():
        bs_text_embed, text_cu_seqlens = text_embedder.encode([caption], type_of_content=type_of_content)
        bs_null_text_embed, null_text_cu_seqlens = text_embedder.encode([negative_caption], type_of_content=type_of_content)
    for key in bs_text_embed:
        bs_text_embed[key] = bs_text_embed[key].to(device=device)
        bs_null_text_embed[key] = bs_null_text_embed[key].to(device=device)
    text_len = int(text_cu_seqlens[-1])
    null_len = int(null_text_cu_seqlens[-1])
    patch = _get(conf, "model.dit_params.patch_size")
    visual_rope_pos = [torch.arange(duration), torch.arange(shape[-3] // patch[1]), torch.arange(shape[-2] // patch[2])]
    latent = generate(dit, device, (bs * duration, height, width, dim), num_steps, bs_text_embed, bs_null_text_embed,
                      visual_rope_pos, torch.arange(text_len), torch.arange(null_len), guidance_weight, scheduler_scale,
                      conf, seed=seed, progress=progress)
    if vae is None:
        return latent
    with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        images = latent.reshape(bs, -1, latent.shape[-3], latent.shape[-2], latent.shape[-1]).to(device=vae_device)
        images = (images / vae.config.scaling_factor).permute(0, 4, 1, 2, 3)
        images = vae.decode(images).sample
        images = ((images.clamp(-1.0, 1.0) + 1.0) * 127.5).to(torch.uint8)
    return images
