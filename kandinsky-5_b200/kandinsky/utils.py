"""Drop-in for the reference's `kandinsky/utils.py` factory: same `get_T2V_pipeline` signature and YAML
schema.  OmegaConf is used when installed, PyYAML otherwise (the YAML files parse unchanged)."""
import os
from typing import Optional, Union

import torch

from .models.dit import get_dit
from .t2v_pipeline import Kandinsky5T2VPipeline


class _Node(dict):
    """dict with attribute access, enough for `conf.model.dit_params.*` style reads."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(obj):
    if isinstance(obj, dict):
        return _Node({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [_wrap(v) for v in obj]
    return obj


def load_conf(conf_path):
    try:
        from omegaconf import OmegaConf

        return OmegaConf.load(conf_path)
    except ImportError:
        import yaml

        with open(conf_path) as f:
            return _wrap(yaml.safe_load(f))


def get_default_conf(dit_path=None, vae_path=None, text_encoder_path=None, text_encoder2_path=None):
    """Mirrors kandinsky/utils.py:137-198 (= configs/config_5s_sft.yaml)."""
    return _wrap({
        "metrics": {"scale_factor": (1, 2, 2)},
        "model": {
            "checkpoint_path": dit_path, "num_steps": 50, "guidance_weight": 5.0,
            "dit_params": {"in_visual_dim": 16, "out_visual_dim": 16, "time_dim": 512, "patch_size": [1, 2, 2],
                           "model_dim": 1792, "ff_dim": 7168, "num_text_blocks": 2, "num_visual_blocks": 32,
                           "axes_dims": [16, 24, 24], "visual_cond": True, "in_text_dim": 3584, "in_text_dim2": 768},
            "attention": {"type": "flash", "causal": False, "local": False, "glob": False, "window": 3},
            "vae": {"checkpoint_path": vae_path, "name": "hunyuan"},
            "text_embedder": {"qwen": {"emb_size": 3584, "checkpoint_path": text_encoder_path, "max_length": 256},
                              "clip": {"checkpoint_path": text_encoder2_path, "emb_size": 768, "max_length": 77}},
        },
    })


def get_T2V_pipeline(device_map: Union[str, torch.device, dict], resolution: int = 512, cache_dir: str = "./weights/",
                     dit_path: Optional[str] = None, text_encoder_path: Optional[str] = None,
                     text_encoder2_path: Optional[str] = None, vae_path: Optional[str] = None,
                     conf_path: Optional[str] = None, offload: bool = False, magcache: bool = False,
                     text_embedder=None, vae=None, state_dict=None, max_tokens: Optional[int] = None
                     ) -> Kandinsky5T2VPipeline:
    """kandinsky/utils.py:23-134.  Extra keyword-only-in-spirit arguments (`text_embedder`, `vae`, `state_dict`,
    `max_tokens`) let callers inject side models / weights; the reference's Hugging Face downloads are out of
    scope (no network) — missing checkpoints raise FileNotFoundError instead."""
    assert resolution in [512]
    if offload:
        raise NotImplementedError("offload moves the reference's torch modules between CPU and GPU; the engine keeps "
                                  "its repacked weights resident in HBM")
    if not isinstance(device_map, dict):
        device_map = {"dit": device_map, "vae": device_map, "text_embedder": device_map}
    try:
        local_rank, world_size = int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    except (KeyError, ValueError):
        local_rank, world_size = 0, 1
    if world_size > 1:
        for k in device_map:
            device_map[k] = torch.device(f"cuda:{local_rank}")
    conf = load_conf(conf_path) if conf_path is not None else get_default_conf(dit_path, vae_path, text_encoder_path,
                                                                                text_encoder2_path)
    if dit_path is not None:
        conf.model.checkpoint_path = dit_path
    params = dict(conf.model.dit_params)
    att_type = conf.model.attention.type
    if max_tokens is None:
        max_tokens = 93696 if att_type == "nabla" else 47616
    dit = get_dit({**{k: (list(v) if isinstance(v, (list, tuple)) or hasattr(v, "__iter__") and not isinstance(v, str) else v)
                      for k, v in params.items()}, "max_tokens": max_tokens})
    if state_dict is None:
        path = conf.model.checkpoint_path
        if path is None or not os.path.exists(path):
            raise FileNotFoundError(f"DiT checkpoint not found: {path} (downloads are not available; pass state_dict=...)")
        from safetensors.torch import load_file

        state_dict = load_file(path)
    if magcache:
        # kandinsky/utils.py:107-113
        from .magcache_utils import set_magcache_params

        set_magcache_params(dit, conf.magcache.mag_ratios, conf.model.num_steps, conf.model.guidance_weight == 1.0)
    dit.load_state_dict(state_dict, assign=True)
    dit = dit.to(device_map["dit"])
    if world_size > 1:
        # kandinsky/utils.py:40-45,80-87: the reference initialises NCCL and applies its tensor-parallel plan here;
        # this engine shards the latent along time instead (models/parallelize.py)
        import torch.distributed as dist

        from .models.parallelize import parallelize_dit

        if not dist.is_initialized():
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local_rank}"))
        dit = parallelize_dit(dit)
    if vae is None:
        # kandinsky/utils.py:118-120: build_vae(conf.model.vae); the engine-backed decoder loads the same diffusers folder
        vae_conf = conf.model.vae
        if vae_path is not None:
            vae_conf.checkpoint_path = vae_path
        ck = vae_conf.checkpoint_path
        if ck is not None and os.path.exists(os.path.join(ck, "vae", "diffusion_pytorch_model.safetensors")):
            from .models.vae import build_vae

            vae = build_vae(vae_conf).to(device_map["vae"])
    if text_embedder is None:
        raise FileNotFoundError("no text embedder: Qwen2.5-VL / CLIP checkpoints are not on disk; pass text_embedder=... "
                                "(contract: .encode(texts, type_of_content) -> ({'text_embeds','pooled_embed'}, cu_seqlens))")
    return Kandinsky5T2VPipeline(device_map=device_map, dit=dit, text_embedder=text_embedder, vae=vae, resolution=resolution,
                                 local_dit_rank=local_rank, world_size=world_size, conf=conf, offload=offload)
