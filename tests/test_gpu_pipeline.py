"""The drop-in surface end to end on a real B200: get_T2V_pipeline -> Kandinsky5T2VPipeline.__call__ -> generate_sample
-> k5_sample -> vae.decode -> uint8 video, with a stand-in text embedder (Qwen2.5-VL / CLIP are outside the hot path)
and narrow random-init DiT / VAE weights.  Checks the reference's output contract (kandinsky/t2v_pipeline.py:90-189,
generation_utils.py:132-228) and that the result equals the same computation done step by step through the oracle-free
public pieces (generate + vae.decode)."""
import os

import pytest
import torch

from oracle import dit_oracle as O
from oracle import vae_oracle as VO

pytestmark = pytest.mark.gpu

TINY = dict(in_visual_dim=16, out_visual_dim=16, time_dim=512, patch_size=(1, 2, 2), model_dim=256, ff_dim=1024,
            num_text_blocks=2, num_visual_blocks=2, axes_dims=(16, 24, 24), visual_cond=True, in_text_dim=3584,
            in_text_dim2=768)
VAE_WIDTHS = (64, 64, 128, 128)


class FakeTextEmbedder:
    """Contract of kandinsky/models/text_embedders.py::Kandinsky5TextEmbedder.encode as generate_sample uses it."""

    def encode(self, texts, type_of_content="video"):
        g = torch.Generator().manual_seed(len(texts[0]) + 7)
        n = 24 if texts[0] else 9
        emb = {"text_embeds": torch.randn(n, 3584, generator=g).to(torch.bfloat16),
               "pooled_embed": torch.randn(1, 768, generator=g).to(torch.bfloat16)}
        return emb, torch.tensor([0, n], dtype=torch.int32)


def _pipeline(tmp_path, steps, guidance):
    import yaml

    from kandinsky import get_T2V_pipeline
    from kandinsky.models.vae import AutoencoderKLHunyuanVideo

    conf = {"metrics": {"scale_factor": [1.0, 2.0, 2.0]},
            "model": {"checkpoint_path": None, "num_steps": steps, "guidance_weight": guidance, "dit_params": dict(TINY, patch_size=[1, 2, 2], axes_dims=[16, 24, 24]),
                      "attention": {"type": "flash", "causal": False, "local": False, "glob": False, "window": 3},
                      "vae": {"checkpoint_path": None, "name": "hunyuan"},
                      "text_embedder": {"qwen": {"emb_size": 3584, "checkpoint_path": None, "max_length": 256},
                                        "clip": {"checkpoint_path": None, "emb_size": 768, "max_length": 77}}}}
    path = os.path.join(tmp_path, "conf.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(conf, f)
    vae = AutoencoderKLHunyuanVideo(block_out_channels=VAE_WIDTHS, max_latent=(5, 64, 96))
    vae.load_state_dict(VO.synthetic_state_dict(VAE_WIDTHS, seed=0))
    vae.to("cuda")
    pipe = get_T2V_pipeline("cuda:0", conf_path=path, text_embedder=FakeTextEmbedder(), vae=vae,
                            state_dict=O.synthetic_state_dict(TINY, seed=0), max_tokens=7 * 32 * 32)
    return pipe


def test_pipeline_returns_uint8_video_with_the_reference_shape(tmp_path):
    pipe = _pipeline(str(tmp_path), steps=3, guidance=5.0)
    out = pipe("a red fox running through snow", time_length=1, width=512, height=512, seed=6554, scheduler_scale=5.0,
               progress=False)
    # time_length 1 s -> 7 latent frames -> 25 video frames (t2v_pipeline.py:150-160), two (17, 8) temporal tiles
    assert out.dtype == torch.uint8 and tuple(out.shape) == (1, 3, 25, 512, 512)
    assert 5.0 < float(out.float().mean()) < 250.0 and float(out.float().std()) > 1.0
    again = pipe("a red fox running through snow", time_length=1, width=512, height=512, seed=6554, scheduler_scale=5.0,
                 progress=False)
    assert float((out.float() - again.float()).abs().max()) <= 1.0     # same seed, same video (GroupNorm atomics: +-1 level)
    with pytest.raises(ValueError):
        pipe("x", time_length=1, width=640, height=512)                  # t2v_pipeline.py:122-125


def test_pipeline_image_mode_and_latent_path_agree(tmp_path):
    """time_length = 0 is the image mode (one latent frame); the pipeline's video equals decode(generate(...))."""
    from kandinsky.generation_utils import generate

    pipe = _pipeline(str(tmp_path), steps=2, guidance=1.0)
    out = pipe("still life", time_length=0, width=768, height=512, seed=11, scheduler_scale=5.0, progress=False)
    assert out.dtype == torch.uint8 and tuple(out.shape) == (1, 3, 1, 512, 768)
    emb, cu = pipe.text_embedder.encode(["still life"])
    negative = ("Static, 2D cartoon, cartoon, 2d animation, paintings, images, worst quality, low quality, ugly, deformed, "
                "walking backwards")                                  # the pipeline's default negative caption
    nemb, ncu = pipe.text_embedder.encode([negative])
    emb = {k: v.cuda() for k, v in emb.items()}
    nemb = {k: v.cuda() for k, v in nemb.items()}
    pos = [torch.arange(1), torch.arange(32), torch.arange(48)]
    lat = generate(pipe.dit, "cuda", (1, 64, 96, 16), 2, emb, nemb, pos, torch.arange(int(cu[-1])), torch.arange(int(ncu[-1])),
                   1.0, 5.0, pipe.conf, seed=11)
    z = (lat.reshape(1, 1, 64, 96, 16) / pipe.vae.config.scaling_factor).permute(0, 4, 1, 2, 3)
    vid = ((pipe.vae.decode(z).sample.clamp(-1.0, 1.0) + 1.0) * 127.5).to(torch.uint8)
    assert float((vid.float() - out.float()).abs().max()) <= 1.0
