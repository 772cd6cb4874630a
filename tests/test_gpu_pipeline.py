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


def _pipeline(tmp_path, steps, guidance, max_tokens=7 * 32 * 32):
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
                            state_dict=O.synthetic_state_dict(TINY, seed=0), max_tokens=max_tokens)
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
    assert torch.equal(out, again)                                        # same seed, same video, bit for bit
    with pytest.raises(ValueError):
        pipe("x", time_length=1, width=640, height=512)                  # t2v_pipeline.py:122-125


def test_pipeline_image_mode_and_latent_path_agree(tmp_path):
    """time_length = 0 is the image mode (one latent frame): a list of PIL images (t2v_pipeline.py:166-176), written as
    png when save_path is given; the pixels equal decode(generate(...)) done by hand through the public pieces."""
    import numpy as np
    from PIL import Image

    from kandinsky.generation_utils import generate

    pipe = _pipeline(str(tmp_path), steps=2, guidance=1.0)
    png = os.path.join(str(tmp_path), "still.png")
    imgs = pipe("still life", time_length=0, width=768, height=512, seed=11, scheduler_scale=5.0, progress=False,
                save_path=png)
    assert isinstance(imgs, list) and len(imgs) == 1 and imgs[0].size == (768, 512) and imgs[0].mode == "RGB"
    out = torch.from_numpy(np.asarray(imgs[0]).copy()).permute(2, 0, 1)[None, :, None]       # [1, 3, 1, H, W]
    assert np.array_equal(np.asarray(Image.open(png).convert("RGB")), np.asarray(imgs[0]))
    emb, cu = pipe.text_embedder.encode(["still life"])
    negative = ("Static, 2D cartoon, cartoon, 2d animation, paintings, images, worst quality, low quality, ugly, deformed, "
                "walking backwards")                                  # the pipeline's default negative caption
    nemb, ncu = pipe.text_embedder.encode([negative])
    emb = {k: v.cuda() for k, v in emb.items()}
    nemb = {k: v.cuda() for k, v in nemb.items()}
    pos = [torch.arange(1), torch.arange(32), torch.arange(48)]
    # positional call, argument for argument as comfyui/nodes_kandinsky.py:221-226 makes it (under its autocast)
    with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        lat = generate(pipe.dit, "cuda:0", (1, 64, 96, 16), 2, emb, nemb, pos, torch.arange(int(cu[-1])),
                       torch.arange(int(ncu[-1])), 1.0, 5.0, pipe.conf)
    lat11 = generate(pipe.dit, "cuda", (1, 64, 96, 16), 2, emb, nemb, pos, torch.arange(int(cu[-1])), torch.arange(int(ncu[-1])),
                     1.0, 5.0, pipe.conf, seed=11)
    assert lat.shape == lat11.shape == (1, 64, 96, 16) and lat.dtype == torch.float32 and not torch.equal(lat, lat11)
    z = (lat11.reshape(1, 1, 64, 96, 16) / pipe.vae.config.scaling_factor).permute(0, 4, 1, 2, 3)
    vid = ((pipe.vae.decode(z).sample.clamp(-1.0, 1.0) + 1.0) * 127.5).to(torch.uint8)
    assert float((vid.float().cpu() - out.float()).abs().max()) <= 1.0


def test_pipeline_writes_mp4_and_decodes_portrait(tmp_path):
    """save_path for a video (t2v_pipeline.py:177-188) and the 768 x 512 portrait resolution (:122-125), whose latent
    96 x 64 must decode in the default landscape VAE workspace."""
    import cv2

    pipe = _pipeline(str(tmp_path), steps=1, guidance=1.0, max_tokens=7 * 48 * 32)
    mp4 = os.path.join(str(tmp_path), "clip.mp4")
    out = pipe("a tall waterfall", time_length=1, width=512, height=768, seed=3, scheduler_scale=5.0, progress=False,
               save_path=[mp4])
    assert out.dtype == torch.uint8 and tuple(out.shape) == (1, 3, 25, 768, 512)
    cap = cv2.VideoCapture(mp4)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 25
    assert (int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))) == (512, 768)
    ok, frame = cap.read()
    cap.release()
    assert ok and abs(float(frame.mean()) - float(out[0, :, 0].float().mean())) < 8.0
    # a list of the wrong length writes nothing and still returns the video (:180)
    miss = os.path.join(str(tmp_path), "none.mp4")
    pipe("a tall waterfall", time_length=1, width=512, height=768, seed=3, scheduler_scale=5.0, progress=False,
         save_path=[miss, miss])
    assert not os.path.exists(miss)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_checkpoint_files_round_trip_through_the_factory(tmp_path, dtype):
    """kandinsky/utils.py:89-120: get_T2V_pipeline(conf_path=<YAML in the reference's schema>) reads the DiT from a
    .safetensors file (load_file + load_state_dict(assign=True)) and the VAE from a diffusers folder, whatever dtype
    the files were saved in.  The engine loaded from a file must equal the engine loaded from the same tensors in
    memory, bit for bit, and stay within bf16 noise of the fp32 weights."""
    import json

    import yaml
    from safetensors.torch import save_file

    from kandinsky import get_T2V_pipeline
    from kandinsky.models.dit import DiffusionTransformer3D

    tmp = str(tmp_path)
    sd32 = O.synthetic_state_dict(TINY, seed=0)
    sd = {k: v.to(dtype).contiguous() for k, v in sd32.items()}
    os.makedirs(os.path.join(tmp, "model"))
    os.makedirs(os.path.join(tmp, "vae_root", "vae"))
    ckpt = os.path.join(tmp, "model", "kandinsky5lite_t2v_tiny.safetensors")
    save_file(sd, ckpt)
    vsd = {k: v.to(torch.float16 if dtype != torch.float32 else dtype).contiguous()
           for k, v in VO.synthetic_state_dict(VAE_WIDTHS, seed=0).items()}          # the published VAE file is fp16
    save_file(vsd, os.path.join(tmp, "vae_root", "vae", "diffusion_pytorch_model.safetensors"))
    with open(os.path.join(tmp, "vae_root", "vae", "config.json"), "w") as f:
        json.dump({"_class_name": "AutoencoderKLHunyuanVideo", "block_out_channels": list(VAE_WIDTHS), "latent_channels": 16,
                   "scaling_factor": 0.476986}, f)
    conf = {"metrics": {"scale_factor": [1.0, 2.0, 2.0], "resolution": 512},
            "model": {"checkpoint_path": ckpt, "num_steps": 2, "guidance_weight": 5.0,
                      "dit_params": dict(TINY, patch_size=[1, 2, 2], axes_dims=[16, 24, 24]),
                      "attention": {"type": "flash", "causal": False, "local": False, "glob": False, "window": 3},
                      "vae": {"checkpoint_path": os.path.join(tmp, "vae_root"), "name": "hunyuan"},
                      "text_embedder": {"qwen": {"emb_size": 3584, "checkpoint_path": "./weights/text_encoder/", "max_length": 256},
                                        "clip": {"checkpoint_path": "./weights/text_encoder2/", "emb_size": 768, "max_length": 77}}},
            "magcache": {"mag_ratios": [1.0] * 8}}
    path = os.path.join(tmp, "config_tiny.yaml")
    with open(path, "w") as f:
        yaml.safe_dump(conf, f)
    pipe = get_T2V_pipeline({"dit": "cuda:0", "vae": "cuda:0", "text_embedder": "cuda:0"}, conf_path=path,
                            text_embedder=FakeTextEmbedder(), max_tokens=32 * 32)
    assert pipe.vae is not None and abs(pipe.vae.config.scaling_factor - 0.476986) < 1e-9
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 16, 16, 33, generator=g).cuda()
    text = torch.randn(24, 3584, generator=g).to(torch.bfloat16).cuda()
    pooled = torch.randn(1, 768, generator=g).to(torch.bfloat16).cuda()
    pos = [torch.arange(3), torch.arange(8), torch.arange(8)]

    def fwd(model):
        return model(x, text, pooled, torch.tensor([700.0]), pos, torch.arange(24), scale_factor=(1.0, 2.0, 2.0))

    def direct(state):
        m = DiffusionTransformer3D(**TINY, max_tokens=32 * 32)
        m.load_state_dict(state, assign=True)
        return fwd(m.to("cuda"))

    from_file = fwd(pipe.dit)
    assert torch.equal(from_file, direct(sd))
    ref = direct(sd32).float()
    assert float((from_file.float() - ref).norm() / ref.norm()) < (1e-6 if dtype == torch.float32 else 3e-2)
    out = pipe("a fox", time_length=0, width=512, height=512, seed=1, num_steps=1, progress=False)
    assert isinstance(out, list) and out[0].size == (512, 512)
    with pytest.raises(FileNotFoundError):
        conf["model"]["checkpoint_path"] = os.path.join(tmp, "missing.safetensors")
        with open(path, "w") as f:
            yaml.safe_dump(conf, f)
        get_T2V_pipeline("cuda:0", conf_path=path, text_embedder=FakeTextEmbedder())
